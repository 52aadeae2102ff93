# round 2, call C: MMA shape-mix probe + correctness of fold / global skip / register-resident tree phase
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 300 python tests/gpu_debug_mma.py ) > gpurun_out/c_mma_probe.log 2>&1
( timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -n 30 ) > gpurun_out/c_pytest.log 2>&1
( LZ_TC_GENERIC_TREE=1 timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -n 30 ) > gpurun_out/c_pytest_generic_tree.log 2>&1
( timeout 300 python tests/gpu_debug_search.py ) > gpurun_out/c_phases.log 2>&1
cat gpurun_out/c_mma_probe.log; tail -30 gpurun_out/c_pytest.log; tail -5 gpurun_out/c_pytest_generic_tree.log; cat gpurun_out/c_phases.log
