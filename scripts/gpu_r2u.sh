# round 2, call U: UniZero driver, fused EfficientZero search_with_reuse, device collector state; full parity suite
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -n 40 ) > gpurun_out/u_pytest.log 2>&1
tail -40 gpurun_out/u_pytest.log | cut -c1-250
