# round 2, call I: canonical softmax order (fused == step-wise again?) + ncu source-level profile of the persistent kernel
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -n 12 ) > gpurun_out/i_pytest.log 2>&1
( timeout 200 python tests/gpu_debug_search.py ) > gpurun_out/i_phases.log 2>&1
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_net_tc -s 2 -c 1 -f -o gpurun_out/prof_net_tc_r02i python tests/gpu_profile_search.py ) > gpurun_out/i_ncu.log 2>&1
tail -12 gpurun_out/i_pytest.log | cut -c1-200
cat gpurun_out/i_phases.log | cut -c1-260
tail -5 gpurun_out/i_ncu.log
ls -la gpurun_out/prof_net_tc_r02i.ncu-rep
