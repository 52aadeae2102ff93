# round 2, call Q: hand-off through shared memory, staged FC2 bias, pipelined mean_q; A/B, parity suite, phases, ncu
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for cfg in "head:" ":0" ":5"; do
  tag=${cfg%%:*}; sp=${cfg##*:}
  ( export LZ_LIB_TAG=$tag; [ -z "$tag" ] && unset LZ_LIB_TAG; [ -n "$sp" ] && export LZ_TC_SPLIT=$sp; timeout 200 python tests/gpu_time_search.py 2>&1 | tail -1 )
done | tee gpurun_out/q_ab.log
( timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -n 25 ) > gpurun_out/q_pytest.log 2>&1
tail -4 gpurun_out/q_pytest.log | cut -c1-200
( timeout 300 python tests/gpu_debug_search.py ) > gpurun_out/q_phases.log 2>&1
cat gpurun_out/q_phases.log | cut -c1-260
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_net_tc -s 1 -c 1 -f -o gpurun_out/prof_net_tc_r02q python tests/gpu_profile_search.py ) > gpurun_out/q_ncu.log 2>&1
tail -3 gpurun_out/q_ncu.log
