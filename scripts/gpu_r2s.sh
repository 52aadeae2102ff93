# round 2, call S: fast exp in the canonical softmax, unrolled head scatter; parity suite + timing + phases
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for cfg in ":0" ":5"; do
  tag=${cfg%%:*}; sp=${cfg##*:}
  ( export LZ_LIB_TAG=$tag; [ -z "$tag" ] && unset LZ_LIB_TAG; [ -n "$sp" ] && export LZ_TC_SPLIT=$sp; timeout 200 python tests/gpu_time_search.py 2>&1 | tail -1 )
done | tee gpurun_out/s_ab.log
( timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -n 25 ) > gpurun_out/s_pytest.log 2>&1
tail -12 gpurun_out/s_pytest.log | cut -c1-220
( timeout 300 python tests/gpu_debug_search.py ) > gpurun_out/s_phases.log 2>&1
cat gpurun_out/s_phases.log | cut -c1-260
