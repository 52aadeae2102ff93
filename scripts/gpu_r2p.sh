# round 2, call P: A/B (head vs rolled read-out + single tree copy, split 0/5), parity suite, phases
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for cfg in "head:" ":0" ":5" ":4"; do
  tag=${cfg%%:*}; sp=${cfg##*:}
  ( export LZ_LIB_TAG=$tag; [ -z "$tag" ] && unset LZ_LIB_TAG; [ -n "$sp" ] && export LZ_TC_SPLIT=$sp; timeout 200 python tests/gpu_time_search.py 2>&1 | tail -1 )
done | tee gpurun_out/p_ab.log
( timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -n 25 ) > gpurun_out/p_pytest.log 2>&1
tail -4 gpurun_out/p_pytest.log | cut -c1-200
( timeout 300 python tests/gpu_debug_search.py ) > gpurun_out/p_phases.log 2>&1
cat gpurun_out/p_phases.log | cut -c1-260
