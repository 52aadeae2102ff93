# round 2, call R: ring retention experiment (group Y starts with the taps group X left in the ring)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for cfg in ":5" "yret:5" "yret:0"; do
  tag=${cfg%%:*}; sp=${cfg##*:}
  ( export LZ_LIB_TAG=$tag; [ -z "$tag" ] && unset LZ_LIB_TAG; [ -n "$sp" ] && export LZ_TC_SPLIT=$sp; timeout 200 python tests/gpu_time_search.py 2>&1 | tail -1 )
done | tee gpurun_out/r_ab.log
( export LZ_LIB_TAG=yret; timeout 600 python -m pytest tests -q -m gpu -x -k "search or model" 2>&1 | tail -n 5 ) | cut -c1-200
( export LZ_LIB_TAG=yret; timeout 300 python tests/gpu_debug_search.py ) 2>&1 | cut -c1-260
