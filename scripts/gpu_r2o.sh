# round 2, call O: A/B of kernel variants (uninstrumented search time)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for cfg in "head:" "oldho:0" "oldho:5" ":0" ":5"; do
  tag=${cfg%%:*}; sp=${cfg##*:}
  ( export LZ_LIB_TAG=$tag; [ -z "$tag" ] && unset LZ_LIB_TAG; [ -n "$sp" ] && export LZ_TC_SPLIT=$sp; timeout 200 python tests/gpu_time_search.py 2>&1 | tail -1 )
done | tee gpurun_out/o_ab.log
