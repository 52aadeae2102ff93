# round 2, call W: A/B: joint two-head read-out, L2-only scratch accesses
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for rep in 1 2; do
for cfg in "nojoint:" "cg:" ":"; do
  tag=${cfg%%:*}
  ( export LZ_LIB_TAG=$tag; [ -z "$tag" ] && unset LZ_LIB_TAG; timeout 200 python tests/gpu_time_search.py 2>&1 | tail -1 )
done; done | tee gpurun_out/w_ab.log
( timeout 900 python -m pytest tests -q -m gpu -x -k "search or model" 2>&1 | tail -n 4 ) | cut -c1-200
