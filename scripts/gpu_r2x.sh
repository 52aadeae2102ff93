# round 2, call X: A/B of the idle-warp poll back-off; parity suite with the L2-only scratch accesses
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for rep in 1 2; do
for cfg in ":" "poll20:" "poll200:"; do
  tag=${cfg%%:*}
  ( export LZ_LIB_TAG=$tag; [ -z "$tag" ] && unset LZ_LIB_TAG; timeout 200 python tests/gpu_time_search.py 2>&1 | tail -1 )
done; done | tee gpurun_out/x_ab.log
( timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -n 4 ) | cut -c1-200
