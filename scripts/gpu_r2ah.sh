# round 2, call AH: compute-sanitizer memcheck over every kernel family (tiny sizes + one 893-root batch for the root-group split)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 120 python tests/gpu_sanitize.py 2>&1 | tail -n 6 ) | cut -c1-200
( timeout 900 compute-sanitizer --tool memcheck python tests/gpu_sanitize.py 2>&1 | tail -n 12 ) > gpurun_out/ah_memcheck.log 2>&1
cat gpurun_out/ah_memcheck.log | cut -c1-220
