cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all python tests/gpu_sanitize.py ) > gpurun_out/san_racecheck_full.log 2>&1
grep -c "Race reported\|hazard" gpurun_out/san_racecheck_full.log
grep -A3 "hazard detected" gpurun_out/san_racecheck_full.log | head -30; tail -n 3 gpurun_out/san_racecheck_full.log
