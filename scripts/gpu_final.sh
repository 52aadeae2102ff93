# final session of the round: sanitizer, launch list, full suite + benches
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 700 compute-sanitizer --tool memcheck python tests/gpu_sanitize.py 2>&1 | tail -n 4 ) > gpurun_out/f_memcheck.log 2>&1
( timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python tests/gpu_sanitize.py 2>&1 | tail -n 4 ) > gpurun_out/f_racecheck.log 2>&1
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 210 --csv --log-file gpurun_out/f_launches_ez.csv python bench.py --workload efficientzero --steps 1 --warmup 3 --no-cpu-baseline ) > gpurun_out/f_ncu.log 2>&1
bash scripts/gpu_round.sh
tail -n 3 gpurun_out/f_memcheck.log gpurun_out/f_racecheck.log
