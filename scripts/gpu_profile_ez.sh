cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 420 --csv --log-file gpurun_out/launches_r01d_ez.csv python bench.py --workload efficientzero --steps 2 --warmup 3 --no-cpu-baseline ) > gpurun_out/ncu_ez.log 2>&1
( timeout 900 compute-sanitizer --tool memcheck python tests/gpu_sanitize.py 2>&1 | tail -n 6 ) > gpurun_out/san_memcheck_r01d.log 2>&1
( timeout 1200 compute-sanitizer --tool racecheck python tests/gpu_sanitize.py 2>&1 | tail -n 6 ) > gpurun_out/san_racecheck_r01d.log 2>&1
tail -n 4 gpurun_out/san_memcheck_r01d.log gpurun_out/san_racecheck_r01d.log
python profiles/summarize.py gpurun_out/launches_r01d_ez.csv | head -20
