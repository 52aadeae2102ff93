cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_ez.py tests/test_gpu_tree.py tests/test_gpu_search.py -x -q -m gpu 2>&1 | tail -n 4 ) > gpurun_out/q_pytest.log 2>&1
( timeout 600 python bench.py --workload efficientzero --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/q_bench_ez.json 2> gpurun_out/q_bench_ez.err
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 210 --csv --log-file gpurun_out/q_launches_ez.csv python bench.py --workload efficientzero --steps 1 --warmup 3 --no-cpu-baseline ) > gpurun_out/q_ncu.log 2>&1
tail -n 2 gpurun_out/q_pytest.log
python - <<'PY'
import json
b=json.loads(open("gpurun_out/q_bench_ez.json").read().strip().splitlines()[-1])
print(round(b["value"]), round(b["ms_per_step"],3), b["config"]["search_only_ms"], b["e2e"]["value"], b["roofline"]["frac"])
PY
python profiles/summarize.py gpurun_out/q_launches_ez.csv 2>/dev/null | head -7
