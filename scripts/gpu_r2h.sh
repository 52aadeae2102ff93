# round 2, call H: epilogue in 16-channel halves, parked tree state, position-independent bias add; launch list of a bench step
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 200 python tests/gpu_debug_search.py ) > gpurun_out/h_phases.log 2>&1
( timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -n 25 ) > gpurun_out/h_pytest.log 2>&1
( timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline ) > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 60 --csv --log-file gpurun_out/h_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline ) > gpurun_out/h_ncu.log 2>&1
cat gpurun_out/h_phases.log | cut -c1-260
tail -25 gpurun_out/h_pytest.log | cut -c1-200
python - <<'PY'
import json, csv, collections
for f in ("h_bench",):
    try:
        b=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(b["value"]), round(b["ms_per_step"],3), b["config"].get("search_only_ms"), b["e2e"]["value"], b["roofline"]["frac"], b["roofline"]["kernel_ms"])
    except Exception as e: print(f, "ERR", e, open(f"gpurun_out/{f}.err").read()[-800:])
try:
    rows=[r for r in csv.reader(open("gpurun_out/h_launches.csv")) if len(r)>5 and r[0].isdigit()]
    for r in rows[:40]: print(r[4][:60], r[-1])
except Exception as e: print("launch list ERR", e)
PY
