# one GPU session: smoke, full parity suite, both bench workloads (with cpu_baseline), reference arm for EfficientZero
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 3 ) > gpurun_out/r_smoke.log 2>&1
( timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 5 ) > gpurun_out/r_pytest.log 2>&1
( timeout 600 python bench.py --steps 10 --warmup 3 ) > gpurun_out/r_bench_muzero.json 2> gpurun_out/r_bench_muzero.err
( timeout 600 python bench.py --workload efficientzero --steps 10 --warmup 3 ) > gpurun_out/r_bench_ez.json 2> gpurun_out/r_bench_ez.err
cat gpurun_out/r_smoke.log; tail -n 2 gpurun_out/r_pytest.log
python - <<'PY'
import json
for f in ("r_bench_muzero","r_bench_ez"):
    try:
        b=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(b["value"]), round(b["ms_per_step"],3), b.get("config",{}).get("search_only_ms"), b.get("e2e",{}).get("value"), (b.get("roofline") or {}).get("frac"), (b.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(f, "ERR", e, open(f"gpurun_out/{f}.err").read()[-1500:])
PY
