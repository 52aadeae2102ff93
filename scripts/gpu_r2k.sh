# round 2, call K: new tests + new bench.py + CUPTI trace of a step
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -n 25 ) > gpurun_out/k_pytest.log 2>&1
( timeout 300 python tests/gpu_trace_step.py ) > gpurun_out/k_trace.log 2>&1
( timeout 600 python bench.py --steps 5 --warmup 3 ) > gpurun_out/k_bench.json 2> gpurun_out/k_bench.err
( timeout 300 python bench.py --impl reference --steps 2 --warmup 3 ) > gpurun_out/k_bench_ref.json 2> gpurun_out/k_bench_ref.err
tail -25 gpurun_out/k_pytest.log | cut -c1-200
cat gpurun_out/k_trace.log | cut -c1-220 | head -60
python - <<'PY'
import json
for f in ("k_bench","k_bench_ref"):
    try:
        b=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(b["value"]), round(b["ms_per_step"],3), b["config"].get("search_only_ms"), b["e2e"], b.get("gpu_launches"), (b.get("roofline") or {}).get("frac"))
        print("   strong:", json.dumps(b.get("strong_scaling"))[:600])
        print("   extra:", json.dumps(b.get("extra"))[:700])
        print("   cpu_baseline:", json.dumps(b.get("cpu_baseline"))[:300])
    except Exception as e: print(f, "ERR", e, open(f"gpurun_out/{f}.err").read()[-1500:])
PY
