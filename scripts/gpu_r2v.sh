# round 2, call V: parity report at the north-star sizes + full bench line (frame-stack e2e arm)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 900 python tests/gpu_parity_report.py ) > gpurun_out/v_parity.log 2>&1
tail -40 gpurun_out/v_parity.log | cut -c1-220
( timeout 900 python bench.py --steps 10 --warmup 3 ) > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err
python - <<'PY'
import json
try:
    b=json.loads(open("gpurun_out/v_bench.json").read().strip().splitlines()[-1])
    print("bench", round(b["value"]), round(b["ms_per_step"],3), b["config"].get("search_only_ms"), "e2e", round(b["e2e"]["value"]), round(b["e2e"]["ms_per_step"],3), b["e2e"]["h2d_bytes_per_step"], "full", b["e2e"].get("full_stack_upload",{}).get("ms_per_step"), b.get("gpu_launches"), b["roofline"]["frac"])
    print(json.dumps(b.get("strong_scaling"))[:500]); print(json.dumps(b.get("extra"))[:700]); print(json.dumps(b.get("cpu_baseline")))
except Exception as e: print("ERR", e, open("gpurun_out/v_bench.err").read()[-2000:])
PY
