# round 2, call G: heads v2 (FC1 lane mapping, fused softmax, mbarrier release, 5-stage ring) + epilogue prefetch
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 200 python tests/gpu_debug_search.py ) > gpurun_out/g_phases.log 2>&1
( timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -n 25 ) > gpurun_out/g_pytest.log 2>&1
( timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline ) > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err
cat gpurun_out/g_phases.log | cut -c1-260
tail -25 gpurun_out/g_pytest.log | cut -c1-200
python - <<'PY'
import json
for f in ("g_bench",):
    try:
        b=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(b["value"]), round(b["ms_per_step"],3), b["config"].get("search_only_ms"), b["e2e"]["value"], b["roofline"]["frac"], b["roofline"]["kernel_ms"])
    except Exception as e: print(f, "ERR", e, open(f"gpurun_out/{f}.err").read()[-800:])
PY
