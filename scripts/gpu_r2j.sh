# round 2, call J: coalesced internal latent layout; where does the non-search part of a step go
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 200 python tests/gpu_debug_search.py ) > gpurun_out/j_phases.log 2>&1
( timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -n 12 ) > gpurun_out/j_pytest.log 2>&1
( timeout 300 python tests/gpu_debug_step.py ) > gpurun_out/j_step.log 2>&1
( timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline ) > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err
cat gpurun_out/j_phases.log | cut -c1-260
tail -12 gpurun_out/j_pytest.log | cut -c1-200
cat gpurun_out/j_step.log
python - <<'PY'
import json
for f in ("j_bench",):
    try:
        b=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(b["value"]), round(b["ms_per_step"],3), b["config"].get("search_only_ms"), b["e2e"]["value"], b["roofline"]["frac"], b["roofline"]["kernel_ms"])
    except Exception as e: print(f, "ERR", e, open(f"gpurun_out/{f}.err").read()[-800:])
PY
