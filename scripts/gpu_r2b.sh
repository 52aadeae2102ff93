# round 2, call B: fold (N=128) + global skip + channels-last pool
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 25 ) > gpurun_out/b_pytest.log 2>&1
( timeout 300 python tests/gpu_debug_search.py ) > gpurun_out/b_phases.log 2>&1
( timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline ) > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err
( timeout 300 python bench.py --workload efficientzero --steps 5 --warmup 3 --no-cpu-baseline ) > gpurun_out/b_bench_ez.json 2> gpurun_out/b_bench_ez.err
cat gpurun_out/b_phases.log; tail -25 gpurun_out/b_pytest.log
python - <<'PY'
import json
for f in ("b_bench","b_bench_ez"):
    try:
        b=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(b["value"]), round(b["ms_per_step"],3), b["config"].get("search_only_ms"), b["e2e"]["value"], b["roofline"]["frac"], b["roofline"]["kernel_ms"])
    except Exception as e: print(f, "ERR", e, open(f"gpurun_out/{f}.err").read()[-1500:])
PY
