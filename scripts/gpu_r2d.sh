# round 2, call D: fold variants (phases) + correctness of the streamed heads (f0 build)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for f in 0 1 2 3; do
  ( LZ_LIB_TAG=f$f timeout 200 python tests/gpu_debug_search.py ) > gpurun_out/d_phases_f$f.log 2>&1
done
( LZ_LIB_TAG=f0 timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -n 30 ) > gpurun_out/d_pytest_f0.log 2>&1
( LZ_LIB_TAG=f0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline ) > gpurun_out/d_bench_f0.json 2> gpurun_out/d_bench_f0.err
for f in 0 1 2 3; do echo "== f$f"; cat gpurun_out/d_phases_f$f.log; done
tail -30 gpurun_out/d_pytest_f0.log
python - <<'PY'
import json
for f in ("d_bench_f0",):
    try:
        b=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(b["value"]), round(b["ms_per_step"],3), b["config"].get("search_only_ms"), b["e2e"]["value"], b["roofline"]["frac"], b["roofline"]["kernel_ms"])
    except Exception as e: print(f, "ERR", e, open(f"gpurun_out/{f}.err").read()[-1500:])
PY
