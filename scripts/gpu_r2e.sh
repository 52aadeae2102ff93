# round 2, call E: divergence hypothesis (all-lane mbarrier waits in the MMA warp) + ring wait accounting
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for f in f1w f1o f0w; do
  ( LZ_LIB_TAG=$f timeout 200 python tests/gpu_debug_search.py ) > gpurun_out/e_phases_$f.log 2>&1
done
( LZ_LIB_TAG=f1w timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -n 12 ) > gpurun_out/e_pytest_f1w.log 2>&1
( LZ_LIB_TAG=f1w timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline ) > gpurun_out/e_bench_f1w.json 2> gpurun_out/e_bench_f1w.err
for f in f1w f1o f0w; do echo "== $f"; cat gpurun_out/e_phases_$f.log | cut -c1-260; done
tail -12 gpurun_out/e_pytest_f1w.log | cut -c1-200
python - <<'PY'
import json
for f in ("e_bench_f1w",):
    try:
        b=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(b["value"]), round(b["ms_per_step"],3), b["config"].get("search_only_ms"), b["e2e"]["value"], b["roofline"]["frac"], b["roofline"]["kernel_ms"])
    except Exception as e: print(f, "ERR", e, open(f"gpurun_out/{f}.err").read()[-800:])
PY
