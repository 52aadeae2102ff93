# round 2, call F: warp-uniform role branches (descriptors in uniform registers, no ELECT / R2UR per MMA)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 200 python tests/gpu_debug_search.py ) > gpurun_out/f_phases_u1.log 2>&1
for f in u0 u2; do
  ( LZ_LIB_TAG=$f timeout 200 python tests/gpu_debug_search.py ) > gpurun_out/f_phases_$f.log 2>&1
done
( timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -n 12 ) > gpurun_out/f_pytest.log 2>&1
( timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline ) > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
( timeout 300 python bench.py --workload efficientzero --steps 5 --warmup 3 --no-cpu-baseline ) > gpurun_out/f_bench_ez.json 2> gpurun_out/f_bench_ez.err
for f in u1 u0 u2; do echo "== $f"; cat gpurun_out/f_phases_$f.log | cut -c1-260; done
tail -12 gpurun_out/f_pytest.log | cut -c1-200
python - <<'PY'
import json
for f in ("f_bench","f_bench_ez"):
    try:
        b=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(b["value"]), round(b["ms_per_step"],3), b["config"].get("search_only_ms"), b["e2e"]["value"], b["roofline"]["frac"], b["roofline"]["kernel_ms"])
    except Exception as e: print(f, "ERR", e, open(f"gpurun_out/{f}.err").read()[-800:])
PY
