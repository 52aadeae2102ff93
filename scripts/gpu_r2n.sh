# round 2, call N: root-group split {5,2} (epilogue of one group under the other group's MMAs) + joint heads read-out
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 300 python tests/gpu_debug_search.py ) > gpurun_out/n_phases.log 2>&1
( LZ_TC_SPLIT=0 timeout 300 python tests/gpu_debug_search.py ) > gpurun_out/n_phases_nosplit.log 2>&1
( timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -n 25 ) > gpurun_out/n_pytest.log 2>&1
( timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras ) > gpurun_out/n_bench.json 2> gpurun_out/n_bench.err
cat gpurun_out/n_phases.log | cut -c1-260
echo "--- LZ_TC_SPLIT=0"
cat gpurun_out/n_phases_nosplit.log | cut -c1-260
tail -8 gpurun_out/n_pytest.log | cut -c1-200
python - <<'PY'
import json
for f in ("n_bench",):
    try:
        b=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(b["value"]), round(b["ms_per_step"],3), b["config"].get("search_only_ms"), b["e2e"]["value"], b["e2e"]["ms_per_step"], b.get("gpu_launches"), b["roofline"]["frac"])
    except Exception as e: print(f, "ERR", e, open(f"gpurun_out/{f}.err").read()[-1500:])
PY
