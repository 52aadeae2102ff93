# round 2, call L: tensor-core heads (tag tch): tests + phases; e2e trace; ncu of two tower layers
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export LZ_LIB_TAG=tch
( timeout 300 python tests/gpu_debug_search.py ) > gpurun_out/l_phases.log 2>&1
( timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -n 25 ) > gpurun_out/l_pytest.log 2>&1
( timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras ) > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err
( timeout 300 python tests/gpu_trace_step.py ) > gpurun_out/l_trace.log 2>&1
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 0 -c 2 -f -o gpurun_out/prof_conv_tc_r02l python tests/gpu_profile_search.py ) > gpurun_out/l_ncu.log 2>&1
cat gpurun_out/l_phases.log | cut -c1-260
tail -25 gpurun_out/l_pytest.log | cut -c1-200
grep -A40 "end-to-end step" gpurun_out/l_trace.log | cut -c1-120
python - <<'PY'
import json
for f in ("l_bench",):
    try:
        b=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(b["value"]), round(b["ms_per_step"],3), b["config"].get("search_only_ms"), b["e2e"]["value"], b["e2e"]["ms_per_step"], b.get("gpu_launches"), b["roofline"]["frac"])
    except Exception as e: print(f, "ERR", e, open(f"gpurun_out/{f}.err").read()[-1500:])
PY
