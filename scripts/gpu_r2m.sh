# round 2, call M (session 2 baseline): phases, full parity suite, full bench, launch list, ncu --set full of the persistent kernel, CUPTI trace
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 300 python tests/gpu_debug_search.py ) > gpurun_out/m_phases.log 2>&1
( timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -n 25 ) > gpurun_out/m_pytest.log 2>&1
( timeout 600 python bench.py --steps 10 --warmup 3 ) > gpurun_out/m_bench.json 2> gpurun_out/m_bench.err
( timeout 300 python tests/gpu_trace_step.py ) > gpurun_out/m_trace.log 2>&1
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/m_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras ) > gpurun_out/m_ncu_list.log 2>&1
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_net_tc -s 1 -c 1 -f -o gpurun_out/prof_net_tc_r02m python tests/gpu_profile_search.py ) > gpurun_out/m_ncu.log 2>&1
cat gpurun_out/m_phases.log | cut -c1-260
tail -8 gpurun_out/m_pytest.log | cut -c1-200
grep -A30 "end-to-end step" gpurun_out/m_trace.log | cut -c1-140 | head -50
python - <<'PY'
import json
for f in ("m_bench",):
    try:
        b=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(b["value"]), round(b["ms_per_step"],3), b["config"].get("search_only_ms"), b["e2e"]["value"], b["e2e"]["ms_per_step"], b.get("gpu_launches"), b["roofline"]["frac"])
        print(json.dumps(b.get("strong_scaling"))[:600]); print(json.dumps(b.get("extra"))[:900])
    except Exception as e: print(f, "ERR", e, open(f"gpurun_out/{f}.err").read()[-1500:])
PY
