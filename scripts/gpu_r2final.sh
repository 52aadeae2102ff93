# round 2, final measurement session: smoke, full parity suite, full bench line (not under a profiler), launch list, ncu --set full, phases, CUPTI trace
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2 ) | cut -c1-200
( timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -n 3 ) | cut -c1-200
( timeout 900 python bench.py --steps 20 --warmup 3 ) > gpurun_out/fin_bench.json 2> gpurun_out/fin_bench.err
( timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/fin_bench_ref.json 2> gpurun_out/fin_bench_ref.err
( timeout 300 python tests/gpu_debug_search.py ) > gpurun_out/fin_phases.log 2>&1
( timeout 300 python tests/gpu_trace_step.py ) > gpurun_out/fin_trace.log 2>&1
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/fin_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras ) > gpurun_out/fin_ncu_list.log 2>&1
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_net_tc -s 1 -c 1 -f -o gpurun_out/prof_net_tc_r02fin python tests/gpu_profile_search.py ) > gpurun_out/fin_ncu.log 2>&1
cat gpurun_out/fin_phases.log | cut -c1-250
python - <<'PY'
import json
for f in ("fin_bench","fin_bench_ref"):
    try:
        b=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(b["value"]), round(b.get("ms_per_step",0),3), (b.get("config") or {}).get("search_only_ms"), "e2e", round(b["e2e"]["value"]), b["e2e"].get("ms_per_step"), b.get("gpu_launches"), (b.get("roofline") or {}).get("frac"), (b.get("clocks") or {}))
    except Exception as e: print(f, "ERR", e, open(f"gpurun_out/{f}.err").read()[-1500:])
PY
