# round 2, call A: baseline phase stamps + uniform-issue variant validation
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_gpu.txt 2>&1
( timeout 300 python tests/gpu_debug_search.py ) > gpurun_out/a_phases_default.log 2>&1
( timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline ) > gpurun_out/a_bench_default.json 2> gpurun_out/a_bench_default.err
export LZ_LIB_TAG=uni
( timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 8 ) > gpurun_out/a_pytest_uni.log 2>&1
( timeout 300 python tests/gpu_debug_search.py ) > gpurun_out/a_phases_uni.log 2>&1
( timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline ) > gpurun_out/a_bench_uni.json 2> gpurun_out/a_bench_uni.err
( timeout 300 python bench.py --workload efficientzero --steps 5 --warmup 3 --no-cpu-baseline ) > gpurun_out/a_bench_uni_ez.json 2> gpurun_out/a_bench_uni_ez.err
cat gpurun_out/a_phases_default.log; cat gpurun_out/a_phases_uni.log; tail -3 gpurun_out/a_pytest_uni.log
python - <<'PY'
import json
for f in ("a_bench_default","a_bench_uni","a_bench_uni_ez"):
    try:
        b=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(b["value"]), round(b["ms_per_step"],3), b["config"].get("search_only_ms"), b["e2e"]["value"], b["roofline"]["frac"], b["roofline"]["kernel_ms"])
    except Exception as e: print(f, "ERR", e, open(f"gpurun_out/{f}.err").read()[-1500:])
PY
