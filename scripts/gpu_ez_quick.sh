cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_ez.py -x -q -m gpu 2>&1 | tail -n 3 ) > gpurun_out/q_pytest.log 2>&1
for i in 1 2 3; do
( timeout 600 python bench.py --workload efficientzero --steps 20 --warmup 3 --no-cpu-baseline ) > gpurun_out/q_bench_ez_$i.json 2> gpurun_out/q_bench_ez.err
done
tail -n 1 gpurun_out/q_pytest.log
python - <<'PY'
import json
for i in (1,2,3):
    b=json.loads(open(f"gpurun_out/q_bench_ez_{i}.json").read().strip().splitlines()[-1])
    print(round(b["value"]), round(b["ms_per_step"],3), round(b["config"]["search_only_ms"],3), round(b["e2e"]["value"]))
PY
