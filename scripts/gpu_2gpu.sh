# weak-scaling check on 2 GPUs of one box (the driver runs 1/2/4/8 itself at round end)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for wl in muzero efficientzero; do
  ( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload $wl --steps 10 --warmup 3 ) > gpurun_out/s2_bench_$wl.json 2> gpurun_out/s2_bench_$wl.err
done
python - <<'PY'
import json
for wl in ("muzero","efficientzero"):
    try:
        b=json.loads(open(f"gpurun_out/s2_bench_{wl}.json").read().strip().splitlines()[-1])
        print(wl, b["n_gpus"], round(b["value"]), round(b["ms_per_step"],3), b["e2e"]["value"])
    except Exception as e: print(wl, "ERR", e, open(f"gpurun_out/s2_bench_{wl}.err").read()[-1200:])
PY
