# round 2, call AE: DownSample tower with the [B_hi | B_lo] fold for N <= 64; parity + tower timing
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -q -m gpu -x -k "model or search or ez or reanalyze or collector" 2>&1 | tail -n 8 ) | cut -c1-220
( timeout 300 python tests/gpu_trace_step.py ) > gpurun_out/ae_trace.log 2>&1
grep -A13 "^initial_inference" gpurun_out/ae_trace.log | cut -c1-130
