# round 2, call AG: conv_tc with 8 epilogue warps + hoisted residual loads; parity, tower phases, tower timing
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -q -m gpu -x -k "model or search or ez or reanalyze or collector" 2>&1 | tail -n 6 ) | cut -c1-220
( timeout 300 python tests/gpu_debug_tower.py 2>&1 | tail -9 ) | cut -c1-200
( timeout 300 python tests/gpu_trace_step.py ) > gpurun_out/ag_trace.log 2>&1
grep -A13 "^initial_inference" gpurun_out/ag_trace.log | cut -c1-130
