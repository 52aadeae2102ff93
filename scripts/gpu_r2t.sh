# round 2, call T: new stem kernel (weights as kernel-parameter operands); model/search parity, tower trace
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -q -m gpu -x -k "model or search or ez or reanalyze" 2>&1 | tail -n 8 ) | cut -c1-200
( timeout 300 python tests/gpu_trace_step.py ) > gpurun_out/t_trace.log 2>&1
grep -A14 "end-to-end step" gpurun_out/t_trace.log | cut -c1-140
grep -B2 -A12 "initial_inference" gpurun_out/t_trace.log | head -40 | cut -c1-140
