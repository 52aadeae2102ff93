# round 2, call AB: reward-hook features consumed one layer later; timing, phases, parity suite
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for rep in 1 2; do ( timeout 200 python tests/gpu_time_search.py 2>&1 | tail -1 ); done | tee gpurun_out/ab_ab.log
( timeout 300 python tests/gpu_debug_search.py ) 2>&1 | cut -c1-260
( timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -n 4 ) | cut -c1-200
