cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ez_lstm_tc -s 20 -c 1 -o gpurun_out/prof_ez_lstm_tc_r01d -f python bench.py --workload efficientzero --steps 1 --warmup 3 --no-cpu-baseline ) > gpurun_out/ncu_full_ez.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
( timeout 600 python -m pytest tests/test_gpu_reanalyze.py -x -q -m gpu 2>&1 | tail -n 3 )
