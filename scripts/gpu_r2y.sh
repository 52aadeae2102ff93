# round 2, call Y: profiles of the current state: launch list of a bench run, ncu --set full of the persistent kernel, phases, CUPTI trace
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 300 python tests/gpu_debug_search.py ) > gpurun_out/y_phases.log 2>&1
cat gpurun_out/y_phases.log | cut -c1-260
( timeout 300 python tests/gpu_trace_step.py ) > gpurun_out/y_trace.log 2>&1
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/y_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras ) > gpurun_out/y_ncu_list.log 2>&1
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_net_tc -s 1 -c 1 -f -o gpurun_out/prof_net_tc_r02y python tests/gpu_profile_search.py ) > gpurun_out/y_ncu.log 2>&1
tail -2 gpurun_out/y_ncu.log
grep -c k_net_tc gpurun_out/y_launches.csv
