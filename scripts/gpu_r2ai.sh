# round 2, call AI: EfficientZero launch list + full test suite on the current code
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -n 4 ) | cut -c1-200
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 420 --csv --log-file gpurun_out/ai_launches_ez.csv python bench.py --workload efficientzero --steps 2 --warmup 1 --no-cpu-baseline --no-extras ) > gpurun_out/ai_ncu_list.log 2>&1
python profiles/summarize.py gpurun_out/ai_launches_ez.csv | head -24
