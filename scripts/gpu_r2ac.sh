# round 2, call AC: CTA stagger A/B (de-phasing the CTAs' L2-bandwidth-bound weight-stream phases)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for st in 0 4000 12000 25000 40000; do ( export LZ_TC_STAGGER=$st; echo -n "stagger=$st "; timeout 200 python tests/gpu_time_search.py 2>&1 | tail -1 ); done | tee gpurun_out/ac_ab.log
