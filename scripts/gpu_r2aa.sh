# round 2, call AA: ncu --set full of the two <32> tower layers and the <128> layer
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( DBG_N=1 DBG_S=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 0 -c 3 -f -o gpurun_out/prof_conv_tc_r02aa python tests/gpu_profile_search.py ) > gpurun_out/aa_ncu.log 2>&1
tail -3 gpurun_out/aa_ncu.log
