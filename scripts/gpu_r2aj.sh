cd $GRAFT_REPO_ROOT
for cfg in "0:0" "1:6000" "1:10000" "1:14000" "2:6000" "2:10000" "2:14000"; do
  ( export LZ_CONV_STAGGER=${cfg%%:*} LZ_CONV_STAGGER_CYCLES=${cfg##*:}; [ "${cfg%%:*}" = "0" ] && unset LZ_CONV_STAGGER; timeout 120 python tests/gpu_time_tower.py 2>&1 | tail -1 )
done
