set -x
cd $GRAFT_REPO_ROOT
timeout 120 python tools/profile_lebesgue.py 10000000 20 > gpurun_out/r2_run47_plain.log 2>&1 || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_run47_lebesgue_launches.csv python tools/profile_lebesgue.py 10000000 20 > gpurun_out/r2_run47_ncu_l.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_run47_tree_launches.csv python tools/profile_tree.py 10000000 20 > gpurun_out/r2_run47_ncu_t.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:v2_bottom -s 1 -c 1 -f -o gpurun_out/r02b_v2_bottom_full python tools/profile_tree.py 10000000 20 > gpurun_out/r2_run47_ncu_bt.log 2>&1
echo finished
