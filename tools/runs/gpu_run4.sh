set -x
cd $GRAFT_REPO_ROOT
python tools/profile_lebesgue.py 10000000 20 > gpurun_out/r2_run4_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02_lebesgue_v2_launches.csv python tools/profile_lebesgue.py 10000000 20 > gpurun_out/r2_run4_ncu.log 2>&1
echo finished
