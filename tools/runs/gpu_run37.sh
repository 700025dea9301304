set -x
cd $GRAFT_REPO_ROOT
timeout 300 python bench.py --steps 1 --warmup 1 --no-evidence --no-rjmcmc --no-cpu > gpurun_out/r2_run37_plain.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mh_balanced -s 1 -c 1 -f -o gpurun_out/r02b_mh160 python bench.py --steps 1 --warmup 1 --no-evidence --no-rjmcmc --no-cpu > gpurun_out/r2_run37_ncu_mh.log 2>&1
timeout 300 python tools/bench_ellipse.py --n 2000000 --reps 1 --cpu-n 20000 > gpurun_out/r2_run37_ell_plain.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:el_cov_partial -s 3 -c 1 -f -o gpurun_out/r02b_el_cov python tools/bench_ellipse.py --n 2000000 --reps 1 --cpu-n 20000 > gpurun_out/r2_run37_ncu_el1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:el_scatter -s 3 -c 1 -f -o gpurun_out/r02b_el_scatter python tools/bench_ellipse.py --n 2000000 --reps 1 --cpu-n 20000 > gpurun_out/r2_run37_ncu_el2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_run37_ellipse_launches.csv python tools/bench_ellipse.py --n 2000000 --reps 1 --cpu-n 20000 > gpurun_out/r2_run37_ncu_el3.log 2>&1
echo finished
