#!/bin/bash
# ncu --set full captures of the other hot-path kernels (run under gpurun, after each plain run exits 0)
set -e
python tools/bench_rjmcmc.py --chains 262144 --steps 200 --ntree 2000000 > gpurun_out/rj_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rj_ensemble -s 1 -c 1 -o gpurun_out/r01_rj \
    python tools/bench_rjmcmc.py --chains 262144 --steps 200 --ntree 2000000 > gpurun_out/rj_ncu.log 2>&1
python tools/bench_evidence.py --n 2000000 --d 20 --queries 2000000 --reps 1 > gpurun_out/ev2_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'jump_prob_kernel|draw_kernel|part_fused|rs_scatter|node_split|cell_terms' -c 12 -o gpurun_out/r01_tree \
    python tools/bench_evidence.py --n 2000000 --d 20 --queries 2000000 --reps 1 > gpurun_out/ev2_ncu.log 2>&1
python tools/bench_nested.py --nlive 20000 --nmcmc 200 --batch 2048 > gpurun_out/nest_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nest_replace -s 2 -c 1 -o gpurun_out/r01_nest \
    python tools/bench_nested.py --nlive 20000 --nmcmc 200 --batch 2048 > gpurun_out/nest_ncu.log 2>&1
