#!/bin/bash
# one ncu --set full capture per kd-tree / Interpolate_pdf / evidence / stats kernel (run under gpurun)
set -e
CMD="python tools/bench_evidence.py --n 4000000 --d 20 --queries 4000000 --reps 1"
$CMD > gpurun_out/ev4_plain.log 2>&1
for k in part_fused_kernel jump_prob_kernel draw_kernel node_split_kernel cell_terms_kernel rs_scatter_kernel rs_hist_kernel mark_side_kernel; do
  # -s: skip the small warm-up launches; capture a late (large) launch of each kernel
  skip=2; [ "$k" = part_fused_kernel ] && skip=30; [ "$k" = rs_scatter_kernel ] && skip=24; [ "$k" = rs_hist_kernel ] && skip=24
  [ "$k" = node_split_kernel ] && skip=30; [ "$k" = mark_side_kernel ] && skip=30
  ncu --set full --clock-control none -k regex:$k -s $skip -c 1 -o gpurun_out/r01_k_$k $CMD > gpurun_out/ev4_ncu_$k.log 2>&1 || true
done
python bench.py --steps 1 --warmup 1 --no-cpu --no-evidence > gpurun_out/b_plain.log 2>&1
ncu --set full --clock-control none -k regex:block_field_moments -s 1 -c 1 -o gpurun_out/r01_k_moments python bench.py --steps 1 --warmup 1 --no-cpu --no-evidence > gpurun_out/b_ncu.log 2>&1 || true
ls -la gpurun_out/r01_k_*.ncu-rep
