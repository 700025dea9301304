#!/bin/bash
# Threads per subtree CTA of the kd-tree's bottom phase (1,024-point subtrees), and two CTAs per SM (512-point subtrees).
cd "$(dirname "$0")/.."
for bt in 1024 512 256; do
  echo "== MCMC_GPU_KD_BT=$bt"
  MCMC_GPU_KD_BT=$bt python tools/bench_evidence.py --reps 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('lebesgue', round(1e3*d['lebesgue_s'],2), 'tree64', round(1e3*d['tree64_s'],2), 'full', round(1e3*d['tree_full_s'],2))"
done
for kb in 110 75; do
  echo "== MCMC_GPU_KD_SMEM_KB=$kb"
  MCMC_GPU_KD_SMEM_KB=$kb python tools/bench_evidence.py --reps 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('lebesgue', round(1e3*d['lebesgue_s'],2), 'tree64', round(1e3*d['tree64_s'],2), 'full', round(1e3*d['tree_full_s'],2))"
done
