#!/bin/bash
# End-of-round measurements (run under gpurun from the repo root).  One GPU:
#   gpurun -- 'bash tools/final_measure.sh'
# several GPUs (N = 2, 4 or 8):
#   gpurun --gpus N -- 'bash tools/final_measure.sh N'
# Records land in gpurun_out/; copy the ones that matter to profiles/rNN/.
N=${1:-1}
if [ "$N" = 1 ]; then
  python -m pytest tests -m gpu -x -q 2>&1 | tail -3
  python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
  python bench.py > gpurun_out/final_bench_1gpu.json 2> gpurun_out/final_bench_1gpu.err
  python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference.json 2>/dev/null
  python tools/bench_evidence.py --reps 3 > gpurun_out/final_cfg3.json 2>/dev/null            # config 3 and the Interpolate_pdf kernels
  python tools/bench_nested.py > gpurun_out/final_cfg4.json 2>/dev/null                        # config 4
  python tools/bench_rjmcmc.py > gpurun_out/final_cfg5_2_4.json 2>/dev/null                    # config 5, (2,4)-D
  python tools/bench_cfg1.py > gpurun_out/final_cfg1.json 2>/dev/null                          # config 1 at its stated size
  python tools/bench_ellipse.py > gpurun_out/final_ellipse.json 2>/dev/null
  python tools/stress_tree.py --seconds 60 > gpurun_out/final_stress_tree.log 2>&1             # random shapes against the oracle
  python tools/stress_parity.py --seconds 120 > gpurun_out/final_stress_parity.log 2>&1
  # launch list of the bench command (never quote a number measured under ncu)
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_bench_launches.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu --no-evidence --no-rjmcmc > gpurun_out/final_ncu.log 2>&1
else
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
  $TR --master-port 29701 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/final_bench_${N}gpu.json 2> gpurun_out/final_bench_${N}gpu.err
  $TR --master-port 29702 tools/multi_gpu_check.py --samples 10000000 --dim 20 --out gpurun_out/final_multi_gpu_check_${N}.json > gpurun_out/final_mg${N}.log 2>&1
  $TR --master-port 29703 tools/bench_dist_build.py --check --reps 5 > gpurun_out/final_dist_build_${N}.json 2>/dev/null
fi
echo finished
