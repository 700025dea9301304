set -x
cd $GRAFT_REPO_ROOT
for md in 8 64; do
  MCMC_GPU_DRAW_CACHE_MAXD=$md timeout 300 python tools/bench_rjmcmc.py --da 8 --db 16 --chains 262144 > gpurun_out/r2_run25_rj_8_16_maxd$md.json 2>> gpurun_out/r2_run25.err
  MCMC_GPU_DRAW_CACHE_MAXD=$md timeout 300 python tools/bench_rjmcmc.py --da 32 --db 64 --chains 262144 > gpurun_out/r2_run25_rj_32_64_maxd$md.json 2>> gpurun_out/r2_run25.err
done
timeout 300 python tools/bench_rjmcmc.py > gpurun_out/r2_run25_rj_2_4.json 2>> gpurun_out/r2_run25.err
timeout 200 python tools/bench_nested.py > gpurun_out/r2_run25_nested.json 2>> gpurun_out/r2_run25.err
echo finished
