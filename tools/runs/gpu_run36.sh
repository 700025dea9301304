set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_run36_bench8.json 2> gpurun_out/r2_run36_bench8.err
MCMC_GPU_KDD_LEVELS_KERNEL=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29582 tools/multi_gpu_check.py --samples 400000 --dim 6 --out gpurun_out/r02_multi_gpu_check_8_small_run36.json > gpurun_out/r2_run36_mg8s.log 2>&1
echo finished
