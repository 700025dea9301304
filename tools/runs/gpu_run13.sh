set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_run13_bench8.json 2> gpurun_out/r2_run13_bench8.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 tools/multi_gpu_check.py --samples 10000000 --dim 20 --out gpurun_out/r02_multi_gpu_check_8_cfg3.json > gpurun_out/r2_run13_mg8.log 2>&1
echo finished
