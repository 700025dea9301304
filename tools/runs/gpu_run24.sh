set -x
cd $GRAFT_REPO_ROOT
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tools/multi_gpu_check.py --samples 10000000 --dim 20 --out gpurun_out/r02_multi_gpu_check_2_run24.json > gpurun_out/r2_run24_mg2.log 2>&1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 tools/multi_gpu_check.py --samples 200000 --dim 5 --out gpurun_out/r02_multi_gpu_check_2_small_run24.json > gpurun_out/r2_run24_mg2s.log 2>&1
echo finished
