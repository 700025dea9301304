set -x
cd $GRAFT_REPO_ROOT
timeout 300 python tools/multi_gpu_check.py --out gpurun_out/r02_multi_gpu_check_1.json > gpurun_out/r2_run2_mg1.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py --out gpurun_out/r02_multi_gpu_check_2.json > gpurun_out/r2_run2_mg2.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/multi_gpu_check.py --n 10000000 --d 20 --out gpurun_out/r02_multi_gpu_check_2_cfg3.json > gpurun_out/r2_run2_mg2b.log 2>&1
echo finished
