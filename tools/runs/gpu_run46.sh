set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_run46_bench2.json 2> gpurun_out/r2_run46_bench2.err
echo finished
