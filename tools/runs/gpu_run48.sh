set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_run48_bench8.json 2> gpurun_out/r2_run48_bench8.err
echo finished
