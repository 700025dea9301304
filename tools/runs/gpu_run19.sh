set -x
cd $GRAFT_REPO_ROOT
MCMC_GPU_DEBUG=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/bench_dist_build.py --reps 3 > gpurun_out/r2_run19_dbg.json 2> gpurun_out/r2_run19_dbg.err
echo finished
