set -x
cd $GRAFT_REPO_ROOT
CUDA_VISIBLE_DEVICES=0 timeout 300 python tools/stress_tree.py --seconds 40 --seed 16 > gpurun_out/r2_run17_stress.log 2>&1
MCMC_GPU_DEBUG=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/bench_dist_build.py --reps 3 > gpurun_out/r2_run17_dbg.json 2> gpurun_out/r2_run17_dbg.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 tools/bench_dist_build.py --check > gpurun_out/r2_run17_dist2.json 2> gpurun_out/r2_run17_dist2.err
echo finished
