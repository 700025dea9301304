set -x
cd $GRAFT_REPO_ROOT
MCMC_GPU_DEBUG=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/bench_dist_build.py --reps 4 > gpurun_out/r2_run20_dbg.json 2> gpurun_out/r2_run20_dbg.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 tools/bench_dist_build.py --check --reps 6 > gpurun_out/r2_run20_dist2.json 2> gpurun_out/r2_run20_dist2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29553 tools/bench_dist_build.py --check --reps 4 --min-split 64 > gpurun_out/r2_run20_dist2_ms64.json 2> gpurun_out/r2_run20_dist2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29554 tools/bench_dist_build.py --check --reps 4 --n 3000000 --d 4 > gpurun_out/r2_run20_dist2_d4.json 2> gpurun_out/r2_run20_dist2.err
CUDA_VISIBLE_DEVICES=0 timeout 300 python tools/stress_tree.py --seconds 30 --seed 20 > gpurun_out/r2_run20_stress.log 2>&1
echo finished
