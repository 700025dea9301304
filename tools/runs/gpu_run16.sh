set -x
cd $GRAFT_REPO_ROOT
export MCMC_GPU_SKIP_FULL_PARITY=1
CUDA_VISIBLE_DEVICES=0 timeout 600 python -m pytest tests/test_kdtree_gpu.py tests/test_evidence_gpu.py tests/test_ellipse_gpu.py -x -q > gpurun_out/r2_run16_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_run16_tests.log
CUDA_VISIBLE_DEVICES=0 timeout 300 python tools/stress_tree.py --seconds 30 --seed 16 > gpurun_out/r2_run16_stress.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/multi_gpu_check.py --samples 10000000 --dim 20 --out gpurun_out/r02_multi_gpu_check_2_run16.json > gpurun_out/r2_run16_mg2.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 tools/multi_gpu_check.py --samples 300000 --dim 3 --out gpurun_out/r02_multi_gpu_check_2_small_run16.json > gpurun_out/r2_run16_mg2s.log 2>&1
echo finished
