set -x
cd $GRAFT_REPO_ROOT
export MCMC_GPU_SKIP_FULL_PARITY=1
CUDA_VISIBLE_DEVICES=0 timeout 600 python -m pytest tests/test_misc_gpu.py tests/test_nested_gpu.py tests/test_rjmcmc_gpu.py -x -q > gpurun_out/r2_run8_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_run8_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_run8_bench2.json 2> gpurun_out/r2_run8_bench2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 tools/multi_gpu_check.py --samples 10000000 --dim 20 --out gpurun_out/r02_multi_gpu_check_2_cfg3.json > gpurun_out/r2_run8_mg2b.log 2>&1
echo finished
