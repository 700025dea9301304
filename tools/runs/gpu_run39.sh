set -x
cd $GRAFT_REPO_ROOT
CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests/test_evidence_gpu.py tests/test_kdtree_gpu.py tests/test_full_size_gpu.py tests/test_misc_gpu.py -x -q > gpurun_out/r2_run39_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_run39_tests.log
CUDA_VISIBLE_DEVICES=0 timeout 300 python tools/bench_evidence.py --reps 3 > gpurun_out/r2_run39_cfg3.json 2> gpurun_out/r2_run39_cfg3.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29591 tools/multi_gpu_check.py --samples 10000000 --dim 20 --out gpurun_out/r02_multi_gpu_check_2_run39.json > gpurun_out/r2_run39_mg2.log 2>&1
echo finished
