set -x
cd $GRAFT_REPO_ROOT
timeout 900 python tools/parity_full_size.py --n 10000000 --d 20 --dups 0.0 --out gpurun_out/r02_parity_cfg3_1e7_run38.json > gpurun_out/r2_run38_parity.log 2>&1
timeout 300 python tools/stress_parity.py --seconds 120 --seed 38 > gpurun_out/r2_run38_stress.log 2>&1
echo finished
