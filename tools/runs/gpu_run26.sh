set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_run26_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_run26_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_run26_smoke.log 2>&1
( time timeout 900 python bench.py > gpurun_out/r2_run26_bench.json 2> gpurun_out/r2_run26_bench.err ) 2> gpurun_out/r2_run26_bench.time
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_run26_ref.json 2> gpurun_out/r2_run26_ref.err ) 2> gpurun_out/r2_run26_ref.time
echo finished
