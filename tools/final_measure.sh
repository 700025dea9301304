#!/bin/bash
# End-of-round measurement on one B200 (run under gpurun): GPU test suite, smoke, the default bench line,
# the reference (CPU) arm and the ncu launch list of the bench command.
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r01_bench_final.json 2> gpurun_out/r01_bench_final.err
tail -c 400 gpurun_out/r01_bench_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01_bench_reference.json 2>/dev/null
python bench.py --steps 3 --warmup 3 --no-cpu --no-evidence > gpurun_out/plain_ll.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_bench_launches_final.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-evidence > gpurun_out/ncu_ll.log 2>&1
python -c "
import json
d=json.load(open('gpurun_out/r01_bench_final.json'))
print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'traffic',d['roofline']['traffic'],'e2e',d['e2e']['value'],'cpu',d.get('cpu_baseline',{}).get('value'),'ev',d.get('evidence',{}).get('lebesgue_samples_per_s'))
r=json.load(open('gpurun_out/r01_bench_reference.json')); print('reference',r['value'],r['cpu_baseline']['sample'])
"
