#!/bin/bash
# ncu launch list of the config-3 pipeline at N = 1e6 (kernel shares; run under gpurun)
set -e
python tools/bench_evidence.py --n 1000000 --d 20 --queries 1000000 --reps 1 > gpurun_out/ev_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r01_evidence_launches.csv \
    python tools/bench_evidence.py --n 1000000 --d 20 --queries 1000000 --reps 1 > gpurun_out/ev_ncu.log 2>&1
