#!/bin/bash
# How the MH kernel's pass time depends on the number of resident warps per scheduler
# (592 schedulers x k one-warp CTAs = 18944 k chains).  Config 2 itself is 65,536 chains = 3.46 per scheduler.
for c in 18944 37888 56832 65536 75776; do
  python bench.py --chains $c --steps 3 --warmup 3 --no-cpu --no-evidence 2>/dev/null | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print($c, 'ms', round(d['ms_per_step'],3), 'steps/s', '%.4g' % d['value'])"
done
