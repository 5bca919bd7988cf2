#!/bin/bash
# e2e (host-fed) job time against the number of mid-feed sweeps (KATGPU_MID_SWEEPS), one line per setting
for m in 1 2 3 4; do
  KATGPU_MID_SWEEPS=$m timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']
print('mid sweeps $m: resident %.1f ms  e2e %.1f ms (%.2f G k-mers/s)' % (d['ms_per_step'], e['ms_per_step'], e['value']/1e9))"
done
