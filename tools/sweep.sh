#!/bin/bash
# usage: tools/sweep.sh "<bench args>" ...   -- prints one summary line per configuration
for cfg in "$@"; do
  python bench.py --steps 5 --no-cpu-baseline --e2e-steps 0 $cfg 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']
    print('$cfg', '| MS/s %.0f' % d['value'], '| frac %.3f' % r['frac'], '| kern ms %.3f' % r['kernel_ms_per_launch'], '| step ms %.3f' % d['ms_per_step'], '| clk', d['clocks']['sm_mhz'])
except Exception as e:
    print('$cfg', 'FAILED', e)
"
done
