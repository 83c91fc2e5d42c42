#!/bin/bash
# every bench workload once, short, kernel time + roofline fraction per line
for w in trading_step trading_rollout trading_nas100_step np_step np_nas100_step portfolio_step cashpenalty_step stoploss_step; do
  python bench.py --workload $w --steps ${STEPS:-300} --warmup 10 --no-cpu --e2e-steps 3 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']
    print('$w', 'value=%.3e'%d['value'], 'kernel_ms=%.4f'%r['kernel_ms'], 'GB/s=%.0f'%r['achieved'], 'frac=%.3f'%r['frac'], 'e2e=%.3e'%d['e2e']['value'], d['clocks'])
except Exception as e:
    print('$w FAILED', e)"
done
