#!/bin/bash
# A/B: run bench.py against each tuning variant of the library (same ABI), print kernel time.
for lib in default finrl_b200/build/ab/*.so; do
  if [ "$lib" = default ]; then unset FINRL_B200_LIB; else export FINRL_B200_LIB=$PWD/$lib; fi
  python bench.py --workload ${WL:-trading_step} --steps ${STEPS:-600} --warmup 20 --no-cpu --e2e-steps 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$lib', 'kernel_ms=%.4f'%r['kernel_ms'], 'frac=%.3f'%r['frac'], 'value=%.3e'%d['value'], d['clocks'])"
done
