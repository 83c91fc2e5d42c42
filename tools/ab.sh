#!/bin/bash
# A/B of one bench workload: every library under variants/ against the working tree's, interleaved twice.
#   WL=np_nas100_step bash tools/ab.sh
WL=${WL:-trading_step}
run() { python bench.py --workload $WL --steps ${STEPS:-200} --warmup 20 --no-cpu --e2e-steps 3 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%-28s' % '$1', 'kernel_ms', round(d['roofline']['kernel_ms'],5), 'frac', round(d['roofline']['frac'],4))"; }
for rep in 1 2; do
for lib in variants/*.so; do FINRL_B200_LIB=$PWD/$lib run $(basename $lib .so); done
run "working tree"
done
