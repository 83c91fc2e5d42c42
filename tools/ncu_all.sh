#!/bin/bash
# full ncu capture of the dominant kernel of each bench workload (one launch each, after a plain run)
set -u
declare -A PAT=( [trading_step]=trading_rollout [np_step]=np_rollout [portfolio_step]=portfolio_rollout [cashpenalty_step]=cashpenalty_rollout [stoploss_step]=stoploss_rollout [trading_nas100_step]=trading_wide [np_nas100_step]=np_wide )
for w in ${WORKLOADS:-trading_step np_step portfolio_step cashpenalty_step stoploss_step}; do
  python bench.py --workload $w --steps 6 --warmup 3 --no-cpu --e2e-steps 3 > gpurun_out/plain_$w.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:${PAT[$w]} -s 5 -c 1 -o gpurun_out/prof_$w -f \
      python bench.py --workload $w --steps 6 --warmup 3 --no-cpu --e2e-steps 3 > gpurun_out/ncu_$w.log 2>&1
  tail -1 gpurun_out/ncu_$w.log | cut -c1-120
done
