#!/bin/bash
# crossover between the thread-per-env ("tile") and the 8-lanes-per-env ("small") StockTradingEnv kernels
for n in 1024 2048 4096 8192 16384 32768; do
  for k in tile small; do
    FRL_TRADING_KERNEL=$k python bench.py --workload trading_rollout --envs $n --steps 200 --warmup 10 --no-cpu --e2e-steps 3 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=$n', '$k', 'ms=%.4f'%d['roofline']['kernel_ms'], 'env-steps/s=%.3e'%d['value'])"
  done
done
