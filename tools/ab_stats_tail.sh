# A/B of the statistics exchange on the 1M-env StockTradingEnv step: FRL_STATS_EXCHANGE=local (n_peers = 0, the
# kernels only accumulate) vs p2p (every launch pushes its predecessor's sums to the peers' totals)
for rep in 1 2 3; do
for m in local p2p; do
FRL_STATS_EXCHANGE=$m python bench.py --steps 300 --warmup 20 --no-extra --no-cpu --e2e-steps 3 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$m', 'kernel_ms', round(d['roofline']['kernel_ms'],5), 'ms_per_step', round(d['ms_per_step'],5), 'frac', round(d['roofline']['frac'],4))"
done; done
