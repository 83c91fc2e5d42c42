# A/B of the cash-penalty kernel (config 5 shape, 262144 envs, D = 100): batches of four holdings in flight
run() { python bench.py --workload cashpenalty_step --steps 200 --warmup 20 --no-cpu --e2e-steps 3 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', 'kernel_ms', round(d['roofline']['kernel_ms'],5), 'frac', round(d['roofline']['frac'],4))"; }
for rep in 1 2; do
for pf in 3; do FINRL_B200_LIB=$PWD/variants/libpf$pf.so run "prefetch depth $pf"; done
run "prefetch depth 1"
done
