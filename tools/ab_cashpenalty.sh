# A/B of the cash-penalty kernel variants (config 5 shape, 262144 envs, D = 100)
run() { python bench.py --workload cashpenalty_step --steps 200 --warmup 20 --no-cpu --e2e-steps 3 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', 'kernel_ms', round(d['roofline']['kernel_ms'],5), 'frac', round(d['roofline']['frac'],4))"; }
for rep in 1 2; do
FINRL_B200_LIB=$PWD/variants/libnobulk.so run "cp.async staging, division routine        "
run "bulk (TMA) staging, reciprocal table, 128-bit"
done
