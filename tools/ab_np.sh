# A/B of the numpy-env kernel (config 3 shape, 1M envs, D = 30): previous commit vs working tree
run() { python bench.py --workload np_step --steps 200 --warmup 20 --no-cpu --e2e-steps 3 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', 'kernel_ms', round(d['roofline']['kernel_ms'],5), 'frac', round(d['roofline']['frac'],4))"; }
for rep in 1 2; do
FINRL_B200_LIB=$PWD/variants/libhead.so run "HEAD        "
run "working tree"
done
