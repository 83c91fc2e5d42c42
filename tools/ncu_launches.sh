#!/bin/bash
# launch list of the default bench command (plain run first, then the same command under ncu with only the
# duration metric), plus the per-kernel share of device time
set -u
CMD="python bench.py --steps 30 --warmup 3 --no-cpu --no-extra --e2e-steps 2"
$CMD > gpurun_out/plain_launches.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/launches.csv")) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
tot = collections.defaultdict(lambda: [0.0, 0])
for r in rows[hdr + 1:]:
    try:
        v = float(r[mv].replace(",", ""))
    except ValueError:
        continue
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[mu], 1e-6)
    tot[r[kn]][0] += v
    tot[r[kn]][1] += 1
s = sum(v[0] for v in tot.values())
with open("gpurun_out/launch_shares.txt", "w") as f:
    f.write("# kernel share of the bench command (ncu --metrics gpu__time_duration.sum, cold-cache serialised launches)\n")
    for k, (ms, n) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
        f.write("%10.3f ms %4dx %5.1f%%  %s\n" % (ms, n, 100 * ms / s, k[:100]))
print(open("gpurun_out/launch_shares.txt").read())
PY
