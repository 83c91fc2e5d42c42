#!/usr/bin/env python
"""Top SASS instructions by one stall reason, with the CUDA source line (from -lineinfo via ncu's source page).

    python tools/ncu_stall_lines.py gpurun_out/prof_np_step.ncu-rep stall_long_sb [top]
"""
import csv, io, subprocess, sys
rep, col = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hdr]
ic, isrc = h.index(col), h.index("Source")
tot = 0
items = []
for n, r in enumerate(rows[hdr + 1:]):
    if len(r) <= ic:
        continue
    try:
        v = int(r[ic])
    except ValueError:
        continue
    tot += v
    items.append((v, n, r[isrc].strip()))
print(f"# {col}: {tot} samples")
for v, n, src in sorted(items, reverse=True)[:top]:
    # show the instruction and the previous few instructions' producers roughly by index
    print(f"{100.0 * v / max(tot, 1):5.1f}%  #{n:5d}  {src[:110]}")
