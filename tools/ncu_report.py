#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): headline metrics, stall reasons, and executed
instructions / stall samples per CUDA source line (SASS joined to -lineinfo via nvdisasm).

    python tools/ncu_report.py gpurun_out/prof.ncu-rep [kernel-substring] [--units N] [--top 40]
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_not_selected_per_warp_active.pct", "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct",
    "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct", "smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct",
    "smsp__warp_issue_stalled_drain_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    args = [a for a in sys.argv[2:] if not a.startswith("--")]
    pat = args[0] if args else ""
    units = float(sys.argv[sys.argv.index("--units") + 1]) if "--units" in sys.argv else None
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    hdr, unit = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    sel = [r for r in rows[2:] if pat in r[ik]]
    print(f"# {rep}: {len(sel)} launch(es) matching '{pat}'")
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w:75s} {unit[i]:12s} {[r[i] for r in sel]}")
    # ---- per source line ----
    srows = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv"))))
    blocks, cur, name, shdr = [], None, None, None
    for r in srows:
        if r and r[0] == "Kernel Name":
            name = r[1]
            cur = []
            blocks.append((name, cur))
        elif r and r[0] == "Address":
            shdr = r
        elif cur is not None and shdr is not None and len(r) == len(shdr):
            cur.append(r)
    blk = next(((n, b) for n, b in blocks if pat in n), None)
    if blk is None:
        return
    name, b = blk
    ia, isrc, ii, isamp = shdr.index("Address"), shdr.index("Source"), shdr.index("Instructions Executed"), shdr.index("# Samples")
    lib = os.path.join(ROOT, "finrl_b200", "libfinrl_b200.so")
    tmp = "/tmp/ncu_report_cubin"
    os.makedirs(tmp, exist_ok=True)
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
    amap = {}
    m = re.search(r"(\w+)<", name)
    short = (m.group(1) if m else name).split("::")[-1]
    targs = re.findall(r"\(int\)(\d+)|, (float|double),", name)
    for f in os.listdir(tmp):
        if not f.endswith(".cubin"):
            continue
        out = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        func, line, fmap = None, None, {}
        for l in out.split("\n"):
            mm = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
            if mm:
                func = mm.group(1)
                fmap[func] = {}
                continue
            mm = re.search(r'//## File "([^"]+)", line (\d+)', l)
            if mm:
                line = (mm.group(1).split("/")[-1], int(mm.group(2)))
            mm = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
            if mm and func:
                fmap[func][int(mm.group(1), 16)] = line
        for fn, mp in fmap.items():
            if short in fn and len(mp) == len(b):
                amap = mp
    base = int(b[0][ia], 16)
    byline = collections.defaultdict(lambda: [0, 0])
    ops = collections.Counter()
    for r in b:
        ln = amap.get(int(r[ia], 16) - base)
        byline[ln][0] += int(r[ii])
        byline[ln][1] += int(r[isamp])
        toks = r[isrc].split()
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
        ops[op] += int(r[ii])
    ti = sum(v[0] for v in byline.values())
    ts = sum(v[1] for v in byline.values()) or 1
    print(f"\n# {name[:100]}\n# executed warp-instructions {ti}" + (f" = {ti / units:.1f} per unit" if units else ""))
    print("# opcode mix: " + ", ".join(f"{o} {100 * c / ti:.1f}%" for o, c in ops.most_common(14)))
    src = {}
    for ln, v in sorted(byline.items(), key=lambda kv: -kv[1][0])[:top]:
        text = ""
        if ln:
            path = os.path.join(ROOT, "finrl_b200", "csrc", ln[0])
            if ln[0] not in src and os.path.exists(path):
                src[ln[0]] = open(path).read().split("\n")
            if ln[0] in src:
                text = src[ln[0]][ln[1] - 1].strip()[:80]
        per = f"{v[0] / units:6.1f}/u" if units else ""
        print(f"{str(ln):30s} inst {100 * v[0] / ti:5.1f}% {per} samp {100 * v[1] / ts:5.1f}%  {text}")


if __name__ == "__main__":
    main()
