#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full --import-source on): key raw metrics, stall-reason shares,
opcode mix and the most-stalled instructions.  Usage: python profiles/ncu_summary.py file.ncu-rep"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum",
        "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum",
        "sm__cycles_elapsed.max", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum", "smsp__inst_executed_op_shfl.sum"]
print("kernel:", rows[2][hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print("%-70s %s %s" % (w, rows[1][i], [r[i] for r in rows[2:]]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
h = rows[hi[0]]
body = [r for r in rows[hi[0] + 1:(hi[1] - 1 if len(hi) > 1 else len(rows))] if len(r) >= len(h)]
col = {n: i for i, n in enumerate(h)}
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
tot = {s: sum(int(r[col[s]] or 0) for r in body) for s in stalls}
N = sum(int(r[col["# Samples"]] or 0) for r in body)
print("samples", N)
for s, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]:
    print("  %-26s %6d %.3f" % (s, v, v / max(N, 1)))
ops = collections.Counter()
for r in body:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[col["Source"]])
    if m:
        ops[m.group(2).split(".")[0]] += int(r[col["Instructions Executed"]] or 0)
T = sum(ops.values())
print("warp instructions", T)
for o, c in ops.most_common(18):
    print("  %-10s %10d %.3f" % (o, c, c / T))
for r in sorted(body, key=lambda r: -int(r[col["# Samples"]] or 0))[:14]:
    print(r[col["# Samples"]], r[col["Source"]].strip()[:70],
          {s[6:]: r[col[s]] for s in stalls if int(r[col[s]] or 0) > 15})
