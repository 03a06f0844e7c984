#!/usr/bin/env python3
"""Per-SASS-instruction execution counts of one kernel from an ncu report (page `source`): prints the instructions
executed per pixel by address range, so that hot-loop overheads show up.

    python tools/ncu_src.py gpurun_out/x.ncu-rep <pixels> [lo_hex hi_hex]
"""
import csv, subprocess, sys, io, collections
rep, px = sys.argv[1], float(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
base = None
tot = 0
lines = []
for r in rows[2:]:
    if len(r) <= iex: continue
    a = int(r[ia], 16)
    if base is None: base = a
    n = int(r[iex]); tot += n
    lines.append((a - base, r[isrc].strip(), n, int(r[ismp])))
print(f"total warp instructions {tot}, per pixel {tot / px:.1f}")
lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 30
ops = collections.Counter()
for a, s, n, smp in lines:
    if lo <= a <= hi:
        ops[s.split()[0 if not s.startswith('@') else 1].split('.')[0]] += n
        print(f"{a:05x} {n / px * 1.0:8.3f} {smp:6d}  {s}")
print("per-pixel by opcode:", ", ".join(f"{k} {v / px:.2f}" for k, v in ops.most_common(40)))
