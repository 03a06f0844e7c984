#!/usr/bin/env python3
"""Executed-instruction mix of one kernel from an .ncu-rep (source page): per-opcode warp instructions
per 32 pixels (= thread instructions per pixel).  usage: ncu_opmix.py rep kernel_regex n_pixels [top]"""
import collections, csv, io, re, subprocess, sys
rep, kern, npx = sys.argv[1], sys.argv[2], float(sys.argv[3])
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
iS, iI, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
iW = hdr.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in hdr else None; iWi = hdr.index("L1 Wavefronts Shared Ideal") if iW is not None else None
ops, samp = collections.Counter(), collections.Counter(); tot = 0; wf = wfi = 0
lines = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    try: n = int(r[iI])
    except ValueError: continue
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[iS].strip())
    op = m.group(2) if m else r[iS][:8]
    ops[op] += n; tot += n; samp[op] += int(r[iSm] or 0)
    if iW is not None:
        wf += int(r[iW] or 0); wfi += int(r[iWi] or 0)
    lines.append((n, int(r[iSm] or 0), r[iS].strip()))
w = npx / 32
print(f"total {tot/w:.1f} instr/px; shared wavefronts {wf/w:.1f} per 32 px (ideal {wfi/w:.1f})")
print(", ".join(f"{k} {v/w:.1f}" for k, v in ops.most_common(top)))
if "--lines" in sys.argv:
    for n, s, t in lines: print(f"{n/w:7.2f} {s:6d}  {t}")
