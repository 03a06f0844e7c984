#!/usr/bin/env python3
"""Static SASS view of one kernel of librip_b200.so: dumps the kernel, prints the opcode histogram of the whole kernel and
of its longest backward-branch loop body (a proxy for executed instructions per iteration while iterating without a GPU).

    python tools/sass_loop.py k_fused_stripILj31ELb1ELi4 [out.sass]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "raw_image_pipeline_b200", "librip_b200.so")


def main():
    pat = sys.argv[1]
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    for blk in re.split(r"\n\s*Function : ", sass)[1:]:
        name = blk.split("\n", 1)[0].strip()
        if pat not in name:
            continue
        ins = []
        for line in blk.splitlines():
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)(.*?);", line)
            if m:
                ins.append((int(m.group(1), 16), m.group(3), (m.group(2) or "") + m.group(3) + m.group(4)))
        if len(sys.argv) > 2:
            with open(sys.argv[2], "w") as f:
                for a, _, t in ins:
                    f.write(f"{a:05x}  {t}\n")
        # longest loop: backward branch with the largest span
        best = None
        for a, op, t in ins:
            if op.startswith("BRA"):
                m = re.search(r"0x([0-9a-f]+)\s*$", t.strip())
                if m:
                    tgt = int(m.group(1), 16)
                    if tgt < a and (best is None or a - tgt > best[1] - best[0]):
                        best = (tgt, a)
        print(f"{name[-70:]}: {len(ins)} instructions")
        def hist(sel):
            c = collections.Counter(op.split(".")[0] for a, op, t in ins if sel(a))
            return sum(c.values()), ", ".join(f"{k} {v}" for k, v in c.most_common(40))
        n, h = hist(lambda a: True)
        print("  all:", h)
        if best:
            # second-longest (inner) loops too
            loops = []
            for a, op, t in ins:
                if op.startswith("BRA"):
                    m = re.search(r"0x([0-9a-f]+)\s*$", t.strip())
                    if m and int(m.group(1), 16) < a:
                        loops.append((a - int(m.group(1), 16), int(m.group(1), 16), a))
            for span, lo, hi in sorted(loops, reverse=True)[:3]:
                n, h = hist(lambda a: lo <= a <= hi)
                print(f"  loop {lo:#x}..{hi:#x}: {n} instructions\n    {h}")


if __name__ == "__main__":
    main()
