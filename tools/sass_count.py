#!/usr/bin/env python3
"""Static SASS statistics of one kernel of librip_b200.so: instruction count and opcode histogram
(a proxy for executed instructions per pixel while iterating without a GPU).

    python tools/sass_count.py 'k_fused<31, 0>'        # substring of the demangled name
    python tools/sass_count.py --all > profiles/sass_opcodes.txt   # every kernel of the library: size, the TMA / mbarrier /
                                                                   # bulk-group opcodes that prove the sm_100a data path, top opcodes
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "raw_image_pipeline_b200", "librip_b200.so")


PROOF = ("UTMALDG", "UTMASTG", "SYNCS", "UTMACMDFLUSH", "FENCE", "ELECT", "LDS", "STS", "LDG", "STG", "ATOMG", "ATOMS", "RED", "IDP", "PRMT")


def all_kernels(lib):
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    rows = []
    for blk in re.split(r"\n\s*Function : ", sass)[1:]:
        name = blk.split("\n", 1)[0].strip()
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        dem = re.sub(r"\(anonymous namespace\)::", "", dem.replace("rip::", "")).split("(")[0]
        dem = dem.replace("void ", "")
        ops = collections.Counter()
        for line in blk.splitlines():
            m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
            if m:
                ops[m.group(2)] += 1
        rows.append((dem, sum(ops.values()), ops))
    rows.sort()
    print(f"# static SASS opcode counts per kernel of {os.path.relpath(lib, ROOT)} (cuobjdump -sass, sm_100a); tools/sass_count.py --all")
    print("# UTMALDG / UTMASTG = TMA tensor load / store, SYNCS = mbarrier operations, UTMACMDFLUSH = bulk-group commit")
    for dem, n, ops in rows:
        proof = " ".join(f"{k}={ops[k]}" for k in PROOF if ops[k])
        top = ", ".join(f"{k} {v}" for k, v in ops.most_common(10))
        print(f"{dem}\n    {n} instructions | {proof}\n    top: {top}")


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--all":
        all_kernels(sys.argv[2] if len(sys.argv) > 2 else LIB)
        return
    pat = sys.argv[1] if len(sys.argv) > 1 else "k_fused<31"
    lib = sys.argv[2] if len(sys.argv) > 2 else LIB
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    blocks = re.split(r"\n\s*Function : ", sass)
    for blk in blocks[1:]:
        name = blk.split("\n", 1)[0].strip()
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        dem = dem.replace("(unsigned int)", "").replace("(int)", "").replace("rip::", "")
        if pat not in dem:
            continue
        ops = collections.Counter()
        n = 0
        for line in blk.splitlines():
            m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
            if m:
                ops[m.group(2)] += 1
                n += 1
        print(f"{dem}: {n} SASS instructions")
        print("  " + ", ".join(f"{k} {v}" for k, v in ops.most_common(28)))


if __name__ == "__main__":
    main()
