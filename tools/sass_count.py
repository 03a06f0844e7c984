#!/usr/bin/env python3
"""Static SASS statistics of one kernel of librip_b200.so: instruction count and opcode histogram
(a proxy for executed instructions per pixel while iterating without a GPU).

    python tools/sass_count.py 'k_fused<31, 0>'        # substring of the demangled name
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "raw_image_pipeline_b200", "librip_b200.so")


def main():
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
