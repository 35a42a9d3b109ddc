#!/usr/bin/env python
"""Static SASS opcode histogram of selected kernels of csrc/librz_engine.so (cuobjdump -sass; runs without a GPU).

    python tools/sass_counts.py k1_bs_staged k1_find_correct_moves k1_calc_flip > profiles/k1_r02_sass_counts.txt

The bit-sliced K1 kernels are straight-line code per tile (one warp = 32 lanes x 32 positions = 1024 positions per tile),
so the static count of the tile body divided by 32 is the dynamic instruction count per position and lane."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "reversi-alpha-zero_b200", "csrc", "librz_engine.so")


def functions():
    text = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    name, ops = None, None
    for line in text.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name:
                yield name, ops
            name, ops = m.group(1), []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and name:
            ops.append(m.group(1))
    if name:
        yield name, ops


def demangle(n):
    try:
        return subprocess.run(["cu++filt", n], stdout=subprocess.PIPE, text=True).stdout.strip() or n
    except Exception:
        return n


def main(patterns):
    for name, ops in functions():
        if not any(p in name for p in patterns):
            continue
        base = collections.Counter(o.split(".")[0] for o in ops)
        print("== %s" % demangle(name))
        print("   instructions: %d" % len(ops))
        print("   " + "  ".join("%s %d" % kv for kv in base.most_common(18)))
        mem = {k: v for k, v in collections.Counter(ops).items() if k.split(".")[0] in ("LDG", "STG", "LDS", "STS", "UBLKCP", "SYNCS", "LDGSTS", "UTCHMMA", "LDTM", "STTM")}
        if mem:
            print("   memory / async: " + "  ".join("%s %d" % kv for kv in sorted(mem.items())))


if __name__ == "__main__":
    main(sys.argv[1:] or ["k1_"])
