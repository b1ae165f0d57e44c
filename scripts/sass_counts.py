#!/usr/bin/env python
"""SASS evidence for profiles/: per kernel of csrc/libsurs.so the counts of the Blackwell mnemonics that matter
(UTCHMMA = tcgen05.mma incl. .2CTA, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (1-D TMA),
UTMALDG = tensor-map TMA, SYNCS = mbarrier ops, REDUX / CREDUX = warp reductions), registers from the ptxas logs.
    python scripts/sass_counts.py > profiles/r2_sass_counts.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "super-resolution-3d-human-shape-from-a-single-low-resolution-image_b200", "csrc")
WANT = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "SYNCS", "REDUX", "CREDUX", "HMMA", "FFMA", "DFMA", "LDG", "STG", "LDS", "STS", "ATOM", "RED"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", os.path.join(CSRC, "libsurs.so")], capture_output=True, text=True).stdout
    counts, fn = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            fn = re.sub(r"\(anonymous namespace\)::", "", fn).split("(")[0]
            counts[fn] = collections.Counter()
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and fn:
            op = m.group(1)
            counts[fn][op.split(".")[0]] += 1
            if op.startswith("UTCHMMA.2CTA"):
                counts[fn]["UTCHMMA.2CTA"] += 1
    print("SASS mnemonic counts per kernel of csrc/libsurs.so (cuobjdump -sass, sm_100a)")
    print("%-78s %s" % ("kernel", " ".join("%9s" % w for w in WANT)))
    for fn, c in counts.items():
        if sum(c.values()) == 0:
            continue
        print("%-78s %s" % (fn[:78], " ".join("%9d" % c.get(w, 0) for w in WANT)))
    print()
    print("registers / spills (ptxas -v, csrc/*.ptxas.log)")
    for log in sorted(os.listdir(CSRC)):
        if not log.endswith(".ptxas.log"):
            continue
        text = open(os.path.join(CSRC, log)).read()
        for m in re.finditer(r"Compiling entry function '(\S+)'.*?\n.*?\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", text):
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::", "", name).split("(")[0]
            print("%-78s regs %3s  stack %4s  spill st/ld %s/%s" % (name[:78], m.group(5), m.group(2), m.group(3), m.group(4)))


if __name__ == "__main__":
    main()
