#!/usr/bin/env python
"""SASS opcode histogram per kernel of libmopa_scn.so (cuobjdump -sass; runs without a GPU): the mnemonics that prove
what each kernel is made of (UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, UBLKCP = TMA bulk
copy, LDGSTS = cp.async, SYNCS = mbarrier, HMMA = mma.sync) + instruction count, registers, spill bytes.
    python tools/sass_histogram.py > profiles/r02_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mopa_b200", "libmopa_scn.so")
KEY = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UBLKCP", "UTMALDG", "LDGSTS", "ARRIVES", "SYNCS", "HMMA",
       "LDG", "STG", "LDS", "STS", "ATOM", "ATOMS", "ATOMG", "RED", "SHFL", "VOTE", "R2UR", "MEMBAR", "FENCE", "BAR", "STL", "LDL"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    usage = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    regs = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+)", usage):
        regs[m.group(1)] = (int(m.group(2)), int(m.group(3)))
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
            cur["_total"] += 1
    print("# tools/sass_histogram.py on mopa_b200/libmopa_scn.so (sm_100a): instructions, registers / stack bytes, key opcodes")
    for name, c in kernels.items():
        short = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        short = re.sub(r"\(.*", "", short)
        r = regs.get(name, (None, None))
        ops = "  ".join("%s %d" % (k, c[k]) for k in KEY if c[k])
        print("%-58s %6d instr  regs %-4s stack %-4s | %s" % (short[:58], c["_total"], r[0], r[1], ops))


if __name__ == "__main__":
    main()
