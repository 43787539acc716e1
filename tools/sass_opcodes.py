#!/usr/bin/env python3
"""Per-kernel SASS opcode histogram of libadtfe.so (cuobjdump -sass): the auditable form of "hand-written sm_100a".

    python tools/sass_opcodes.py [adt_str_b200/libadtfe.so] > profiles/r02_sass_opcodes.txt

Blackwell-specific mnemonics to look for: UBLKCP (cp.async.bulk), SYNCS.* (mbarrier), FFMA2 / FADD2 / FMUL2 (packed
fp32), REDUX / CREDUX, FMNMX3, UTCHMMA / UTCBAR / LDTM (tcgen05 MMA, commit, tensor-memory load), UTMALDG (TMA tensor load).
"""
import collections
import re
import subprocess
import sys


def main(path):
    text = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in text.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    arch = re.search(r"arch = (sm_\w+)", text)
    print(f"{path}: {len(kernels)} kernels, {arch.group(1) if arch else '?'}")
    for name, ops in kernels.items():
        demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0]
        total = sum(ops.values())
        print(f"\n== {demangled}  [{total} SASS instructions]")
        base = collections.Counter()
        for op, n in ops.items():
            base[op.split(".")[0]] += n
        print("   " + ", ".join(f"{op} {n}" for op, n in base.most_common(18)))
        special = {op: n for op, n in ops.items() if re.match(r"(UBLKCP|UTMA|SYNCS|UTC|LDTM|STTM|REDUX|CREDUX|FMNMX3|F(FMA|ADD|MUL)2|ELECT|UCGABAR|FENCE)", op)}
        if special:
            print("   sm_100a / async: " + ", ".join(f"{op} {n}" for op, n in sorted(special.items())))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "adt_str_b200/libadtfe.so")
