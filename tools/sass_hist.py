#!/usr/bin/env python
"""Opcode histogram of the hot loops of one kernel (cuobjdump -sass of an object file): the loops are the largest
backward branches; counts are per loop body and per row (body / rows_per_body).  ALU-pipe opcodes are summed, since that
pipe bounds the sweep kernels.   usage: sass_hist.py <object> <mangled-kernel-substring> [rows_per_body] [min_bytes]"""
import collections
import re
import subprocess
import sys

ALU = {"LOP3", "LEA", "SHF", "ISETP", "IADD3", "SEL", "VIADD", "PLOP3", "VIMNMX", "VIADDMNMX", "PRMT", "POPC", "FLO", "IABS", "BMSK", "SGXT", "LOP"}


def main(obj, sub, rows=3, min_bytes=0x2000):
    names = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    fn = None
    for m in re.finditer(r"Function : (\S+)", names):
        if sub in m.group(1):
            fn = m.group(1)
            break
    if not fn:
        sys.exit("kernel not found")
    txt = subprocess.run(["cuobjdump", "-sass", "-fun", fn, obj], capture_output=True, text=True).stdout
    ins = []
    for l in txt.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), re.sub(r"^@!?U?P\d+\s+", "", m.group(2).strip()), m.group(2).strip().startswith("@")))
    print(fn, len(ins), "instructions")
    loops = []
    for a, t, pred in ins:
        m = re.search(r"BRA\S*\s+(?:!?U?P\d+,\s*|UR\d+,\s*)?0x([0-9a-f]+)", t)
        if m and pred and int(m.group(1), 16) < a and min_bytes <= a - int(m.group(1), 16) < 0x8000:
            loops.append((int(m.group(1), 16), a))
    for lo, hi in loops:
        c = collections.Counter(t.split()[0].split(".")[0] for a, t, _ in ins if lo <= a <= hi)
        tot = sum(c.values())
        alu = sum(v for k, v in c.items() if k in ALU)
        print(f"loop {lo:#x}..{hi:#x}: {tot} instr = {tot / rows:.1f} per row, ALU pipe {alu} = {alu / rows:.1f} per row")
        print("   " + "  ".join(f"{k} {v / rows:.1f}" for k, v in c.most_common(14)))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], *(int(x, 0) for x in sys.argv[3:]))
