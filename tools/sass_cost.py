#!/usr/bin/env python3
"""Static issue-cost estimate of a kernel's SASS (not product code): sums, over an address range of the
disassembly, a per-opcode cost in SM-clocks per warp instruction taken from tools/microbench/int_throughput.cu
measured on B200 (profiles/r01_int_throughput.txt).  Usage:
    cuobjdump -sass obj.o | python tools/sass_cost.py <kernel-substring> [lo_hex hi_hex]"""
import re
import sys

COST = {"IADD3": 0.30, "IADD3.X": 0.62, "IMAD.X": 0.62, "IMAD.WIDE": 0.90, "IMAD.WIDE.U32": 0.90, "IMAD.WIDE.U32.X": 1.0, "IMAD.HI.U32": 1.0, "IMAD.HI": 1.0,
        "IMAD": 0.5, "IMAD.U32": 0.5, "IMAD.MOV": 0.5, "IMAD.MOV.U32": 0.5, "IMAD.IADD": 0.5, "IMAD.SHL": 0.5, "IMAD.SHL.U32": 0.5, "LOP3.LUT": 0.5, "SHF": 0.5, "PRMT": 0.5,
        "SEL": 0.5, "MOV": 0.3, "VIADD": 0.3, "ISETP": 0.5, "LEA": 0.5, "LEA.HI": 0.5, "LEA.HI.X": 0.62}


def main():
    name = sys.argv[1]
    lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
    hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 30
    inside = False
    counts, total, n = {}, 0.0, 0
    for line in sys.stdin:
        if "Function :" in line:
            inside = name in line
            continue
        if not inside:
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        addr = int(m.group(1), 16)
        if addr < lo or addr >= hi:
            continue
        op = m.group(3)
        key = op
        while key and key not in COST:
            key = key.rsplit(".", 1)[0] if "." in key else ""
        c = COST.get(key, 0.5)
        counts[op] = counts.get(op, 0) + 1
        total += c
        n += 1
    print("%d instructions, estimated %.1f SM-clk per warp (%.2f per instruction)" % (n, total, total / max(n, 1)))
    for op, k in sorted(counts.items(), key=lambda kv: -kv[1])[:14]:
        print("  %5d %s" % (k, op))


if __name__ == "__main__":
    main()
