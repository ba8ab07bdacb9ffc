#!/usr/bin/env python
"""Instruction mix of one iteration of a code-specialised kernel from its SASS: everything between the first LDTM of
block row 0 and the last barrier before the per-iteration syndrome, grouped by issue pipe.  The per-edge figures that
bench.py's `issue` block and DESIGN.md §4.1 quote come from here.
    cuobjdump -sass -fun lmst_spec_c2t ldpc-lib_b200/build/lms_spec_aot.o | python tools/sass_mix.py EDGES"""
import re
import sys

ALU = {"FMNMX", "FMNMX3", "LOP3", "SEL", "FSEL", "FSETP", "ISETP", "VIADD", "IADD3", "SHF", "PRMT", "PLOP3", "VOTE", "LEA", "FSET", "VIMNMX", "IABS", "POPC"}
FMA = {"FADD", "FMUL", "FFMA", "IMAD", "FADD2", "FMUL2", "FFMA2"}
LSU = {"LDS", "STS", "LDG", "STG", "ATOMS", "LDSM"}
TMEM = {"LDTM", "STTM"}


def main():
    edges = int(sys.argv[1])
    ops = []
    for line in sys.stdin:
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m:
            ops.append(m.group(2))
    first = next(i for i, o in enumerate(ops) if o == "LDTM")
    last_sttm = max(i for i, o in enumerate(ops) if o == "STTM")
    end = next(i for i in range(last_sttm, len(ops)) if ops[i] == "BAR")
    body = ops[first:end + 1]
    cnt = {}
    for o in body:
        cnt[o] = cnt.get(o, 0) + 1
    grp = {"alu": 0, "fma": 0, "lsu": 0, "tmem": 0, "other": 0}
    for o, n in cnt.items():
        g = "alu" if o in ALU else "fma" if o in FMA else "lsu" if o in LSU else "tmem" if o in TMEM else "other"
        grp[g] += n
    print("instructions per iteration and lane: %d (%.2f per edge)" % (len(body), len(body) / edges))
    for g, n in grp.items():
        print("  %-5s %5d  %.2f per edge" % (g, n, n / edges))
    print("  " + ", ".join("%s %d" % kv for kv in sorted(cnt.items(), key=lambda kv: -kv[1])))


if __name__ == "__main__":
    main()
