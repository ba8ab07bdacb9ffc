#!/usr/bin/env python
"""Instruction mix of ONE iteration of a code-specialised kernel from its SASS, grouped by issue pipe.

The iteration is the straight-line code from the first LDTM (block row 0's old messages) to the BLOCKROWS-th BAR.SYNC
after it (every block row ends in exactly one).  Blocks that only warp 0 executes (the repeat stores of lms_tmem's PP
layout: a forward branch around a loop that contains STS) are counted separately; the per-edge figures are the
average over the WARPS warps of a CTA.  bench.py's roofline block and DESIGN.md quote this tool's JSON output
(profiles/*_sass_mix.json); tests/test_abi.py fails when the built object no longer matches the committed file.

    cuobjdump -sass -fun lmst_spec_c2t ldpc-lib_b200/build/lms_spec_aot.o | python tools/sass_mix.py EDGES [BLOCKROWS [WARPS]] [--json]
"""
import json
import re
import sys

ALU = {"FMNMX", "FMNMX3", "LOP3", "SEL", "FSEL", "FSETP", "ISETP", "VIADD", "IADD3", "SHF", "PRMT", "PLOP3", "VOTE", "LEA", "FSET", "VIMNMX",
       "VIMNMX3", "IABS", "POPC", "HMNMX2", "VIADDMNMX", "LOP", "IADD", "MOV", "R2P", "P2R", "CS2R", "BMSK", "SGXT", "FLO", "BREV"}
FMA = {"FADD", "FMUL", "FFMA", "IMAD", "FADD2", "FMUL2", "FFMA2", "HADD2", "HMUL2", "HFMA2", "IMAD.MOV"}
LSU = {"LDS", "STS", "LDG", "STG", "ATOMS", "LDSM", "SHFL", "LDL", "STL", "LDC"}
TMEM = {"LDTM", "STTM"}


def parse(text):
    """-> list of (address, opcode, predicate or None, branch target or None)"""
    out = []
    for line in text.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)\s*(.*?);", line)
        if not m:
            continue
        addr, pred, op, rest = int(m.group(1), 16), (m.group(2) or "").strip() or None, m.group(3), m.group(5)
        tgt = None
        if op == "BRA":
            t = re.search(r"0x([0-9a-f]+)", rest)
            tgt = int(t.group(1), 16) if t else None
            u = re.match(r"(!?UP\d+)\s*,", rest)                 # BRA.U UP0, target: a branch on a uniform predicate
            if u and not pred:
                pred = u.group(1)
        out.append((addr, op, pred, tgt))
    return out


def mix(text, edges, blockrows=None, warps=8):
    ins = parse(text)
    first = next(i for i, x in enumerate(ins) if x[1] == "LDTM")
    addr_all = {x[0]: i for i, x in enumerate(ins)}
    # blocks behind a predicated forward branch: a loop with stores = warp 0's repeat stores (weight 1 / warps); a block with a
    # barrier = the full syndrome pass after a clean quick look (once or twice per frame: weight 0)
    weight = [1.0] * len(ins)
    for i, (addr, op, pred, tgt) in enumerate(ins):
        if op == "BRA" and pred and tgt in addr_all and i < addr_all[tgt] <= i + 600:
            j = addr_all[tgt]
            blk = ins[i + 1:j]
            if any(x[1] == "STS" for x in blk) and any(x[1] == "BRA" and x[3] is not None and x[3] <= x[0] for x in blk):
                for k in range(i + 1, j):
                    weight[k] = min(weight[k], 1.0 / warps)
            elif any(x[1] == "BAR" for x in blk):
                for k in range(i + 1, j):
                    weight[k] = 0.0
    if blockrows:
        # the iteration = the window [an LDTM, the BLOCKROWS-th unconditional BAR after it] that holds the most three-input minima
        # (frame load and syndrome code also contain LDTM / BAR)
        def window(start):
            bars = 0
            for i in range(start, len(ins)):
                if ins[i][1] == "BAR" and weight[i] == 1.0:
                    bars += 1
                    if bars == blockrows:
                        return i
            return None
        best = None
        for i, x in enumerate(ins):
            if x[1] != "LDTM" or (i > 0 and ins[i - 1][1] == "LDTM"):
                continue
            e = window(i)
            if e is None:
                continue
            score = sum(1 for k in range(i, e + 1) if ins[k][1] in ("FMNMX3", "FMNMX", "HMNMX2", "VIMNMX") and weight[k] == 1.0)
            if best is None or score > best[0]:
                best = (score, i, e)
        if best is None:
            raise SystemExit("fewer than %d BAR after any LDTM" % blockrows)
        first, end = best[1], best[2]
    else:
        last_sttm = max(i for i, x in enumerate(ins) if x[1] == "STTM")
        end = next(i for i in range(last_sttm, len(ins)) if ins[i][1] == "BAR")
    body = ins[first:end + 1]
    wbody = weight[first:end + 1]
    warp0_only = [w < 1.0 and w > 0.0 for w in wbody]
    body = [x for x, w in zip(body, wbody) if w > 0.0]
    warp0_only = [f for f, w in zip(warp0_only, wbody) if w > 0.0]

    def group(op):
        return "alu" if op in ALU else "fma" if op in FMA else "lsu" if op in LSU else "tmem" if op in TMEM else "other"

    res = {"edges": edges, "blockrows": blockrows, "warps": warps}
    cnt_all, cnt_w0 = {}, {}
    for flag, (addr, op, pred, tgt) in zip(warp0_only, body):
        d = cnt_w0 if flag else cnt_all
        d[op] = d.get(op, 0) + 1
    grp = {g: 0.0 for g in ("alu", "fma", "lsu", "tmem", "other")}
    for op, n in cnt_all.items():
        grp[group(op)] += n
    for op, n in cnt_w0.items():
        grp[group(op)] += n / warps
    total = sum(grp.values())
    res["instructions_per_iteration_every_warp"] = sum(cnt_all.values())
    res["instructions_per_iteration_warp0_only"] = sum(cnt_w0.values())
    res["per_edge"] = {"total": total / edges, **{g: n / edges for g, n in grp.items()}}
    # shared-memory wavefronts: one per LDS / STS of a full warp (a block only warp 0 runs: 1 / warps)
    wf = sum(n for op, n in cnt_all.items() if op in ("LDS", "STS", "ATOMS")) + sum(n for op, n in cnt_w0.items() if op in ("LDS", "STS")) / warps
    res["per_edge"]["shared_memory_wavefronts"] = wf / edges
    res["opcodes"] = dict(sorted(cnt_all.items(), key=lambda kv: -kv[1]))
    res["opcodes_warp0_only"] = dict(sorted(cnt_w0.items(), key=lambda kv: -kv[1]))
    return res


def main():
    argv = [a for a in sys.argv[1:] if a != "--json"]
    edges = int(argv[0])
    blockrows = int(argv[1]) if len(argv) > 1 else None
    warps = int(argv[2]) if len(argv) > 2 else 8
    res = mix(sys.stdin.read(), edges, blockrows, warps)
    if "--json" in sys.argv:
        print(json.dumps(res, indent=1))
        return
    pe = res["per_edge"]
    print("instructions per iteration and lane: %d (+ %d in warp 0 only)  (%.2f per edge)" % (res["instructions_per_iteration_every_warp"],
                                                                                       res["instructions_per_iteration_warp0_only"], pe["total"]))
    for g in ("alu", "fma", "lsu", "tmem", "other"):
        print("  %-5s %.2f per edge" % (g, pe[g]))
    print("  shared-memory wavefronts %.2f per edge" % pe["shared_memory_wavefronts"])
    print("  " + ", ".join("%s %d" % kv for kv in res["opcodes"].items()))
    if res["opcodes_warp0_only"]:
        print("  warp 0 only: " + ", ".join("%s %d" % kv for kv in res["opcodes_warp0_only"].items()))


if __name__ == "__main__":
    main()
