#!/usr/bin/env python
"""Writes the generated part of a code-specialised LMS_DEC kernel (see ldpc-lib_b200/csrc/lms_spec.cuh): the base
matrix as compile-time tables plus the extern "C" kernel entry.  Used at build time for the frozen benchmark
matrices (ldpc-lib_b200/Makefile -> build/lms_spec_aot_gen.h); the library's run-time generator (spec_jit.cpp)
emits the same text for any other code.

    python tools/gen_lms_spec.py NAME=configs/file.jsonx:Z[:MINB[:lms|lmst|lmst2|ms|ims|mst|imst|imsh|imsh2]] ... > out.h
"""
import os
import re
import sys

import numpy as np


def load(path):
    txt = open(path).read()
    m = re.search(r"code\s*=\s*matrix\s*\((\d+)\s+(\d+)\)\s*\{(.*?)\}", txt, re.S)
    b, c = int(m.group(1)), int(m.group(2))
    return np.array(m.group(3).split(), dtype=np.int64).reshape(b, c)


def tmem_tables(b, c, Z, rp, col, sh, zp):
    """Tables of the tensor-memory kernel (lms_tmem.cuh): ROT[k] = shift of the last block row of column k (the
    column's rotation at iteration boundaries), DELTA[e] = (SH[e] - shift of the cyclically previous block row
    of the same column) mod Z, RI = (Z - ROT) mod Z, SYNSH[e] = (SH[e] - ROT[col]) mod Z, TCOLS = TMEM columns."""
    E = len(col)
    last = {}
    for e in range(E):
        last[col[e]] = sh[e]
    rot = [last.get(k, 0) for k in range(c)]
    cur = dict(last)
    delta = []
    for e in range(E):
        delta.append((sh[e] - cur[col[e]]) % Z)
        cur[col[e]] = sh[e]
    ri = [(Z - r) % Z for r in rot]
    synsh = [(sh[e] - rot[col[e]]) % Z for e in range(E)]
    need = E * ((zp // 32 + 3) // 4)
    tcols = 32
    while tcols < need:
        tcols *= 2
    lastw = [0] * E
    seen = set()
    for e in reversed(range(E)):
        if col[e] not in seen:
            lastw[e] = 1
            seen.add(col[e])
    return rot, delta, ri, synsh, tcols, lastw


def early_table(b, rp, col):
    """EARLY[e]: the column of edge e (block row j >= 1) is not touched by block row j - 1, so its value can be loaded one
    block row ahead (lms_tmem.cuh prefetch)."""
    early = []
    for j in range(b):
        prev = set(col[rp[j - 1]:rp[j]]) if j > 0 else None
        for e in range(rp[j], rp[j + 1]):
            early.append(1 if (j > 0 and col[e] not in prev) else 0)
    return early


def emit(name, hd, Z, minb, kind="lms"):
    b, c = hd.shape
    rp, col, sh = [0], [], []
    for j in range(b):
        for i in range(c):
            if hd[j, i] != -1:
                col.append(i); sh.append(int(hd[j, i]) % Z)
        rp.append(len(col))
    E = len(col)
    zp = (Z + 31) // 32 * 32
    arr = lambda v: "{ " + ", ".join(str(x) for x in v) + " }"
    out = []
    out.append("namespace ldpcb200 { namespace gen_%s {" % name)
    out.append("__constant__ int RT_RP[%d] = %s;" % (b + 1, arr(rp)))
    out.append("__constant__ int RT_COL[%d] = %s;" % (E, arr(col)))
    out.append("__constant__ int RT_SH[%d] = %s;" % (E, arr(sh)))
    rot, delta, ri, synsh, tcols, lastw = tmem_tables(b, c, Z, rp, col, sh, zp)
    out.append("__constant__ int RT_ROT[%d] = %s;" % (c, arr(rot)))
    out.append("__constant__ int RT_RI[%d] = %s;" % (c, arr(ri)))
    out.append("__constant__ int RT_SYNSH[%d] = %s;" % (E, arr(synsh)))
    out.append("struct Code {")
    maxdeg = max(rp[j + 1] - rp[j] for j in range(b))
    out.append("    static constexpr int B = %d, C = %d, Z = %d, E = %d, ZP = %d, MINB = %d, MAXDEG = %d;" % (b, c, Z, E, zp, minb, maxdeg))
    out.append("    static constexpr bool DOUBLED = true, PS_SMEM = false;")
    out.append("    static constexpr int RP[%d] = %s;" % (b + 1, arr(rp)))
    out.append("    static constexpr int COL[%d] = %s;" % (E, arr(col)))
    out.append("    static constexpr int SH[%d] = %s;" % (E, arr(sh)))
    seen, first = set(), []
    for e in range(E):
        first.append(0 if col[e] in seen else 1)
        seen.add(col[e])
    out.append("    static constexpr bool FIRST[%d] = %s;     // first edge of its block column (ascending block rows)" % (E, arr(first)))
    # lms_tmem.cuh: every block column is kept in the rotation of its last writer
    rot, delta, ri, synsh, tcols, lastw = tmem_tables(b, c, Z, rp, col, sh, zp)
    out.append("    static constexpr int TCOLS = %d;" % tcols)
    out.append("    static constexpr int DELTA[%d] = %s;" % (E, arr(delta)))
    out.append("    static constexpr int ROT[%d] = %s;" % (c, arr(rot)))
    out.append("    static constexpr int RI[%d] = %s;" % (c, arr(ri)))
    out.append("    static constexpr int SYNSH[%d] = %s;" % (E, arr(synsh)))
    out.append("    static constexpr bool LAST[%d] = %s;      // last edge of its block column (ascending block rows)" % (E, arr(lastw)))
    out.append("    static constexpr bool EARLY[%d] = %s;     // column untouched by the previous block row" % (E, arr(early_table(b, rp, col))))
    out.append("    static __device__ __forceinline__ const int* rt_rot() { return RT_ROT; }")
    out.append("    static __device__ __forceinline__ const int* rt_ri() { return RT_RI; }")
    out.append("    static __device__ __forceinline__ const int* rt_synsh() { return RT_SYNSH; }")
    out.append("    static __device__ __forceinline__ const int* rt_rp() { return RT_RP; }")
    out.append("    static __device__ __forceinline__ const int* rt_col() { return RT_COL; }")
    out.append("    static __device__ __forceinline__ const int* rt_sh() { return RT_SH; }")
    out.append("};")
    out.append("} }")
    if kind == "lmst":
        out.append('extern "C" __global__ void __launch_bounds__(%d, %d) lmst_spec_%s(const __grid_constant__ ldpcb200::FrameIO io)' % (zp, minb, name))
        out.append("{ ldpcb200::LmsTmem<ldpcb200::gen_%s::Code>::kernel(io); }" % name)
        out.append("LDPC_MS_SPEC_REGISTER(lmst, %s, %d, %d, %d, %d, %d, %d)" % (name, b, c, Z, E, zp, minb))
    elif kind == "lmst2":
        out.append('extern "C" __global__ void __launch_bounds__(%d, 1) lmst2_spec_%s(const __grid_constant__ ldpcb200::FrameIO io)' % (zp, name))
        out.append("{ ldpcb200::LmsTmem2<ldpcb200::gen_%s::Code>::kernel(io); }" % name)
        out.append("LDPC_MS_SPEC_REGISTER(lmst2, %s, %d, %d, %d, %d, %d, %d)" % (name, b, c, Z, E, zp, 1))
    elif kind in ("mst", "imst"):
        out.append('extern "C" __global__ void __launch_bounds__(%d, %d) %s_spec_%s(const __grid_constant__ ldpcb200::FrameIO io, const ldpcb200::MsSpecParams sp)' % (zp, minb, kind, name))
        out.append("{ ldpcb200::MsTmem<ldpcb200::gen_%s::Code, %s>::kernel(io, sp); }" % (name, "true" if kind == "imst" else "false"))
        out.append("LDPC_MS_SPEC_REGISTER(%s, %s, %d, %d, %d, %d, %d, %d)" % (kind, name, b, c, Z, E, zp, minb))
    elif kind in ("imsh", "imsh2"):                         # IMS_DEC as fp16 pairs: one or two groups of zp threads per CTA
        groups = 2 if kind == "imsh2" else 1
        out.append('extern "C" __global__ void __launch_bounds__(%d, %d) %s_spec_%s(const __grid_constant__ ldpcb200::FrameIO io, const ldpcb200::MsSpecParams sp)' % (groups * zp, minb, kind, name))
        out.append("{ ldpcb200::ImsH2<ldpcb200::gen_%s::Code, %d>::kernel(io, sp); }" % (name, groups))
        out.append("LDPC_MS_SPEC_REGISTER(%s, %s, %d, %d, %d, %d, %d, %d)" % (kind, name, b, c, Z, E, groups * zp, minb))
    elif kind == "lms":
        out.append('extern "C" __global__ void __launch_bounds__(%d, %d) lms_spec_%s(const __grid_constant__ ldpcb200::FrameIO io)' % (zp, minb, name))
        out.append("{ ldpcb200::LmsSpec<ldpcb200::gen_%s::Code>::kernel(io); }" % name)
        out.append("LDPC_SPEC_REGISTER(%s, %d, %d, %d, %d, %d, %d)" % (name, b, c, Z, E, zp, minb))
    else:
        out.append('extern "C" __global__ void __launch_bounds__(%d, %d) %s_spec_%s(const __grid_constant__ ldpcb200::FrameIO io, const ldpcb200::MsSpecParams sp)' % (zp, minb, kind, name))
        out.append("{ ldpcb200::MsSpec<ldpcb200::gen_%s::Code, %s>::kernel(io, sp); }" % (name, "true" if kind == "ims" else "false"))
        out.append("LDPC_MS_SPEC_REGISTER(%s, %s, %d, %d, %d, %d, %d, %d)" % (kind, name, b, c, Z, E, zp, minb))
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    print("// generated by tools/gen_lms_spec.py -- do not edit")
    for spec in sys.argv[1:]:
        name, rest = spec.split("=")
        parts = rest.split(":")
        hd = load(parts[0])
        Z = int(parts[1])
        minb = int(parts[2]) if len(parts) > 2 else 1
        kind = parts[3] if len(parts) > 3 else "lms"
        sys.stdout.write(emit(name, hd, Z, minb, kind))
