#!/usr/bin/env python
"""One simulate() call of a named configuration -- the fixed small workload that ncu captures are taken of.
    python tools/profile_case.py CODE Z DECODER MAXITER SNR FRAMES [PRECISION [MODULATION [PUNCTURED_BLOCKS]]]
    e.g.  ref32x16_a 126 IMS 15 3.0 100000      c3_bg1_46x68 384 LMS 10 1.5 20000 32 3 2"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "ldpc-lib_b200")); sys.path.insert(0, ROOT)
from codes import load_code                                   # noqa: E402
import pyldpcb200 as L                                        # noqa: E402

code, Z, dec, maxiter, snr, nf = sys.argv[1], int(sys.argv[2]), sys.argv[3], int(sys.argv[4]), float(sys.argv[5]), int(sys.argv[6])
prec = int(sys.argv[7]) if len(sys.argv) > 7 else (32 if dec in ("LMS", "MS") else 64)
mod = int(sys.argv[8]) if len(sys.argv) > 8 else 0
punct = int(sys.argv[9]) if len(sys.argv) > 9 else 0
hd, _ = load_code(code)
with L.Decoder(hd, Z, getattr(L, dec + "_DEC"), precision=prec, use_fast=2) as d:
    print(d.kernel_info())
    d.simulate(snr, 2000, maxiter, modulation=mod, punct=punct, seed=1)
    r = d.simulate(snr, nf, maxiter, modulation=mod, punct=punct, seed=1, stream=1)
    ms, nl = d.last_kernel_ms()
    K = (hd.shape[1] - hd.shape[0]) * Z
    print("frames", r["frames"], "avg iters", r["iter_sum"] / r["frames"], "kernel ms", ms, "launches", nl, "info Gbit/s", nf * K / ms / 1e6)
