#!/usr/bin/env python
"""In-kernel QAM channel: the factored demodulator (channel.cuh pam_demod_factored, the default) against the reference's
evaluation order (LDPCB200_QAM_EXACT=1) -- per-frame records of the same simulate() call and kernel time.
    python tools/qam_ab.py [CODE Z DECODER MAXITER SNR FRAMES MODULATION PUNCTURED_BLOCKS]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "ldpc-lib_b200")); sys.path.insert(0, ROOT)
import numpy as np                                            # noqa: E402
from codes import load_code                                   # noqa: E402
import pyldpcb200 as L                                        # noqa: E402

a = sys.argv[1:] or ["c3_bg1_46x68", "384", "LMS", "10", "1.5", "20000", "3", "2"]
code, Z, dec, maxiter, snr, nf, mod, punct = a[0], int(a[1]), a[2], int(a[3]), float(a[4]), int(a[5]), int(a[6]), int(a[7])
hd, _ = load_code(code)
K = (hd.shape[1] - hd.shape[0]) * Z
out = {"case": a, "rows": []}
recs = {}
for exact in ("1", "0"):
    os.environ["LDPCB200_QAM_EXACT"] = exact
    with L.Decoder(hd, Z, getattr(L, dec + "_DEC"), precision=32 if dec in ("LMS", "MS") else 64, use_fast=2) as d:
        d.simulate(snr, 2000, maxiter, modulation=mod, punct=punct, seed=1)
        r = d.simulate(snr, nf, maxiter, modulation=mod, punct=punct, seed=1, stream=1, want_per_frame=True)
        ms, _ = d.last_kernel_ms()
    recs[exact] = r["per_frame"]
    out["rows"].append({"demodulator": "reference order (pam_demod)" if exact == "1" else "factored (pam_demod_factored)", "kernel_ms": ms,
                        "info_gbps": nf * K / ms / 1e6, "frames": r["frames"], "frame_errors": r["frame_errors"], "bit_errors": r["bit_errors"],
                        "iter_sum": r["iter_sum"]})
out["frames_with_a_different_record"] = int((recs["0"] != recs["1"]).sum())
print(json.dumps(out, indent=1))
