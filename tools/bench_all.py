#!/usr/bin/env python
"""Throughput + parity of every decoder on the benchmark configurations (BASELINE.json configs C1..C5), GPU next to the
compiled reference on one host core.  Not the driver's bench (that is bench.py, config C2): this is the per-row
evidence table of DESIGN.md §6.    python tools/bench_all.py > profiles/r01_all_decoders.json
"""
import importlib.util
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from codes import load_code                                   # noqa: E402
from oracle import pyoracle as po                             # noqa: E402

spec = importlib.util.spec_from_file_location("pyldpcb200", os.path.join(ROOT, "ldpc-lib_b200", "pyldpcb200.py"))
L = importlib.util.module_from_spec(spec)
sys.modules["pyldpcb200"] = L
spec.loader.exec_module(L)

# (label, code, Z, decoder, maxiter, Eb/N0 dB, modulation, punctured blocks, precision, GPU frames, CPU frames)
CASES = [
    ("C1 TASP_DEC 50it", "ref32x16_b", 126, "TASP", 50, 2.0, 0, 0, 64, 20000, 60),
    ("C1 ASP_DEC 50it", "ref32x16_b", 126, "ASP", 50, 2.0, 0, 0, 64, 20000, 60),
    ("C1 LMS_DEC 15it fp32", "ref32x16_b", 126, "LMS", 15, 2.0, 0, 0, 32, 400000, 200),
    ("C1 LMS_DEC 15it double", "ref32x16_b", 126, "LMS", 15, 2.0, 0, 0, 64, 40000, 200),
    ("C2 LMS_DEC 10it fp32", "ref32x16_b", 256, "LMS", 10, 2.0, 0, 0, 32, 400000, 100),
    ("C2 MS_DEC 10it fp32", "ref32x16_b", 256, "MS", 10, 3.0, 0, 0, 32, 400000, 60),
    ("C2 MS_DEC 10it double", "ref32x16_b", 256, "MS", 10, 3.0, 0, 0, 64, 20000, 60),
    ("C3 LMS_DEC 10it QAM-64 fp32", "c3_bg1_46x68", 384, "LMS", 10, 1.5, 3, 2, 32, 20000, 20),
    ("C3 LMS_DEC 10it QAM-64 double", "c3_bg1_46x68", 384, "LMS", 10, 1.5, 3, 2, 64, 4000, 20),
    ("C4 BP_DEC 20it", "c4_wifi_12x24", 81, "BP", 20, 2.0, 0, 0, 64, 40000, 100),
    ("C4 ASP_DEC 20it", "c4_wifi_12x24", 81, "ASP", 20, 2.0, 0, 0, 64, 40000, 100),
    ("C4 SP_DEC 20it", "c4_wifi_12x24", 81, "SP", 20, 2.0, 0, 0, 64, 40000, 100),
    ("C4 TASP_DEC 20it", "c4_wifi_12x24", 81, "TASP", 20, 2.0, 0, 0, 64, 40000, 100),
    ("C4 LCHE_DEC 20it", "c4_wifi_12x24", 81, "LCHE", 20, 2.0, 0, 0, 64, 40000, 60),
    ("C4 LMS_DEC 20it fp32 (NVRTC)", "c4_wifi_12x24", 81, "LMS", 20, 2.0, 0, 0, 32, 400000, 200),
    ("C5 IMS_DEC 15it", "ref32x16_a", 126, "IMS", 15, 3.0, 0, 0, 64, 400000, 200),
    ("C2-size IMS_DEC 10it (NVRTC)", "ref32x16_b", 256, "IMS", 10, 3.0, 0, 0, 64, 400000, 100),
    ("C5 IASP_DEC 15it", "ref32x16_a", 126, "IASP", 15, 3.0, 0, 0, 64, 40000, 100),
]


def main():
    rows = []
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    for label, code, Z, dec, maxiter, snr, mod, punct, prec, nf_gpu, nf_cpu in CASES:
        if only and only not in label:
            continue
        hd, _ = load_code(code)
        did = getattr(po, dec)
        K = (hd.shape[1] - hd.shape[0]) * Z
        with L.Decoder(hd, Z, did, precision=prec, use_fast=2) as d:
            info = d.kernel_info()
            d.simulate(snr, min(nf_gpu, 2000), maxiter, modulation=mod, punct=punct, seed=1)          # warm-up
            t0 = time.perf_counter()
            sim = d.simulate(snr, nf_gpu, maxiter, modulation=mod, punct=punct, seed=1, stream=1)
            wall = time.perf_counter() - t0
            ms, _ = d.last_kernel_ms()
            llr = d.generate_llr(snr, nf_cpu, modulation=mod, punct=punct, seed=1, stream=1)
            got = d.decode(llr, maxiter)
        ref_avail = po.have_ref()
        t0 = time.perf_counter()
        if ref_avail:
            want = po.ref_decode(did, hd, Z, llr.astype(np.float64), maxiter, fresh=(dec == "BP"), want_post=False)
        else:
            want = po.orc_decode(did, hd, Z, llr.astype(np.float64), maxiter)
        cpu = time.perf_counter() - t0
        bad = (got["iters"] != want["iters"]) | (got["hard"] != want["hard"]).any(axis=1)
        rows.append({"case": label, "N": hd.shape[1] * Z, "K": K, "kernel": info["name"], "gpu_frames": nf_gpu,
                     "gpu_info_gbps_kernel": nf_gpu * K / (ms * 1e-3) / 1e9, "gpu_info_gbps_call": nf_gpu * K / wall / 1e9,
                     "avg_iterations": sim["iter_sum"] / sim["frames"], "fer": sim["frame_errors"] / sim["frames"],
                     "cpu_kind": "reference" if ref_avail else "port", "cpu_info_mbps_one_core": nf_cpu * K / cpu / 1e6,
                     "parity_frames": int(nf_cpu), "parity_mismatch_frames": int(bad.sum())})
        print(json.dumps(rows[-1]), file=sys.stderr)
    print(json.dumps({"rows": rows, "note": "early exit on (reference semantics); GPU = ldpcb200_simulate (noise + LLR generated in the "
                      "decoder's first load); CPU = compiled reference decoder on the same LLR buffers, one core"}, indent=1))


if __name__ == "__main__":
    main()
