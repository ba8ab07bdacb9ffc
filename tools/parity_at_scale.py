#!/usr/bin/env python
"""Parity at scale: the GPU decoders against the COMPILED REFERENCE (oracle/_ref, all host cores) on identical LLR buffers,
with the frame counts the north-star bars are stated for -- fixed-point bit-exact on >= 10^5 frames, float decoders
>= 99.99 % frames with identical decisions and iteration counts, posteriors within 1e-4.  Writes one JSON document:
    python tools/parity_at_scale.py [CASE SUBSTRING] > profiles/r02_parity_at_scale.json
The reference side is the checker here, never the thing shipped (falls back to the C oracle port when oracle/_ref is absent)."""
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "ldpc-lib_b200"))
from codes import load_code                                   # noqa: E402

# (label, code, Z, decoder, GPU precision, maxiter, Eb/N0 dB, frames)
CASES = [
    ("C2 LMS_DEC fp32 vs reference double", "ref32x16_b", 256, "LMS", 32, 10, 2.0, 200000),
    ("C2 LMS_DEC fp32 vs reference double", "ref32x16_b", 256, "LMS", 32, 10, 3.0, 200000),
    ("C2 MS_DEC fp32 vs reference double", "ref32x16_b", 256, "MS", 32, 10, 3.0, 100000),
    ("C5 IMS_DEC fixed point", "ref32x16_a", 126, "IMS", 64, 15, 3.0, 200000),
    ("C5 IMS_DEC fixed point", "ref32x16_a", 126, "IMS", 64, 15, 4.0, 100000),
    ("C1 TASP_DEC double", "ref32x16_b", 126, "TASP", 64, 50, 2.0, 100000),
    ("C1 ASP_DEC double", "ref32x16_b", 126, "ASP", 64, 50, 2.0, 50000),
    ("C4 LMS_DEC fp32 vs reference double", "c4_wifi_12x24", 81, "LMS", 32, 20, 2.0, 200000),
    ("C4 LCHE_DEC double", "c4_wifi_12x24", 81, "LCHE", 64, 20, 2.0, 50000),
    ("C4 BP_DEC double (product form)", "c4_wifi_12x24", 81, "BP", 64, 20, 2.0, 100000),
    ("C4 SP_DEC double (own message divided out)", "c4_wifi_12x24", 81, "SP", 64, 20, 2.0, 100000),
]


def _worker(args):
    dec, code, Z, maxiter, llr, want_post = args
    from oracle import pyoracle as po
    hd, _ = load_code(code)
    did = getattr(po, dec)
    if po.have_ref():
        r = po.ref_decode(did, hd, Z, llr, maxiter, fresh=(dec == "BP"), want_post=want_post)
    else:
        r = po.orc_decode(did, hd, Z, llr, maxiter)
    return r["hard"], r["iters"], r["post"] if want_post else None


def main():
    import pyldpcb200 as L
    from oracle import pyoracle as po
    cores = os.cpu_count() or 1
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    rows = []
    for label, code, Z, dec, prec, maxiter, snr, nf in CASES:
        if only and only not in label:
            continue
        hd, _ = load_code(code)
        did = getattr(L, dec + "_DEC")
        bad = post_bad = post_n = 0
        worst_rel = 0.0
        t_cpu = t_gpu = 0.0
        chunk = 20000
        kernel = None
        with L.Decoder(hd, Z, did, precision=prec, use_fast=2) as d:
            kernel = d.kernel_info()["name"]
            for f0 in range(0, nf, chunk):
                n = min(chunk, nf - f0)
                llr = d.generate_llr(snr, n, seed=7, stream=int(snr * 100), first_frame=f0)     # fp32 values of the engine's channel
                want_post = f0 == 0 and dec in ("LMS", "MS", "IMS")
                t0 = time.perf_counter()
                got = d.decode(llr, maxiter, want_post=want_post)
                t_gpu += time.perf_counter() - t0
                ref_in = llr.astype(np.float64)
                t0 = time.perf_counter()
                with mp.get_context("fork").Pool(cores) as pool:
                    res = pool.map(_worker, [(dec, code, Z, maxiter, ref_in[i::cores], want_post) for i in range(cores)])
                t_cpu += time.perf_counter() - t0
                r_hard = np.zeros((n, d.N), np.uint8); r_iters = np.zeros(n, np.int32)
                r_post = np.zeros((n, d.N), np.float64) if want_post else None
                for i, (h, it, p) in enumerate(res):
                    r_hard[i::cores] = h; r_iters[i::cores] = it
                    if want_post:
                        r_post[i::cores] = p
                m = (got["iters"] != r_iters) | (got["hard"] != r_hard).any(axis=1)
                bad += int(m.sum())
                if want_post:
                    g = got["post"].astype(np.float64)
                    if dec == "IMS":
                        post_bad += int((g != r_post).sum())
                    else:
                        ok = ~m                                              # frames whose decisions and iteration counts agree
                        rel = np.abs(g[ok] - r_post[ok]) / np.maximum(np.abs(r_post[ok]), 1.0)
                        worst_rel = max(worst_rel, float(rel.max()))
                        post_bad += int((rel > 1e-4).sum())
                        post_n = int(rel.size)
        K = (hd.shape[1] - hd.shape[0]) * Z
        row = {"case": label, "snr_db": snr, "frames": nf, "kernel": kernel, "reference": "compiled reference" if po.have_ref() else "C oracle port",
               "mismatch_frames": bad, "mismatch_fraction": bad / nf, "cpu_cores": cores,
               "cpu_info_mbps_all_cores": nf * K / t_cpu / 1e6, "gpu_call_info_gbps_host_buffers": nf * K / t_gpu / 1e9}
        if dec == "IMS":
            row["posterior_values_differing_first_chunk"] = post_bad
        elif dec in ("LMS", "MS"):
            row["posterior_max_error_over_max(|LLR|,1)_first_chunk"] = worst_rel
            row["posterior_values_above_1e-4_first_chunk"] = "%d of %d (agreeing frames)" % (post_bad, post_n)
        rows.append(row)
        print(json.dumps(row), file=sys.stderr)
    print(json.dumps({"rows": rows, "note": "identical LLR buffers (the engine's channel, fp32 values widened to double for the reference); "
                      "a frame mismatches when its iteration count or any hard decision differs"}, indent=1))


if __name__ == "__main__":
    main()
