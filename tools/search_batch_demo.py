#!/usr/bin/env python
"""The code-search caller's inner loop on the engine: K candidate matrices of the shape of
files/input32_16.jsonx's code (16 x 32, Z = 126, TASP_DEC, 50 iterations, one SNR, a few thousand frames each -- what
main_good_code_search.cpp:320-338 asks of bp_simulation per candidate), scored (a) one candidate per handle and call, as
the level-2 link does, and (b) all in one ldpcb200_simulate_codes launch on one handle.
    python tools/search_batch_demo.py [K] [frames] > profiles/r02_search_batch.json"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "ldpc-lib_b200")); sys.path.insert(0, ROOT)
from codes import load_code                                   # noqa: E402
import pyldpcb200 as L                                        # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nf = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
hd, _ = load_code("ref32x16_b")
Z, snr, it = 126, 2.0, 50
rng = np.random.default_rng(1)
hds = []
for k in range(K):
    m = hd.copy()
    nz = m >= 0
    m[nz] = rng.integers(0, Z, int(nz.sum()))
    hds.append(m)
hds = np.stack(hds)

with L.Decoder(hd, Z, L.TASP_DEC) as d:                       # CUDA context + warm-up
    d.simulate(snr, 100, it, seed=1)
t0 = time.perf_counter()
single = []
for k in range(K):
    with L.Decoder(hds[k], Z, L.TASP_DEC) as d:
        single.append(d.simulate(snr, nf, it, seed=1))
t_single = time.perf_counter() - t0
with L.Decoder(hd, Z, L.TASP_DEC) as d:
    d.simulate_codes(hds[:2], snr, 16, it, seed=1)
    t0 = time.perf_counter()
    batch = d.simulate_codes(hds, snr, nf, it, seed=1)
    t_batch = time.perf_counter() - t0
    ms, _ = d.last_kernel_ms()
assert batch == single, "batched results differ from single calls"
fer = sorted(b["frame_errors"] / b["frames"] for b in batch)
print(json.dumps({"codes": K, "frames_per_code": nf, "shape": "16 x 32, Z = 126, TASP_DEC, 50 iterations, %.1f dB" % snr,
                  "one_handle_and_call_per_code_s": t_single, "one_launch_s": t_batch, "kernel_ms": ms,
                  "speedup": t_single / t_batch, "identical_results": True,
                  "fer_best_median_worst": [fer[0], fer[len(fer) // 2], fer[-1]],
                  "info_gbps_one_launch": K * nf * 2016 / t_batch / 1e9}, indent=1))
