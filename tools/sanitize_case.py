#!/usr/bin/env python
"""A small, fixed workload for compute-sanitizer: every kernel family once (table-driven double + fp32, the
table-driven shared-memory LMS kernel, the ahead-of-time and run-time compiled specialised kernels, simulate with the
fused channel, QAM-64 first load).  usage: compute-sanitizer --tool memcheck|racecheck python tools/sanitize_case.py"""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from codes import load_code, awgn_llr          # noqa: E402

spec = importlib.util.spec_from_file_location("pyldpcb200", os.path.join(ROOT, "ldpc-lib_b200", "pyldpcb200.py"))
L = importlib.util.module_from_spec(spec)
sys.modules["pyldpcb200"] = L
spec.loader.exec_module(L)

hd, _ = load_code("ref32x16_b")
b, c = hd.shape
rng = np.random.default_rng(0)
for Z, fast in ((256, 1), (126, 1), (126, 0)):
    llr = awgn_llr(rng, 6, c * Z, b, c, 2.0, dtype=np.float32)
    os.environ["LDPCB200_NO_SPEC"] = "0" if fast else "1"
    with L.Decoder(hd, Z, L.LMS_DEC, precision=32) as d:
        print(d.kernel_info()["name"], d.decode(llr, 5, want_post=True)["iters"])
        print(d.simulate(2.0, 8, 5, seed=1)["frame_errors"])
os.environ["LDPCB200_NO_SPEC"] = "0"
hd4, _ = load_code("c4_wifi_12x24")
llr = awgn_llr(rng, 6, 24 * 81, 12, 24, 2.0, dtype=np.float32)
with L.Decoder(hd4, 81, L.LMS_DEC, precision=32, use_fast=2) as d:
    print(d.kernel_info()["name"], d.decode(llr, 5)["iters"])
for dec in (L.LMS_DEC, L.MS_DEC, L.IMS_DEC, L.TASP_DEC, L.ASP_DEC, L.BP_DEC, L.SP_DEC, L.LCHE_DEC, L.IASP_DEC):
    with L.Decoder(hd4, 81, dec) as d:
        print(L.DECODER_NAMES[dec], d.decode(llr.astype(np.float64), 4)["iters"], d.simulate(2.0, 4, 3, seed=2)["frames"])
hd3, _ = load_code("c3_bg1_46x68")
with L.Decoder(hd3, 96, L.LMS_DEC, precision=32, use_fast=2) as d:
    print(d.kernel_info()["name"], d.simulate(3.0, 6, 4, modulation=L.MOD_QAM64, punct=2, seed=3))
print("sanitize case done")
