#!/usr/bin/env python
"""FER / BER-vs-SNR curves of the benchmark configurations on the GPU, with the reference's stop rule (frame errors per
point) and, where BASELINE.md records the reference's own curve, a check that each point falls inside the reference's
95 % (Wilson) interval.      python tools/fer_curves.py > profiles/r01_fer_curves.json
"""
import importlib.util
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from codes import load_code                                   # noqa: E402


def load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


L = load("pyldpcb200", os.path.join(ROOT, "ldpc-lib_b200", "pyldpcb200.py"))
SH = load("simhost", os.path.join(ROOT, "ldpc-lib_b200", "simhost.py"))


def wilson(k, n, z=1.96):
    p = k / n
    d = 1 + z * z / n
    c = (p + z * z / (2 * n)) / d
    h = z * np.sqrt(p * (1 - p) / n + z * z / (4 * n * n)) / d
    return c - h, c + h


# reference curves recorded in BASELINE.md §2 (reference `main simulation`, REF-32x16-B, Z = 126, 50 iterations,
# 50 frame errors per point, abort rule off)
REF_C1 = {"snr": [1.0, 1.25, 1.5, 1.75, 2.0, 2.25, 2.5],
          "TASP": [0.373, 0.0899, 0.0457, 0.0263, 0.0182, 0.0113, 0.00735],
          "LMS": [0.769, 0.427, 0.158, 0.0734, 0.0478, 0.0263, 0.0170]}

# (label, code, Z, decoder, precision, maxiter, snrs, frame errors, max frames, reference curve or None[, modulation, punctured blocks])
RUNS = [
    # BASELINE config 4 as written: BP flooding vs layered, FER down to 1e-6 (>= 30 frame errors per point), 8 GPUs
    ("C4F LMS_DEC 20it (layered)", "c4_wifi_12x24", 81, "LMS_DEC", 32, 20, [2.0, 2.25, 2.5, 2.75], 100, 400000000, None),
    ("C4F TASP_DEC 20it (layered)", "c4_wifi_12x24", 81, "TASP_DEC", 64, 20, [1.75, 2.0, 2.25, 2.5], 100, 200000000, None),
    ("C4F BP_DEC 20it (flooding)", "c4_wifi_12x24", 81, "BP_DEC", 64, 20, [2.0, 2.25, 2.5, 2.75, 3.0], 60, 120000000, None),
    ("C4F ASP_DEC 20it (flooding)", "c4_wifi_12x24", 81, "ASP_DEC", 64, 20, [2.0, 2.25, 2.5, 2.75, 3.0], 60, 120000000, None),
    ("C4G TASP_DEC 20it (layered)", "c4_wifi_12x24", 81, "TASP_DEC", 64, 20, [2.75, 3.0, 3.25, 3.5], 60, 300000000, None),
    ("C4G ASP_DEC 20it (flooding)", "c4_wifi_12x24", 81, "ASP_DEC", 64, 20, [3.25, 3.5], 60, 300000000, None),
    # BASELINE config 3 as written: BG1-shaped Z = 384, QAM-64 demodulation fused with layered min-sum, 8 GPUs
    ("C3 LMS_DEC 10it QAM-64 punct 2", "c3_bg1_46x68", 384, "LMS_DEC", 32, 10, [0.5, 0.75, 1.0, 1.5], 200, 3000000, None, 3, 2),
    ("C1 TASP_DEC 50it", "ref32x16_b", 126, "TASP_DEC", 64, 50, REF_C1["snr"], 400, 400000, REF_C1["TASP"]),
    ("C1 LMS_DEC 50it", "ref32x16_b", 126, "LMS_DEC", 32, 50, REF_C1["snr"], 2000, 4000000, REF_C1["LMS"]),
    ("C4 LMS_DEC 20it (layered)", "c4_wifi_12x24", 81, "LMS_DEC", 32, 20, [1.0, 1.5, 2.0, 2.25, 2.5, 2.75, 3.0], 200, 300000000, None),
    ("C4 TASP_DEC 20it (layered)", "c4_wifi_12x24", 81, "TASP_DEC", 64, 20, [1.0, 1.5, 2.0, 2.25], 100, 6000000, None),
    ("C4 BP_DEC 20it (flooding)", "c4_wifi_12x24", 81, "BP_DEC", 64, 20, [1.0, 1.5, 2.0, 2.5], 100, 6000000, None),
    ("C4 ASP_DEC 20it (flooding)", "c4_wifi_12x24", 81, "ASP_DEC", 64, 20, [1.0, 1.5, 2.0, 2.5], 100, 6000000, None),
    ("C5 IMS_DEC 15it (fixed point)", "ref32x16_a", 126, "IMS_DEC", 64, 15, [2.0, 2.5, 3.0, 3.5, 4.0, 4.5], 1000, 20000000, None),
]


def main():
    """Under torchrun (one rank per GPU) the frames of every point are striped over the ranks (simhost.frame_loop: one
    all-reduce of four counters per round); results do not depend on the number of ranks.  FIXED_FRAMES=n in the
    environment decodes exactly n frames per point instead of stopping at the error count (timing runs)."""
    out = []
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    group = device = None
    if world > 1:
        import torch
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        device = torch.device("cuda", local)
        dist.init_process_group("nccl", device_id=device)
        group = dist.group.WORLD
    fixed = int(os.environ.get("FIXED_FRAMES", "0"))
    for run in RUNS:
        label, code, Z, dec, prec, maxiter, snrs, nerr, nmax, ref = run[:10]
        mod, punct = (run[10], run[11]) if len(run) > 10 else (0, 0)
        if only and only not in label:
            continue
        hd, _ = load_code(code)
        pts = []
        if fixed:
            nerr, nmax = 1 << 62, fixed - 1
        with L.Decoder(hd, Z, getattr(L, dec), precision=prec, use_fast=2, device=local) as d:
            for k, snr in enumerate(snrs):
                t0 = time.perf_counter()
                ber, fer, r = SH.bp_simulation(d, maxiter, nerr, nmax, snr, 1.0, modulation=mod, punctured_blocks=punct, seed=1, stream=k,
                                               round_frames=1 << 14, max_round_frames=1 << 20, group=group, device=device)
                secs = time.perf_counter() - t0
                pt = {"snr_db": snr, "fer": fer, "ber": ber, "frames": r.experiment, "frame_errors": r.nde, "undetected": r.nue,
                      "seconds": round(secs, 2), "info_gbps_all_gpus": r.decoded * d.K / secs / 1e9}
                if ref:
                    lo, hi = wilson(50, 50 / ref[k])
                    mylo, myhi = wilson(r.nde, r.experiment)
                    pt.update(reference_fer=ref[k], reference_ci95=[lo, hi], inside=bool(myhi >= lo and mylo <= hi))
                pts.append(pt)
                if rank == 0:
                    print(label, pt, file=sys.stderr)
            out.append({"case": label, "kernel": d.kernel_info()["name"], "points": pts})
    if rank == 0:
        txt = json.dumps({"runs": out, "n_gpus": world,
                          "note": "stop rule: n frame errors or the frame budget, applied in frame order (simhost.frame_loop)"}, indent=1)
        if len(sys.argv) > 2:                    # libraries print banners on stdout (NCCL): a file keeps the JSON clean
            with open(sys.argv[2], "w") as f:
                f.write(txt + "\n")
        else:
            print(txt)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
