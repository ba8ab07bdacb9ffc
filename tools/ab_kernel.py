#!/usr/bin/env python
"""A/B timing of compile-time variants of a code-specialised kernel on one GPU (development tool).

Every variant is the run-time compiled kernel (LDPCB200_NO_AOT=1) with its own LDPCB200_JIT_DEFINES; "aot" is the
ahead-of-time instance as built.  For each: the device-resident decode of bench.py (fixed iterations and with the
reference's early exit), and a bitwise comparison of decisions, iteration counts and posteriors with the first
variant.

    python tools/ab_kernel.py --frames 131072 aot jit LMS_TMEM_PP=0 LMS_TMEM_PP=1,LMS_TMEM_FFMA2=1
(a variant is "aot", "jit" (no defines) or a comma-separated list of NAME=VALUE macro definitions)
"""
import argparse
import importlib.util
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def binding():
    spec = importlib.util.spec_from_file_location("pyldpcb200", os.path.join(ROOT, "ldpc-lib_b200", "pyldpcb200.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["pyldpcb200"] = mod
    spec.loader.exec_module(mod)
    return mod


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("variants", nargs="+")
    ap.add_argument("--code", default="ref32x16_b")
    ap.add_argument("--Z", type=int, default=256)
    ap.add_argument("--decoder", type=int, default=8)
    ap.add_argument("--maxiter", type=int, default=10)
    ap.add_argument("--snr", type=float, default=2.0)
    ap.add_argument("--frames", type=int, default=1 << 17)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--precision", type=int, default=32)
    args = ap.parse_args()
    import torch
    from codes import load_code
    L = binding()
    hd, _ = load_code(args.code)
    b, c = hd.shape
    dev = torch.device("cuda", 0)
    llr = hard = iters = None
    first = None
    for v in args.variants:
        os.environ.pop("LDPCB200_GRID_PER_SM", None)
        os.environ.pop("LDPCB200_TMEM2", None)
        if v == "t2" or v.startswith("t2,"):                        # "t2": two frames per CTA (lms_tmem2.cuh, run-time compiled); "t2,NAME=VALUE": with macros
            os.environ["LDPCB200_TMEM2"] = "1"
            v = "jit" if v == "t2" else v[3:]
        if v.startswith("grid"):                                    # "grid1": the ahead-of-time kernel with one CTA per SM
            os.environ["LDPCB200_GRID_PER_SM"] = v[4:]
            v = "aot"
        if v == "aot":
            os.environ.pop("LDPCB200_NO_AOT", None)
            os.environ.pop("LDPCB200_JIT_DEFINES", None)
        else:
            os.environ["LDPCB200_NO_AOT"] = "1"
            os.environ["LDPCB200_JIT_DEFINES"] = "" if v == "jit" else " ".join("-D" + d for d in v.split(","))
        try:
            dec = L.Decoder(hd, args.Z, args.decoder, precision=args.precision, device=0, use_fast=1 if v == "aot" else 2)
        except Exception as e:                                      # a variant that does not compile must not end the run
            print(json.dumps({"variant": v, "error": str(e)[:400]}), flush=True)
            continue
        info = dec.kernel_info()
        if llr is None:
            llr = torch.empty((args.frames, dec.N), dtype=torch.float32, device=dev)
            for f0 in range(0, args.frames, 1 << 16):
                n = min(1 << 16, args.frames - f0)
                dec.generate_llr(args.snr, n, seed=1, stream=0, first_frame=f0, out=llr[f0:f0 + n])
            hard = torch.empty((args.frames, dec.nwords), dtype=torch.int32, device=dev)
            iters = torch.empty(args.frames, dtype=torch.int32, device=dev)
            host = llr[:512].cpu().numpy()
        fixed, early = [], []
        for _ in range(args.reps):
            dec.decode_device(llr, args.maxiter, hard_words=hard, iters=iters, no_early_exit=True)
            fixed.append(dec.last_kernel_ms()[0])
        for _ in range(args.reps):
            dec.decode_device(llr, args.maxiter, hard_words=hard, iters=iters)
            early.append(dec.last_kernel_ms()[0])
        torch.cuda.synchronize()
        res = (hard.cpu().numpy().copy(), iters.cpu().numpy().copy())
        post = dec.decode(host, args.maxiter, want_post=True)
        gb = lambda ms: args.frames * dec.K / (ms * 1e-3) / 1e9
        line = {"variant": v, "kernel": info["name"], "tmem": info["tmem"], "ctas_per_sm": info["ctas_per_sm"], "smem": info["smem_bytes"],
                "fixed_ms": min(fixed), "fixed_gbps": gb(min(fixed)), "early_ms": min(early), "early_gbps": gb(min(early)),
                "avg_iters": float(np.abs(res[1]).mean())}
        if first is None:
            first = (res, post)
        else:
            line["same_hard"] = bool(np.array_equal(res[0], first[0][0]))
            line["same_iters"] = bool(np.array_equal(res[1], first[0][1]))
            line["same_post"] = bool(np.array_equal(post["post"].view(np.uint32), first[1]["post"].view(np.uint32)))
        print(json.dumps(line), flush=True)
        dec.close()


if __name__ == "__main__":
    main()
