#!/usr/bin/env python
"""Generates the golden fixtures in this directory from the UNMODIFIED reference (oracle/_ref, built by
`make -C oracle ref refmain` from /root/reference).  Run in the development container only:

    python tests/golden/make_golden.py [decoders] [demod] [jsonx] [sim]

  decoders_c4_z27.npz   identical LLR buffers -> the reference's decword / return value / posterior for
                        each of the nine binary decoders (12x24 "802.11n-shaped" matrix lifted to Z = 27)
  demod.npz             Demodulate() with m = log2(Q) and QAM_modulator() on seeded inputs
  jsonx_expected/*      settings::from_file -> to_file of the reference for tests/golden/jsonx_cases/*
  ref_sim_*.jsonx       result files of the reference's `main simulation` on configs/sim_c1_*.jsonx
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import pyoracle as po          # noqa: E402
from codes import load_code, awgn_llr      # noqa: E402


def decoders():
    hd, _ = load_code("c4_wifi_12x24")
    Z = 27
    b, c = hd.shape
    rng = np.random.default_rng(2026)
    llr = np.concatenate([awgn_llr(rng, 5, c * Z, b, c, 1.0), awgn_llr(rng, 5, c * Z, b, c, 3.0)])
    llr[9] = np.abs(llr[9]) + 0.5                       # an error-free frame: exercises the "already a codeword" returns
    out = {"hd": hd, "Z": np.int32(Z), "llr": llr, "maxiter": np.int32(12)}
    for dec in (po.BP, po.SP, po.ASP, po.MS, po.IMS, po.IASP, po.TASP, po.LMS, po.LCHE):
        r = po.ref_decode(dec, hd, Z, llr.copy(), 12, fresh=True)
        name = po.NAMES[dec]
        out[name + "_hard"] = np.packbits(r["hard"], axis=1)
        out[name + "_iters"] = r["iters"]
        out[name + "_post"] = r["post"]
        if dec == po.IMS:
            out[name + "_aux"] = r["aux"].astype(np.int16)
    r = po.ref_decode(po.BP, hd, Z, llr.copy(), 3, fresh=False)      # one state for all frames: the stale-syndrome quirk
    out["BP_chain_iters"] = r["iters"]
    out["BP_chain_hard"] = np.packbits(r["hard"], axis=1)
    np.savez_compressed(os.path.join(HERE, "decoders_c4_z27.npz"), **out)
    print("decoders:", {k: v.tolist() for k, v in out.items() if k.endswith("_iters")})


def demod():
    rng = np.random.default_rng(7)
    out = {}
    for Q in (16, 64, 256):
        m = int(np.log2(Q))
        ns = 400
        sq = int(np.sqrt(Q))
        x = rng.integers(0, sq, 2 * ns) * 2.0 - (sq - 1) + rng.standard_normal(2 * ns) * 0.9
        x[:8] = [-30, 30, 0, 0.5, -0.5, sq + 3.0, -(sq + 3.0), 1e-3]       # far outside, on thresholds
        sigma = 0.8
        out["x%d" % Q] = x
        out["sigma%d" % Q] = np.float64(sigma)
        with np.errstate(all="ignore"):
            out["llr%d" % Q] = po.ref_demodulate(Q, ns, sigma, x, 26.0, 0)
            out["p1_%d" % Q] = po.ref_demodulate(Q, ns, sigma, x, 26.0, 1)
        bits = rng.integers(0, 2, ns * m).astype(np.uint8)
        out["bits%d" % Q] = bits
        out["mod%d" % Q] = po.ref_modulate(Q, bits)
    x4 = rng.standard_normal(200)
    out["x4"] = x4
    out["llr4"] = po.ref_demodulate(4, 100, 0.7, x4, 26.0, 0)
    np.savez_compressed(os.path.join(HERE, "demod.npz"), **out)
    print("demod ok")


def encoder():
    """random_codeword() of the reference on the benchmark matrices (seeded): exit code and codeword."""
    out = {}
    for name, Z in (("ref32x16_a", 126), ("ref32x16_b", 126), ("ref32x16_b", 256), ("c4_wifi_12x24", 27), ("c3_bg1_46x68", 16)):
        hd, _ = load_code(name)
        hd = np.where(hd > 0, hd % Z, hd)
        for seed in (1, 5):
            rc, cw = po.ref_random_codeword(hd, Z, seed)
            key = "%s_Z%d_s%d" % (name, Z, seed)
            out[key + "_rc"] = np.int32(rc)
            if cw is not None:
                out[key + "_cw"] = np.packbits(cw)
    np.savez_compressed(os.path.join(HERE, "encoder.npz"), **out)
    print("encoder:", {k: int(v) for k, v in out.items() if k.endswith("_rc")})


JSONX_CASES = [("basic.jsonx", ""), ("matrices.jsonx", ""), ("refs_main.jsonx", ""), ("refs_main.jsonx", "settings"),
               ("refs_main.jsonx", "settings/more"),
               ("refs_main.jsonx", "top_default"), ("refs_main.jsonx", "settings/snrs"),
               ("refs_main.jsonx", "settings/via_inner_defaults"), ("refs_main.jsonx", "results")]


# select() paths on which the reference dies (the outer `defaults` is NOT consulted once the first key matched)
JSONX_DIES = [("refs_main.jsonx", "settings/fallback_only"), ("refs_main.jsonx", "nothing"), ("basic.jsonx", "alpha/x")]


INTERLEAVER_CASES = [("c4_wifi_12x24", 27, 0, 1, 1, 1), ("c4_wifi_12x24", 27, 3, 1, 1, 1), ("c4_wifi_12x24", 27, 2, 2, 1, 1),
                     ("c4_wifi_12x24", 27, 3, 2, 1, 1), ("c4_wifi_12x24", 27, 4, 2, 1, 1), ("c4_wifi_12x24", 27, 3, 3, 96, 1),
                     ("c4_wifi_12x24", 27, 3, 3, 100, 1), ("c4_wifi_12x24", 27, 2, 4, 1, 12), ("ref32x16_b", 21, 3, 2, 1, 1),
                     ("ref32x16_b", 21, 4, 2, 1, 1), ("ref32x16_b", 21, 2, 2, 1, 1), ("c3_bg1_46x68", 8, 3, 2, 1, 1)]


def interleaver():
    """The reference's interleaver tables (Permutations_Open / Permutation_Init / Permutation on index ramps)."""
    out = {"cases": np.array([c[1:] for c in INTERLEAVER_CASES], np.int32), "codes": np.array([c[0] for c in INTERLEAVER_CASES])}
    for k, (code, Z, mod, mode, block, inter) in enumerate(INTERLEAVER_CASES):
        hd, _ = load_code(code)
        d, i = po.ref_permutation(hd, Z, mod, mode, block, inter)
        out["direct_%d" % k], out["inverse_%d" % k] = d, i
    np.savez_compressed(os.path.join(HERE, "interleaver.npz"), **out)
    print("interleaver.npz written")


def jsonx_name(name, sel):
    return name.replace(".jsonx", "") + ("__" + sel.replace("/", "_") if sel else "") + ".jsonx"


def jsonx():
    cases = os.path.join(HERE, "jsonx_cases")
    exp = os.path.join(HERE, "jsonx_expected")
    os.makedirs(exp, exist_ok=True)
    tool = os.path.join(ROOT, "oracle", "_ref", "jsonx_rt")
    for name, sel in JSONX_CASES:
        out = os.path.join(exp, jsonx_name(name, sel))
        subprocess.check_call([tool, os.path.join(cases, name), sel, out])
    for name, sel in JSONX_DIES:
        rc = subprocess.call([tool, os.path.join(cases, name), sel, "/dev/null"], stderr=subprocess.DEVNULL)
        assert rc == 1, (name, sel, rc)
    print("jsonx:", sorted(os.listdir(exp)))


def sim():
    main = os.path.join(ROOT, "oracle", "_ref", "main")
    for name in ("sim_c1_lms", "sim_c1_tasp"):
        out = os.path.join(HERE, "ref_%s.jsonx" % name)
        if os.path.exists(out):
            os.remove(out)
        with open(os.path.join(HERE, "ref_%s.log" % name), "w") as log:
            subprocess.check_call([main, "simulation", os.path.join(ROOT, "configs", name + ".jsonx"), out], stdout=log, stderr=log)
        print("sim:", name, "done")


if __name__ == "__main__":
    what = sys.argv[1:] or ["decoders", "demod", "jsonx", "encoder", "sim", "interleaver"]
    po.build(ref=True)
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "refmain"])
    for w in what:
        globals()[w]()
