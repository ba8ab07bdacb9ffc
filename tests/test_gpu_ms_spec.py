"""The code-specialised flooding min-sum kernels (ms_spec.cuh): MS_DEC in fp32 and the fixed-point IMS_DEC, ahead-of-time
and run-time compiled instances, against the oracle on identical buffers (bit-exact) and the compiled reference's golden
vectors."""
import os

import numpy as np
import pytest

from codes import load_code, awgn_llr
from conftest import ROOT

pytestmark = pytest.mark.gpu


def _case(code, Z, snr, nf, seed=21):
    hd, _ = load_code(code)
    b, c = hd.shape
    return hd, awgn_llr(np.random.default_rng(seed), nf, c * Z, b, c, snr)


@pytest.mark.parametrize("code,Z,snr,fast", [("ref32x16_a", 126, 3.0, 2), ("ref32x16_b", 126, 3.0, 3), ("c4_wifi_12x24", 81, 2.5, 3),
                                              ("ref32x16_b", 256, 3.0, 3)])
def test_ims_spec_bit_exact(ldpc, po, code, Z, snr, fast):
    """IMS_DEC: quantised channel values (incl. the per-frame double energy), ims_soft, decisions, iteration counts."""
    hd, llr = _case(code, Z, snr, 200)
    want = po.orc_decode(po.IMS, hd, Z, llr, 15)
    with ldpc.Decoder(hd, Z, po.IMS, use_fast=2) as d:
        assert d.kernel_info()["fast"] == fast, d.kernel_info()
        got = d.decode(llr, 15, want_post=True, want_aux=True)
        got32 = d.decode(llr.astype(np.float32), 15)
    assert np.array_equal(got["aux"], want["aux"])
    assert np.array_equal(got["iters"], want["iters"])
    assert np.array_equal(got["hard"], want["hard"])
    assert np.array_equal(got["post"], want["post"])
    want32 = po.orc_decode(po.IMS, hd, Z, llr.astype(np.float32).astype(np.float64), 15)
    assert np.array_equal(got32["iters"], want32["iters"]) and np.array_equal(got32["hard"], want32["hard"])


def test_ims_spec_reference_golden(ldpc, po):
    G = np.load(os.path.join(ROOT, "tests", "golden", "decoders_c4_z27.npz"))
    hd, Z, llr = G["hd"], int(G["Z"]), G["llr"]
    with ldpc.Decoder(hd, Z, po.IMS, use_fast=2) as d:
        assert d.kernel_info()["fast"] == 3
        got = d.decode(llr, int(G["maxiter"]), want_post=True, want_aux=True)
    assert np.array_equal(got["iters"], G["IMS_iters"])
    assert np.array_equal(np.packbits(got["hard"], axis=1), G["IMS_hard"])
    assert np.array_equal(got["post"].astype(np.float64), G["IMS_post"])
    assert np.array_equal(got["aux"], G["IMS_aux"])


@pytest.mark.parametrize("code,Z,snr,fast", [("ref32x16_b", 256, 3.0, 2), ("ref32x16_b", 126, 3.0, 3), ("c4_wifi_12x24", 81, 2.5, 3)])
def test_ms_spec_f32_bit_exact_vs_f32_oracle(ldpc, po, code, Z, snr, fast):
    hd, llr = _case(code, Z, snr, 200)
    llr = llr.astype(np.float32)
    want = po.orc_decode(po.MS, hd, Z, llr, 12, dtype=np.float32)
    with ldpc.Decoder(hd, Z, po.MS, precision=32, use_fast=2) as d:
        assert d.kernel_info()["fast"] == fast, d.kernel_info()
        got = d.decode(llr, 12, want_post=True)
        nx = d.decode(llr, 12, no_early_exit=True)
        z = d.decode(llr[:4], 0)
    assert np.array_equal(got["iters"], want["iters"])
    assert np.array_equal(got["hard"], want["hard"])
    assert np.array_equal(got["post"], want["post"])
    assert np.array_equal(nx["iters"], want["iters"])
    assert np.array_equal(z["hard"], (llr[:4] < 0).astype(np.uint8))


def test_ms_spec_f32_vs_reference_double(ldpc, po):
    """>= 99.99 % of frames with the double algorithm's decisions and iteration counts."""
    hd, llr = _case("ref32x16_b", 256, 3.0, 3000, seed=4)
    want = po.orc_decode(po.MS, hd, 256, llr, 10)
    with ldpc.Decoder(hd, 256, po.MS, precision=32) as d:
        got = d.decode(llr, 10)
    bad = (got["iters"] != want["iters"]) | (got["hard"] != want["hard"]).any(axis=1)
    assert bad.mean() <= 1e-4, bad.sum()


def test_ms_ims_simulate_consistency(ldpc, po):
    hd, _ = load_code("ref32x16_a")
    for dec, kw in ((po.IMS, {}), (po.MS, {"precision": 32})):
        with ldpc.Decoder(hd, 126, dec, use_fast=2, **kw) as d:
            sim = d.simulate(3.5, 300, 15, seed=5, want_per_frame=True)
            llr = d.generate_llr(3.5, 300, seed=5)
            out = d.decode(llr, 15)
        errs = out["hard"].sum(axis=1)
        assert sim["frame_errors"] == int((errs > 0).sum()) and sim["bit_errors"] == int(errs.sum())
        assert sim["iter_sum"] == int(np.abs(out["iters"]).sum())


@pytest.mark.parametrize("code,Z,snr,groups", [("ref32x16_a", 126, 3.0, 2), ("ref32x16_a", 126, 3.0, 1), ("c4_wifi_12x24", 81, 2.5, 2),
                                                ("c4_wifi_12x24", 81, 2.5, 1), ("ref32x16_b", 256, 2.0, 1)])
def test_ims_fp16_pairs_two_frames_per_cta(ldpc, po, monkeypatch, code, Z, snr, groups):
    """ims_h2.cuh: frames as fp16 pairs, 2 or 4 slots per CTA that are refilled as their frames stop.  Odd batches (slots run
    dry at different times), a single frame, neighbours that stop at different iterations, fixed iterations, zero iterations --
    all bit-exact against the oracle, and equal to the one-frame-per-CTA kernel (LDPCB200_IMS_H2=0)."""
    monkeypatch.setenv("LDPCB200_IMS_H2_GROUPS", str(groups))
    hd, llr = _case(code, Z, snr, 151, seed=33)
    want = po.orc_decode(po.IMS, hd, Z, llr, 15)
    assert len(set(want["iters"].tolist())) > 3                          # partners stop at different iterations
    with ldpc.Decoder(hd, Z, po.IMS, use_fast=2) as d:
        info = d.kernel_info()
        assert info["frames_per_cta"] == 2 * groups and info["tmem"], info
        got = d.decode(llr, 15, want_post=True, want_aux=True)
        one = d.decode(llr[17:18], 15, want_post=True)
        nx = d.decode(llr, 15, no_early_exit=True)
        z = d.decode(llr[:5], 0, want_post=True)
        sim = d.simulate(snr, 101, 15, seed=8, want_per_frame=True)
        again = d.decode(d.generate_llr(snr, 101, seed=8), 15)
    for key in ("aux", "iters", "hard", "post"):
        assert np.array_equal(got[key], want[key]), key
    assert np.array_equal(one["iters"], want["iters"][17:18]) and np.array_equal(one["post"], want["post"][17:18])
    assert np.array_equal(nx["iters"], want["iters"])
    assert np.array_equal(z["iters"], np.zeros(5, np.int32))              # no pass ran: decisions of the QUANTISED channel values
    assert np.array_equal(z["hard"], (want["aux"][:5] < 0).astype(np.uint8)) and np.array_equal(z["post"], want["aux"][:5])
    errs = again["hard"].sum(axis=1)
    assert np.array_equal(sim["per_frame"] >> 31, (errs > 0).astype(np.uint32))
    assert sim["iter_sum"] == int(np.abs(again["iters"]).sum()) and sim["frames"] == 101
    monkeypatch.setenv("LDPCB200_IMS_H2", "0")
    with ldpc.Decoder(hd, Z, po.IMS, use_fast=2) as d:
        assert d.kernel_info()["frames_per_cta"] == 1
        ref = d.decode(llr, 15, want_post=True)
        nx1 = d.decode(llr, 15, no_early_exit=True, want_post=True)
    assert np.array_equal(ref["post"], got["post"]) and np.array_equal(ref["iters"], got["iters"])
    with ldpc.Decoder(hd, Z, po.IMS, use_fast=2, dbits=9) as d:           # dbits = 9: integers up to 255 * 12 -- not this kernel
        assert d.kernel_info()["frames_per_cta"] == 1


def test_ims_fp16_pairs_other_alpha(ldpc, po):
    """Every scaling constant ialpha = 0 .. 16 of the fp16-pair kernel against the oracle (run-time compiled instance)."""
    hd, llr = _case("c4_wifi_12x24", 81, 2.5, 40, seed=2)
    for alpha in (0.0, 0.07, 0.3, 0.5, 0.8125, 0.95, 1.0):
        want = po.orc_decode(po.IMS, hd, 81, llr, 10, alpha=alpha)
        with ldpc.Decoder(hd, 81, po.IMS, use_fast=2, alpha=alpha) as d:
            assert d.kernel_info()["frames_per_cta"] >= 2
            got = d.decode(llr, 10, want_post=True)
        assert np.array_equal(got["iters"], want["iters"]) and np.array_equal(got["post"], want["post"]), alpha


def test_ims_energy_fallback_on_values_at_a_rounding_boundary(ldpc, po):
    """ims_h2.cuh sums the frame energy in parallel and only falls back to the reference's sequential order when a value lands
    within 1e-9 of a quantiser boundary.  Frames built to sit ON boundaries (one value solved so that |y| coef max_quant / thr
    = k + 1/2 to the last bit, where the order of the additions decides the result) must still be the oracle's, value for value."""
    hd, _ = load_code("ref32x16_a")
    Z, N = 126, 32 * 126
    rng = np.random.default_rng(77)
    frames = []
    for k in (0, 3, 10, 17, 25, 30):
        y = rng.normal(1.0, 0.7, N) * rng.choice([-1.0, 1.0], N)
        j = int(rng.integers(0, N))
        b = 0.3
        for _ in range(200):                                        # fixed point of b coef(b) 31 / 1.4 = k + 1/2
            y[j] = b
            en = 0.0
            for v in y:                                             # the reference's order
                en += v * v
            b = (k + 0.5) * 1.4 / 31.0 / np.sqrt(N / en)
        y[j] = b
        frames.append(y)
    llr = np.array(frames + [rng.normal(1.0, 1.0, N) for _ in range(10)])
    want = po.orc_decode(po.IMS, hd, Z, llr, 8)
    with ldpc.Decoder(hd, Z, po.IMS, use_fast=2) as d:
        assert d.kernel_info()["frames_per_cta"] >= 2
        got = d.decode(llr, 8, want_post=True, want_aux=True)
    for key in ("aux", "iters", "hard", "post"):
        assert np.array_equal(got[key], want[key]), key
