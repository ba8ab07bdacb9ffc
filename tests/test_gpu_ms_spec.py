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
