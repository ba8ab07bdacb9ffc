"""GPU parity of the table-driven kernels against the CPU oracle on identical LLR buffers, through the C ABI.

Bars (BASELINE.json north_star): integer decoders (IMS_DEC, IASP_DEC) bit-exact; LMS_DEC / MS_DEC in double
bit-exact too (add / subtract / compare / one multiply, same order, no FMA); the exp/log decoders identical
decisions and iteration counts on >= 99.99 % of frames and posteriors within 1e-4 relative.
"""
import numpy as np
import pytest

from codes import load_code, awgn_llr

pytestmark = pytest.mark.gpu

CASES = [("ref32x16_b", 126, 2.0), ("c4_wifi_12x24", 81, 2.0)]


def _llr(code, Z, snr, nf, seed=7):
    hd, _ = load_code(code)
    b, c = hd.shape
    return hd, awgn_llr(np.random.default_rng(seed), nf, c * Z, b, c, snr)


@pytest.mark.parametrize("code,Z,snr", CASES)
@pytest.mark.parametrize("dec", ["LMS", "MS"])
def test_minsum_f64_bit_exact(ldpc, po, code, Z, snr, dec):
    hd, llr = _llr(code, Z, snr, 96)
    did = getattr(po, dec)
    want = po.orc_decode(did, hd, Z, llr, 15)
    with ldpc.Decoder(hd, Z, did, precision=64, use_fast=False) as d:
        got = d.decode(llr, 15, want_post=True)
    assert np.array_equal(got["iters"], want["iters"])
    assert np.array_equal(got["hard"], want["hard"])
    assert np.array_equal(got["post"], want["post"])          # bitwise equal doubles


@pytest.mark.parametrize("code,Z,snr", CASES)
@pytest.mark.parametrize("dec", ["LMS", "MS"])
def test_minsum_f32_bit_exact_vs_f32_oracle(ldpc, po, code, Z, snr, dec):
    hd, llr = _llr(code, Z, snr, 96)
    llr = llr.astype(np.float32)
    did = getattr(po, dec)
    want = po.orc_decode(did, hd, Z, llr, 15, dtype=np.float32)
    with ldpc.Decoder(hd, Z, did, precision=32, use_fast=False) as d:
        got = d.decode(llr, 15, want_post=True)
    assert np.array_equal(got["iters"], want["iters"])
    assert np.array_equal(got["hard"], want["hard"])
    assert np.array_equal(got["post"], want["post"])


@pytest.mark.parametrize("code,Z,snr", [("ref32x16_a", 126, 3.0), ("c4_wifi_12x24", 81, 2.5)])
def test_ims_bit_exact(ldpc, po, code, Z, snr):
    hd, llr = _llr(code, Z, snr, 96)
    want = po.orc_decode(po.IMS, hd, Z, llr, 15)
    with ldpc.Decoder(hd, Z, po.IMS, use_fast=False) as d:
        got = d.decode(llr, 15, want_post=True, want_aux=True)
    assert np.array_equal(got["aux"], want["aux"])            # quantiser incl. the per-frame energy
    assert np.array_equal(got["iters"], want["iters"])
    assert np.array_equal(got["hard"], want["hard"])
    assert np.array_equal(got["post"], want["post"])


@pytest.mark.parametrize("code,Z,snr", CASES)
def test_iasp_bit_exact(ldpc, po, code, Z, snr):
    hd, llr = _llr(code, Z, snr, 64)
    want = po.orc_decode(po.IASP, hd, Z, llr, 15)
    with ldpc.Decoder(hd, Z, po.IASP) as d:
        got = d.decode(llr, 15, want_post=True)
    # the only float step is the input quantiser 1/(1+exp(v)) (decoders.cpp:3849-3861): a last-ulp
    # difference of exp() could move a value across a rounding boundary, so allow 1e-4 of frames
    bad = (got["iters"] != want["iters"]) | (got["hard"] != want["hard"]).any(axis=1)
    assert bad.mean() <= 1e-4 + 1.0 / len(bad) * 0, bad.sum()
    assert np.array_equal(got["post"][~bad], want["post"][~bad])


@pytest.mark.parametrize("code,Z,snr", CASES)
@pytest.mark.parametrize("dec", ["TASP", "ASP", "BP", "SP", "LCHE"])
def test_float_sumprod(ldpc, po, code, Z, snr, dec):
    hd, llr = _llr(code, Z, snr, 64)
    did = getattr(po, dec)
    want = po.orc_decode(did, hd, Z, llr, 15)
    with ldpc.Decoder(hd, Z, did) as d:
        got = d.decode(llr, 15, want_post=True)
    bad = (got["iters"] != want["iters"]) | (got["hard"] != want["hard"]).any(axis=1)
    assert bad.sum() == 0, (bad.sum(), got["iters"][:8], want["iters"][:8])
    a, w = got["post"], want["post"]
    rel = np.abs(a - w) / np.maximum(np.abs(w), 1e-300)
    assert np.max(rel) <= 1e-4, np.max(rel)


def test_bp_chain_syndrome(ldpc, po):
    """BP_DEC's stale-syndrome quirk (decoders.cpp:1742-1759) is reproduced when asked for."""
    hd, llr = _llr("ref32x16_b", 126, 1.2, 24)
    want = po.orc_decode(po.BP, hd, 126, llr, 6, chain=True)
    with ldpc.Decoder(hd, 126, po.BP) as d:
        got = d.decode(llr, 6, chain=True)
    assert np.array_equal(got["iters"], want["iters"])
    assert np.array_equal(got["hard"], want["hard"])


def test_packed_hard_and_edge_cases(ldpc, po):
    hd, llr = _llr("c4_wifi_12x24", 81, 3.0, 5)          # N = 1944 is not a multiple of 32
    with ldpc.Decoder(hd, 81, po.LMS, use_fast=False) as d:
        a = d.decode(llr, 10)
        p = d.decode(llr, 10, packed=True)
        bits = ((p["hard"][:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(5, -1)[:, :d.N]
        assert np.array_equal(bits.astype(np.uint8), a["hard"])
        e = d.decode(llr[:0], 10)                        # empty batch
        assert e["iters"].shape == (0,)
        z = d.decode(llr, 0)                             # zero iterations: LMS returns 1 / -0 (decoders.cpp:5424)
        want = po.orc_decode(po.LMS, hd, 81, llr, 0)
        assert np.array_equal(z["iters"], want["iters"]) and np.array_equal(z["hard"], want["hard"])
        big = np.full((2, d.N), 30.0)                    # an all-zero codeword far from the threshold
        assert np.array_equal(d.decode(big, 10)["iters"], [1, 1])
    with pytest.raises(ldpc.LdpcError):
        ldpc.Decoder(hd, 81, 6)                          # FHT_DEC (GF(q)) is out of scope


def test_c3_qam64_fused_llr_and_layered_min_sum(ldpc, po):
    """Config C3 shape: 46 x 68 'BG1-shaped' matrix, QAM-64, two punctured block columns, LMS_DEC.  The LLRs the
    decoder's first load generates equal Demodulate() (m = 6) of the received symbols, negated, and the decode of
    those LLRs equals the oracle's."""
    hd, _ = load_code("c3_bg1_46x68")
    Z, snr, nf = 96, 7.0, 24
    with ldpc.Decoder(hd, Z, po.LMS, precision=64) as d:
        llr = d.generate_llr(snr, nf, modulation=ldpc.MOD_QAM64, punct=2, seed=6, dtype=np.float64)
        sim = d.simulate(snr, nf, 10, modulation=ldpc.MOD_QAM64, punct=2, seed=6, want_per_frame=True)
        got = d.decode(llr, 10, want_post=True)
        N, R = d.N, d.R
    assert np.all(llr[:, N - 2 * Z:] == 0.5)                       # LLR-domain decoder: punctured value 0.5 (bp_simulation.cpp:700)
    want = po.orc_decode(po.LMS, hd, Z, llr, 10)
    assert np.array_equal(got["iters"], want["iters"]) and np.array_equal(got["hard"], want["hard"])
    assert np.array_equal(got["post"], want["post"])
    errs = got["hard"].sum(axis=1)
    assert sim["frame_errors"] == int((errs > 0).sum()) and sim["bit_errors"] == int(errs.sum())
    # LLR -> received symbol -> Demodulate round trip on the I component of the first symbols (bits 0..2)
    sigma = ldpc.sigma(46, 68, 2, snr, ldpc.MOD_QAM64)
    body = llr[:, :N - 2 * Z]
    assert (body > 0).mean() > 0.85                                # all-zero codeword
    # consistency with the function-boundary demodulator: re-demodulating any symbol whose three I-bit LLRs we hold
    # must reproduce them; recover x from the MSB LLR by bisection on the oracle's monotone bit-0 metric
    o = po.oracle()
    for f in range(2):
        for s in range(3):
            target = -body[f, 6 * s:6 * s + 3]                     # Demodulate's sign is log P1/P0
            lo, hi = -12.0, 12.0
            for _ in range(80):
                mid = 0.5 * (lo + hi)
                v = po.orc_demodulate(64, 1, sigma, np.array([mid, 0.0]))[0]
                lo, hi = (mid, hi) if v < target[0] else (lo, mid)
            back = po.orc_demodulate(64, 1, sigma, np.array([0.5 * (lo + hi), 0.0]))[:3]
            if abs(target[0]) < 25:                                # not clipped at T
                assert np.allclose(back, target, rtol=1e-5, atol=1e-5), (back, target)
