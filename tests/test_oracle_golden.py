"""The CPU oracle against (a) the committed golden vectors produced by the unmodified reference
(tests/golden/make_golden.py) and (b) the compiled reference itself when oracle/_ref is present."""
import numpy as np
import pytest

from codes import load_code, awgn_llr
from conftest import ROOT
import os

G = np.load(os.path.join(ROOT, "tests", "golden", "decoders_c4_z27.npz"))
D = np.load(os.path.join(ROOT, "tests", "golden", "demod.npz"))
ALL = ["BP", "SP", "ASP", "MS", "IMS", "IASP", "TASP", "LMS", "LCHE"]
EXACT = {"MS", "IMS", "IASP", "LMS", "LCHE"}           # no exp/log on the path (LCHE: table look-ups only)


@pytest.mark.parametrize("dec", ALL)
def test_oracle_matches_reference_golden(po, dec):
    hd, Z, llr, maxiter = G["hd"], int(G["Z"]), G["llr"], int(G["maxiter"])
    r = po.orc_decode(getattr(po, dec), hd, Z, llr, maxiter)
    assert np.array_equal(r["iters"], G[dec + "_iters"])
    assert np.array_equal(np.packbits(r["hard"], axis=1), G[dec + "_hard"])
    post = r["post"].astype(np.float64)
    if dec in EXACT:
        assert np.array_equal(post, G[dec + "_post"])
    else:
        # same libm here as where the vectors were made: equal in practice; 1e-12 guards a libm change
        assert np.allclose(post, G[dec + "_post"], rtol=1e-12, atol=1e-300)
    if dec == "IMS":
        assert np.array_equal(r["aux"], G["IMS_aux"])


def test_oracle_bp_chain_quirk(po):
    """BP_DEC re-uses the previous frame's syndrome for its pre-iteration check (decoders.cpp:1742-1759)."""
    r = po.orc_decode(po.BP, G["hd"], int(G["Z"]), G["llr"], 3, chain=True)
    assert np.array_equal(r["iters"], G["BP_chain_iters"])
    assert np.array_equal(np.packbits(r["hard"], axis=1), G["BP_chain_hard"])
    assert G["BP_chain_iters"][9] == 1 and G["BP_iters"][9] == 0      # the quirk is visible in the fixture


def test_return_conventions_in_golden():
    """SURVEY.md §8a': a frame that is already a codeword returns 1 from MS/IMS/LMS and 0 from the others."""
    for dec in ALL:
        assert G[dec + "_iters"][9] == (1 if dec in ("MS", "IMS", "LMS") else 0)
        assert G[dec + "_iters"][0] == -12


@pytest.mark.parametrize("Q", [16, 64, 256])
def test_oracle_demodulate_golden(po, Q):
    x, sigma = D["x%d" % Q], float(D["sigma%d" % Q])
    ns = x.size // 2
    with np.errstate(all="ignore"):
        llr = po.orc_demodulate(Q, ns, sigma, x, 26.0, 0)
        p1 = po.orc_demodulate(Q, ns, sigma, x, 26.0, 1)
    assert np.array_equal(np.isnan(llr), np.isnan(D["llr%d" % Q]))     # x = +-30: all points beyond T -> NaN (QAM_demodulator.cpp:198)
    assert np.array_equal(llr[~np.isnan(llr)], D["llr%d" % Q][~np.isnan(llr)])
    assert np.array_equal(p1[~np.isnan(p1)], D["p1_%d" % Q][~np.isnan(p1)])
    assert np.array_equal(po.orc_modulate(Q, D["bits%d" % Q]), D["mod%d" % Q])


def test_oracle_demodulate_qam4_golden(po):
    assert np.array_equal(po.orc_demodulate(4, 100, 0.7, D["x4"], 26.0, 0), D["llr4"])


def test_all_zero_codeword_maps_to_corner(po):
    """bits 0 -> natural index 0 -> gray[0] -> coordinate -(sqrt(Q)-1) (QAM_modulator.cpp:127-140)."""
    for Q in (16, 64, 256):
        m = int(np.log2(Q))
        assert np.all(po.orc_modulate(Q, np.zeros(5 * m, np.uint8)) == -(np.sqrt(Q) - 1))


def test_sigma_formulas(po):
    # bp_simulation.cpp:444-449
    assert po.sigma_bpsk(2.0, 16, 32, 0) == pytest.approx(np.sqrt(10 ** -0.2 / 2 / 0.5))
    assert po.sigma_qam(2.0, 46, 68, 2, 64) == pytest.approx(np.sqrt(10 ** -0.2 / (2 * (22 / 66) * 3 * 2) * 42.0))


needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libldpcref.so")),
                               reason="oracle/_ref not built (needs /root/reference)")


@needs_ref
@pytest.mark.parametrize("dec", ALL)
@pytest.mark.parametrize("code,Z,snr", [("ref32x16_b", 126, 2.0), ("ref32x16_a", 64, 2.5)])
def test_oracle_vs_compiled_reference(po, dec, code, Z, snr):
    hd, _ = load_code(code)
    b, c = hd.shape
    llr = awgn_llr(np.random.default_rng(3), 12, c * Z, b, c, snr)
    did = getattr(po, dec)
    r = po.ref_decode(did, hd, Z, llr.copy(), 10, fresh=True)
    o = po.orc_decode(did, hd, Z, llr.copy(), 10)
    assert np.array_equal(r["iters"], o["iters"])
    assert np.array_equal(r["hard"], o["hard"])
    assert np.allclose(r["post"], o["post"].astype(np.float64), rtol=1e-12, atol=1e-300)


@needs_ref
def test_f32_oracle_tracks_double_reference(po):
    """The fp32 restatement of LMS_DEC (what lms_fast.cu is bit-identical to) against the reference's double."""
    hd, _ = load_code("ref32x16_b")
    llr = awgn_llr(np.random.default_rng(9), 400, 32 * 126, 16, 32, 2.5)
    r = po.ref_decode(po.LMS, hd, 126, llr, 10, fresh=False, want_post=False)
    o = po.orc_decode(po.LMS, hd, 126, llr, 10, dtype=np.float32)
    bad = (r["iters"] != o["iters"]) | (r["hard"] != o["hard"]).any(axis=1)
    assert bad.sum() == 0
