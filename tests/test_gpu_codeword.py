"""Real codewords through the channel (second half): ldpcb200_set_codeword sends a non-zero codeword
through the direct permutation, the Gray map / BPSK map, the noise, Demodulate, the inverse permutation and the puncturing,
and counts errors against it.  The reference's own loop always sends the all-zero word (bp_simulation.cpp:567), so the check
is built from its parts: the codewords are the reference encoder's (tests/golden/encoder.npz), the modulator and
demodulator are the oracle's restatements (pinned to the reference by tests/test_oracle_golden.py), the permutation tables
are the reference's, and the noise samples are read back from the generator."""
import os

import numpy as np
import pytest

from codes import load_code

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = np.load(os.path.join(ROOT, "tests", "golden", "encoder.npz"))


def codeword(name, Z, seed=1):
    return np.unpackbits(G["%s_Z%d_s%d_cw" % (name, Z, seed)])[:32 * Z].astype(np.uint8)


def test_bpsk_llr_of_a_codeword(ldpc, po):
    """-2 (sigma n + 2 c - 1) / sigma^2 (bp_simulation.cpp:600-612): a one moves the zero-codeword LLR by -4 / sigma^2."""
    hd, _ = load_code("ref32x16_b")
    Z, snr = 126, 2.0
    cw = codeword("ref32x16_b", Z)
    direct, inverse = ldpc.interleaver_tables(hd, Z, 0, 1)
    sig = ldpc.sigma(16, 32, 0, snr, 0)
    for perm in (False, True):
        with ldpc.Decoder(hd, Z, po.LMS, precision=32) as d:
            if perm:
                d.set_interleaver(direct, inverse)
            zero = d.generate_llr(snr, 4, seed=5, dtype=np.float64)
            d.set_codeword(cw)
            got = d.generate_llr(snr, 4, seed=5, dtype=np.float64)
            d.set_codeword(None)
            back = d.generate_llr(snr, 4, seed=5, dtype=np.float64)
        want = zero - 4.0 * cw[None, :] / sig ** 2
        assert np.allclose(got, want, rtol=2e-6, atol=2e-5)
        assert np.array_equal(back, zero)
        assert ((got < 0) == (cw[None, :] == 1)).mean() > 0.85           # the signs follow the codeword


@pytest.mark.parametrize("mod,mode", [(2, 0), (3, 1), (4, 3), (3, 2)])
def test_qam_chain_of_a_codeword_against_modulator_and_demodulator(ldpc, po, mod, mode):
    hd, _ = load_code("ref32x16_a")
    Z, snr, nf = 126, {2: 6.0, 3: 10.0, 4: 14.0}[mod], 3
    N, m, Q = 32 * Z, 2 * mod, 1 << (2 * mod)
    ns = N // m
    cw = codeword("ref32x16_a", Z, 5)
    direct, inverse = ldpc.interleaver_tables(hd, Z, mod, mode, 96, 1)
    sig = ldpc.sigma(16, 32, 2, snr, mod)
    with ldpc.Decoder(hd, Z, po.LMS, precision=32) as d:
        d.set_interleaver(direct, inverse)
        d.set_codeword(cw)
        got = d.generate_llr(snr, nf, modulation=mod, punct=2, seed=9, dtype=np.float64)
        noise = d.generate_noise(snr, nf, 2 * ns, modulation=mod, punct=2, seed=9).astype(np.float64)
    sent = cw[direct]                                                # Permutation direction 0, bp_simulation.cpp:573
    x = po.orc_modulate(Q, sent)                                     # QAM_modulator, :577
    for f in range(nf):
        res = po.orc_demodulate(Q, ns, sig, x + sig * noise[f])      # r = s + sigmaQAM n; Demodulate, :626
        y = (-res)[inverse]                                          # :627-628, then Permutation direction 1, :684
        y[N - 2 * Z:] = 0.5                                          # :697-710
        assert np.allclose(got[f], y.astype(np.float32).astype(np.float64), rtol=3e-6, atol=1e-6), (mod, mode, f)
    assert ((got[:, :N - 2 * Z] < 0) == (cw[None, :N - 2 * Z] == 1)).mean() > 0.8


@pytest.mark.parametrize("dec,prec,mod,snr", [("LMS", 32, 0, 2.5), ("LMS", 32, 3, 9.5), ("IMS", 64, 0, 3.5), ("TASP", 64, 2, 5.5), ("BP", 64, 0, 2.5)])
def test_simulate_counts_errors_against_the_codeword(ldpc, po, dec, prec, mod, snr):
    hd, _ = load_code("ref32x16_a")
    Z, nf, it = 126, 120, 12
    N, R = 32 * Z, 16 * Z
    cw = codeword("ref32x16_a", Z)
    direct, inverse = ldpc.interleaver_tables(hd, Z, mod, 1)
    with ldpc.Decoder(hd, Z, getattr(po, dec), precision=prec) as d:
        d.set_interleaver(direct, inverse)
        d.set_codeword(cw)
        sim = d.simulate(snr, nf, it, modulation=mod, seed=21, want_per_frame=True)
        llr = d.generate_llr(snr, nf, modulation=mod, seed=21, dtype=np.float32)
        ref = d.decode(llr, it)
        clean = d.simulate(snr + 8.0, 50, it, modulation=mod, seed=22)
        d.set_codeword(None)
        zero = d.simulate(snr, 400, it, modulation=mod, seed=21)
    wrong = ref["hard"] != cw[None, :]
    err, info = wrong.sum(axis=1), wrong[:, R:].sum(axis=1)
    assert np.array_equal(sim["per_frame"] >> 31, (err > 0).astype(np.uint32))
    assert np.array_equal(sim["per_frame"] & 0xFFFFFF, np.where(err > 0, info, 0).astype(np.uint32))
    assert sim["bit_errors"] == int(err.sum()) and sim["frames"] == nf and sim["iter_sum"] == int(np.abs(ref["iters"]).sum())
    assert clean["frame_errors"] == 0 and clean["frames"] == 50      # far above the waterfall: every frame is the codeword
    assert zero["frames"] == 400
