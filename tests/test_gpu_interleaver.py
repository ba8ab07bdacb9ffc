"""Bit interleaver on the device: ldpcb200_set_interleaver makes the decoder's first load a gather /
scatter through the reference's permutation tables (direct_inverse_perm.cpp; the tables are the compiled reference's own,
tests/golden/interleaver.npz).  Checked: the de-interleaved LLR buffer is what bp_simulation.cpp:684 + :697-710 produce from the
transmitted-order buffer, and every kernel family's fused first load decodes exactly that buffer."""
import os

import numpy as np
import pytest

from codes import load_code

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def golden_tables():
    G = np.load(os.path.join(ROOT, "tests", "golden", "interleaver.npz"))
    for k, (code, (Z, mod, mode, block, inter)) in enumerate(zip(G["codes"], G["cases"])):
        yield str(code), int(Z), int(mod), int(mode), G["direct_%d" % k], G["inverse_%d" % k]


def test_deinterleaved_llrs_equal_the_reference_permutation_of_the_transmitted_ones(ldpc, po):
    """generate_llr with an interleaver == Permutation(direction 1) of generate_llr without, then puncturing."""
    for code, Z, mod, mode, direct, inverse in golden_tables():
        hd, _ = load_code(code)
        N = hd.shape[1] * Z
        if mod >= 2 and N % (2 * mod):
            continue
        for punct in (0, 2):
            ps = N - punct * Z
            with ldpc.Decoder(hd, Z, po.LMS, precision=32, use_fast=1) as d:
                # transmitted order; sigma depends on the punctured blocks (bp_simulation.cpp:444), so the same `punct`, whose
                # own puncturing overwrites the last positions of THIS buffer: those are left out of the comparison below
                sent = d.generate_llr(2.0, 5, modulation=mod, punct=punct, seed=7, dtype=np.float64)
                d.set_interleaver(direct, inverse)
                got = d.generate_llr(2.0, 5, modulation=mod, punct=punct, seed=7, dtype=np.float64)
                d.set_interleaver()
                back = d.generate_llr(2.0, 5, modulation=mod, punct=punct, seed=7, dtype=np.float64)
            known = np.flatnonzero((inverse < ps) & (np.arange(N) < ps))     # y[i] = buffer[inverse[i]], bp_simulation.cpp:684
            assert np.array_equal(got[:, known], sent[:, inverse[known]]), (code, Z, mod, mode, punct)
            assert punct == 0 or np.all(got[:, ps:] == 0.5)                  # :697-710, after the inverse permutation
            assert np.array_equal(back, sent)
            if mode and mod >= 2:
                assert not np.array_equal(got[:, :ps], sent[:, :ps])


FAMILIES = [("LMS", 32, 1, "ref32x16_b", 126, 3, {}),             # lms_tmem (AOT), QAM-64
            ("LMS", 32, 1, "ref32x16_b", 126, 0, {}),             # lms_tmem, BPSK
            ("LMS", 32, 1, "ref32x16_b", 126, 2, {"LDPCB200_NO_TMEM": "1"}),   # lms_spec
            ("LMS", 32, 1, "c4_wifi_12x24", 27, 3, {}),           # lms_fast_kernel (table-driven)
            ("LMS", 32, 2, "c4_wifi_12x24", 27, 2, {}),           # NVRTC instance
            ("MS", 32, 1, "ref32x16_b", 126, 3, {}),              # ms_tmem<float>
            ("IMS", 64, 1, "ref32x16_b", 126, 2, {}),             # ms_tmem<int>
            ("IMS", 64, 1, "ref32x16_b", 126, 0, {"LDPCB200_NO_TMEM": "1"}),   # ms_spec<int>
            ("TASP", 64, 1, "c4_wifi_12x24", 27, 3, {}),          # tasp_fast (per-bit gather)
            ("BP", 64, 1, "c4_wifi_12x24", 27, 2, {})]            # table-driven parity kernel


@pytest.mark.parametrize("dec,prec,fast,code,Z,mod,env", FAMILIES)
def test_fused_first_load_with_an_interleaver_equals_decoding_the_generated_buffer(ldpc, po, monkeypatch, dec, prec, fast, code, Z, mod, env):
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    hd, _ = load_code(code)
    N = hd.shape[1] * Z
    direct, inverse = ldpc.interleaver_tables(hd, Z, mod, 1)        # one random permutation of all N positions
    nf, it, snr = 96, 8, {0: 1.5, 2: 5.0, 3: 9.0}[mod]
    with ldpc.Decoder(hd, Z, getattr(po, dec), precision=prec, use_fast=fast) as d:
        d.set_interleaver(direct, inverse)
        sim = d.simulate(snr, nf, it, modulation=mod, punct=1, seed=11, want_per_frame=True)
        llr = d.generate_llr(snr, nf, modulation=mod, punct=1, seed=11, dtype=np.float32)
        ref = d.decode(llr, it)
    err = ref["hard"].sum(axis=1)
    info_err = ref["hard"][:, hd.shape[0] * Z:].sum(axis=1)
    assert np.array_equal(sim["per_frame"] >> 31, (err > 0).astype(np.uint32))
    assert np.array_equal(sim["per_frame"] & 0xFFFFFF, np.where(err > 0, info_err, 0).astype(np.uint32))
    assert sim["bit_errors"] == int(err.sum()) and sim["frames"] == nf


def test_set_interleaver_rejects_tables_that_are_not_inverse_permutations(ldpc, po):
    hd, _ = load_code("c4_wifi_12x24")
    with ldpc.Decoder(hd, 27, po.LMS, precision=32) as d:
        n = d.N
        with pytest.raises(ldpc.LdpcError):
            d.set_interleaver(np.arange(n), np.roll(np.arange(n), 1))
        with pytest.raises(ldpc.LdpcError):
            d.set_interleaver(np.zeros(n), np.zeros(n))
