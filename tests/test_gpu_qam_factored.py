"""The factored QAM demodulator of the in-kernel channel (channel.cuh pam_demod_factored: one exponential and one division per PAM
component) against the reference's evaluation order (pam_demod, LDPCB200_QAM_EXACT=1): the doubles agree to ~1e-15, the fp32 LLRs
the decoders see are the same values (a last-bit difference about once in 1e7), so whole simulations return the same per-frame
records -- at operating points where the decoders work hard, for QAM-16 / 64 / 256, fp32 and double decoders, with puncturing --
and simulate() still equals decode(generate_llr()) (generate_llr keeps pam_demod)."""
import numpy as np
import pytest

from codes import load_code

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dec,prec", [("LMS", 32), ("TASP", 64)])
@pytest.mark.parametrize("mod,snr", [(2, 5.0), (3, 8.0), (4, 11.5)])
def test_factored_demodulator_gives_the_same_frames(ldpc, po, monkeypatch, dec, prec, mod, snr):
    hd, _ = load_code("c4_wifi_12x24")
    runs = {}
    for exact in ("1", "0"):
        monkeypatch.setenv("LDPCB200_QAM_EXACT", exact)
        with ldpc.Decoder(hd, 81, getattr(po, dec), precision=prec) as d:
            runs[exact] = d.simulate(snr, 3000, 20, modulation=mod, punct=1, seed=9, want_per_frame=True)
            if exact == "0":
                llr = d.generate_llr(snr, 3000, modulation=mod, punct=1, seed=9)
                out = d.decode(llr, 20)
    a, b = runs["1"], runs["0"]
    assert 1.2 < a["iter_sum"] / a["frames"] < 19.0, a                           # the decoders iterate: LLR differences would show
    # 5.8e6 LLRs per run: at most one of them may differ in its last fp32 bit, so at most one frame may tell
    assert int((a["per_frame"] != b["per_frame"]).sum()) <= 1
    assert a["frames"] == b["frames"] and abs(a["frame_errors"] - b["frame_errors"]) <= 1 and abs(a["iter_sum"] - b["iter_sum"]) <= 20
    errs = out["hard"].sum(axis=1)
    assert abs(b["frame_errors"] - int((errs > 0).sum())) <= 1 and abs(b["iter_sum"] - int(np.abs(out["iters"]).sum())) <= 20


def test_factored_demodulator_is_off_where_its_powers_would_overflow(ldpc, po, monkeypatch):
    """QAM-256 at 45 dB: W^15 = exp(15 * 2 x / N0) is far outside the double range; the host must keep the reference order."""
    hd, _ = load_code("c4_wifi_12x24")
    runs = []
    for exact in ("1", "0"):
        monkeypatch.setenv("LDPCB200_QAM_EXACT", exact)
        with ldpc.Decoder(hd, 81, po.LMS, precision=32) as d:
            runs.append(d.simulate(45.0, 500, 20, modulation=4, seed=3, want_per_frame=True))
    a, b = runs
    assert b["frames"] == 500 and b["frame_errors"] == 0 and b["bit_errors"] == 0
    assert a["iter_sum"] == b["iter_sum"] and np.array_equal(a["per_frame"], b["per_frame"])
