"""The tensor-memory kernels (lms_tmem.cuh, ms_tmem.cuh, tasp_fast.cu) against their register-compressed / table-driven
twins on identical buffers: same decisions, iteration counts and posteriors bit for bit, for the ahead-of-time and the
run-time compiled instances, full and ragged lane counts; and a long run that would expose a race between the lanes
of a layer (the write-after-read hazard the split mbarrier closes)."""
import numpy as np
import pytest

from codes import load_code, awgn_llr

pytestmark = pytest.mark.gpu


def _llr(code, Z, snr, nf, seed):
    hd, _ = load_code(code)
    b, c = hd.shape
    return hd, awgn_llr(np.random.default_rng(seed), nf, c * Z, b, c, snr).astype(np.float32)


@pytest.mark.parametrize("dec,prec", [("LMS", 32), ("MS", 32), ("IMS", 64)])
@pytest.mark.parametrize("code,Z,snr", [("ref32x16_b", 256, 2.5), ("ref32x16_b", 126, 2.5), ("c4_wifi_12x24", 81, 2.0),
                                         ("c4_wifi_12x24", 384, 2.0), ("ref32x16_a", 126, 2.5)])
def test_tmem_kernel_equals_register_kernel(ldpc, po, monkeypatch, dec, prec, code, Z, snr):
    hd, llr = _llr(code, Z, snr, 300, 31)
    did = getattr(po, dec)
    monkeypatch.delenv("LDPCB200_NO_TMEM", raising=False)
    with ldpc.Decoder(hd, Z, did, precision=prec, use_fast=2) as d:
        info = d.kernel_info()
        assert info["tmem"] and info["fast"] >= 2, info
        a = d.decode(llr, 12, want_post=True, want_aux=(dec == "IMS"))
        fixed = d.decode(llr, 12, no_early_exit=True)
    monkeypatch.setenv("LDPCB200_NO_TMEM", "1")
    with ldpc.Decoder(hd, Z, did, precision=prec, use_fast=2) as d:
        info = d.kernel_info()
        assert not info["tmem"] and info["fast"] >= 2, info
        b = d.decode(llr, 12, want_post=True, want_aux=(dec == "IMS"))
    assert np.array_equal(a["iters"], b["iters"])
    assert np.array_equal(a["hard"], b["hard"])
    assert np.array_equal(a["post"], b["post"])
    if dec == "IMS":
        assert np.array_equal(a["aux"], b["aux"])
    assert np.array_equal(fixed["iters"], a["iters"])                    # fixed-iteration mode reports the first success


@pytest.mark.parametrize("dec", ["TASP", "ASP", "LCHE", "IASP", "LMS", "MS"])      # LMS, MS: the double kernels on the same skeleton
def test_sumprod_fast_equals_parity_kernel(ldpc, po, monkeypatch, dec):
    did = getattr(po, dec)
    for code, Z, snr in [("ref32x16_b", 126, 2.0), ("c4_wifi_12x24", 81, 1.5), ("ref32x16_b", 256, 2.0)]:
        hd, llr = _llr(code, Z, snr, 200, 33)
        llr = llr.astype(np.float64)
        monkeypatch.delenv("LDPCB200_NO_TASP_FAST", raising=False)
        with ldpc.Decoder(hd, Z, did) as d:
            assert d.kernel_info()["tmem"], d.kernel_info()
            a = d.decode(llr, 30, want_post=True)
            fixed = d.decode(llr, 30, no_early_exit=True)
        monkeypatch.setenv("LDPCB200_NO_TASP_FAST", "1")
        with ldpc.Decoder(hd, Z, did) as d:
            assert d.kernel_info()["fast"] == 0
            b = d.decode(llr, 30, want_post=True)
        assert np.array_equal(a["iters"], b["iters"]) and np.array_equal(a["hard"], b["hard"])
        assert np.array_equal(a["post"], b["post"])                       # same expressions, same order: bitwise
        assert np.array_equal(fixed["iters"], a["iters"])
        want = po.orc_decode(did, hd, Z, llr[:60], 30)
        assert np.array_equal(a["iters"][:60], want["iters"]) and np.array_equal(a["hard"][:60], want["hard"])


@pytest.mark.parametrize("four", [True, False])
@pytest.mark.parametrize("dec", ["BP", "SP"])
def test_bp_sp_fast_kernels_against_the_parity_kernels(ldpc, po, monkeypatch, dec, four):
    """bpsp4_kernel (bpsp4.cu: four threads per check row, the default) and bpsp_fast_kernel (tasp_fast.cu: one thread per check
    row, LDPCB200_NO_BPSP4=1): BP_DEC regroups the reference's log-domain expressions into the product form (one exp,
    one log per edge instead of two each), SP_DEC divides the edge's own message out of the column product -- the float class
    of the parity bar: identical decisions and iteration counts (here: on every frame), posteriors within 1e-4 relative
    (here: 1e-7), against the parity kernel (the reference's literal expressions) and the oracle; fixed-iteration mode
    reports the first success; a single frame; simulate == decode(generate_llr)."""
    did = getattr(po, dec)
    if four:
        monkeypatch.delenv("LDPCB200_NO_BPSP4", raising=False)
    else:
        monkeypatch.setenv("LDPCB200_NO_BPSP4", "1")
    for code, Z, snr in [("c4_wifi_12x24", 81, 2.0), ("ref32x16_b", 126, 2.5), ("c4_wifi_12x24", 81, 6.0)]:
        hd, llr = _llr(code, Z, snr, 300, 35)
        llr = llr.astype(np.float64)
        monkeypatch.delenv("LDPCB200_NO_TASP_FAST", raising=False)
        with ldpc.Decoder(hd, Z, did) as d:
            assert d.kernel_info()["tmem"], d.kernel_info()
            assert (d.kernel_info()["threads"] > (Z + 31) // 32 * 32) == four, d.kernel_info()
            a = d.decode(llr, 20, want_post=True)
            fixed = d.decode(llr, 20, no_early_exit=True)
            one = d.decode(llr[7:8], 20)
            sim = d.simulate(snr, 200, 20, seed=4, want_per_frame=True)
            again = d.decode(d.generate_llr(snr, 200, seed=4, dtype=np.float64), 20)
        monkeypatch.setenv("LDPCB200_NO_TASP_FAST", "1")
        with ldpc.Decoder(hd, Z, did) as d:
            assert d.kernel_info()["fast"] == 0
            b = d.decode(llr, 20, want_post=True)
        assert np.array_equal(a["iters"], b["iters"]) and np.array_equal(a["hard"], b["hard"])
        rel = np.abs(a["post"] - b["post"]) / np.maximum(np.abs(b["post"]), 1e-300)
        assert np.max(rel) <= 1e-7, np.max(rel)
        assert np.array_equal(fixed["iters"], a["iters"]) and np.array_equal(one["iters"], a["iters"][7:8])
        want = po.orc_decode(did, hd, Z, llr[:80], 20)
        assert np.array_equal(a["iters"][:80], want["iters"]) and np.array_equal(a["hard"][:80], want["hard"])
        assert np.array_equal(sim["per_frame"] >> 31, (again["hard"].sum(axis=1) > 0).astype(np.uint32))
        assert sim["iter_sum"] == int(np.abs(again["iters"]).sum())


def test_tmem_lms_long_run_against_oracle(ldpc, po):
    """6 000 frames at the waterfall: any ordering bug between the lanes of a layer shows up as a handful of frames."""
    hd, llr = _llr("ref32x16_b", 256, 2.2, 6000, 35)
    with ldpc.Decoder(hd, 256, po.LMS, precision=32) as d:
        assert d.kernel_info()["tmem"]
        got = d.decode(llr, 10, packed=True)
    want = po.orc_decode(po.LMS, hd, 256, llr, 10, dtype=np.float32)
    assert np.array_equal(got["iters"], want["iters"])
    assert np.array_equal(got["hard"], np.packbits(want["hard"], axis=1, bitorder="little").view(np.uint32)[:, :got["hard"].shape[1]])


@pytest.mark.parametrize("code,Z,snr,maxiter", [("ref32x16_b", 256, 2.2, 10), ("c4_wifi_12x24", 128, 2.0, 7)])
def test_two_frames_per_cta_kernel_equals_the_one_frame_kernel(ldpc, po, monkeypatch, code, Z, snr, maxiter):
    """lms_tmem2.cuh (opt-in, LDPCB200_TMEM2=1: two frames per CTA in lock step, packed f32x2 arithmetic) against lms_tmem.cuh:
    bitwise equal posteriors, decisions and iteration counts -- the slots finish at different iterations, take new frames
    while the neighbour carries on, and an odd frame count leaves one slot empty at the end."""
    hd, llr = _llr(code, Z, snr, 1201, 37)
    monkeypatch.delenv("LDPCB200_TMEM2", raising=False)
    with ldpc.Decoder(hd, Z, po.LMS, precision=32, use_fast=2) as d:
        assert d.kernel_info()["tmem"] and not d.kernel_info()["two_frames"]
        a = d.decode(llr, maxiter, want_post=True)
    monkeypatch.setenv("LDPCB200_TMEM2", "1")
    with ldpc.Decoder(hd, Z, po.LMS, precision=32, use_fast=2) as d:
        assert d.kernel_info()["two_frames"], d.kernel_info()
        b = d.decode(llr, maxiter, want_post=True)
        fixed = d.decode(llr, maxiter, no_early_exit=True)
        sim = d.simulate(snr, 999, maxiter, seed=3)
    assert np.array_equal(a["iters"], b["iters"]) and np.array_equal(a["hard"], b["hard"])
    assert np.array_equal(a["post"].view(np.uint32), b["post"].view(np.uint32))
    assert np.array_equal(fixed["iters"], a["iters"])
    monkeypatch.delenv("LDPCB200_TMEM2", raising=False)
    with ldpc.Decoder(hd, Z, po.LMS, precision=32, use_fast=2) as d:
        sim1 = d.simulate(snr, 999, maxiter, seed=3)
    assert sim == sim1                                                   # the fused channel + counters path


@pytest.mark.parametrize("code,Z,snr,maxiter", [("ref32x16_b", 126, 2.0, 15), ("c4_wifi_12x24", 81, 2.0, 20), ("ref32x16_b", 256, 2.2, 10)])
@pytest.mark.parametrize("ctas", ["1", "2", ""])
def test_race_guard_resident_ctas_and_ragged_lifting(ldpc, po, monkeypatch, code, Z, snr, maxiter, ctas):
    """The write-after-read hazard between the lanes of a block row (split mbarrier, lms_tmem.cuh) under different timings:
    one, two or all resident CTAs per SM (LDPCB200_GRID_PER_SM) and lifting sizes that are not a multiple of 32 (C1: 126,
    C4: 81 -- the doubled-column path with idle lanes), 3 000 frames each against the fp32 oracle, bit for bit."""
    if ctas:
        monkeypatch.setenv("LDPCB200_GRID_PER_SM", ctas)
    else:
        monkeypatch.delenv("LDPCB200_GRID_PER_SM", raising=False)
    hd, llr = _llr(code, Z, snr, 3000, 41)
    with ldpc.Decoder(hd, Z, po.LMS, precision=32, use_fast=2) as d:
        assert d.kernel_info()["tmem"], d.kernel_info()
        got = d.decode(llr, maxiter, want_post=True)
    want = po.orc_decode(po.LMS, hd, Z, llr, maxiter, dtype=np.float32)
    assert np.array_equal(got["iters"], want["iters"]) and np.array_equal(got["hard"], want["hard"])
    assert np.array_equal(got["post"].view(np.uint32), want["post"].view(np.uint32))


def test_c4_posterior_tolerance_of_the_fp32_kernel_against_double(ldpc, po):
    """BASELINE's bar for the float class at C4 (N = 1944, 20 iterations): identical decisions and iteration counts on
    >= 99.99 % of frames against the double arithmetic of the reference, and posteriors within 1e-4 -- evaluated as
    |d| / max(|LLR|, 1), the only meaningful form for fp32 sums that pass through zero: at most 1e-4 of the
    values of agreeing frames may exceed it (fp32 accumulation over 20 layered iterations), none beyond 1e-2 (observed: 3e-3);
    LDPCB200_PRECISION=64 selects the double kernel, whose posteriors are the reference's bit for bit."""
    hd, llr = _llr("c4_wifi_12x24", 81, 2.0, 20000, 43)
    with ldpc.Decoder(hd, 81, po.LMS, precision=32, use_fast=2) as d:
        got = d.decode(llr, 20, want_post=True)
    want = po.orc_decode(po.LMS, hd, 81, llr.astype(np.float64), 20)
    bad = (got["iters"] != want["iters"]) | (got["hard"] != want["hard"]).any(axis=1)
    assert bad.mean() <= 1e-4, int(bad.sum())
    rel = np.abs(got["post"][~bad].astype(np.float64) - want["post"][~bad]) / np.maximum(np.abs(want["post"][~bad]), 1.0)
    assert (rel > 1e-4).mean() <= 1e-4, float((rel > 1e-4).mean())
    assert rel.max() <= 1e-2, float(rel.max())
