"""The shared-memory LMS_DEC throughput kernel (lms_fast.cu, fp32) against the oracle, through the C ABI."""
import numpy as np
import pytest

from codes import load_code, awgn_llr

pytestmark = pytest.mark.gpu


def _case(code, Z, snr, nf, seed=11):
    hd, _ = load_code(code)
    b, c = hd.shape
    return hd, awgn_llr(np.random.default_rng(seed), nf, c * Z, b, c, snr)


@pytest.mark.parametrize("code,Z,snr", [("ref32x16_b", 256, 2.0), ("ref32x16_b", 126, 2.0),
                                         ("c4_wifi_12x24", 81, 1.5), ("ref32x16_a", 126, 2.5)])
def test_fast_f32_bit_exact_vs_f32_oracle(ldpc, po, code, Z, snr):
    """Same fp32 arithmetic, different evaluation order of monotone steps: results must be bit-identical."""
    hd, llr = _case(code, Z, snr, 300)
    llr = llr.astype(np.float32)
    want = po.orc_decode(po.LMS, hd, Z, llr, 10, dtype=np.float32)
    with ldpc.Decoder(hd, Z, po.LMS, precision=32) as d:
        assert d.kernel_info()["fast"] >= 1
        got = d.decode(llr, 10, want_post=True)
    assert np.array_equal(got["iters"], want["iters"])
    assert np.array_equal(got["hard"], want["hard"])
    assert np.array_equal(got["post"], want["post"])


def test_fast_f32_vs_reference_double(ldpc, po):
    """north_star bar for float BP: identical hard decisions and iteration counts on >= 99.99 % of frames,
    posteriors within 1e-4 relative (here: of the LLR scale) against the double algorithm."""
    hd, llr = _case("ref32x16_b", 256, 2.5, 4000, seed=5)
    want = po.orc_decode(po.LMS, hd, 256, llr, 10)                    # double = the reference, bit for bit
    with ldpc.Decoder(hd, 256, po.LMS, precision=32) as d:
        got = d.decode(llr, 10, want_post=True)                      # double LLRs are rounded to fp32 on load
    bad = (got["iters"] != want["iters"]) | (got["hard"] != want["hard"]).any(axis=1)
    assert bad.mean() <= 1e-4, bad.sum()
    err = np.abs(got["post"].astype(np.float64) - want["post"])
    assert np.max(err / np.maximum(np.abs(want["post"]), 1.0)) <= 1e-4


def test_fast_no_early_exit_and_f64_input(ldpc, po):
    hd, llr = _case("ref32x16_b", 126, 3.0, 64)
    with ldpc.Decoder(hd, 126, po.LMS, precision=32) as d:
        a = d.decode(llr.astype(np.float32), 10)
        b = d.decode(llr.astype(np.float32).astype(np.float64), 10)      # same values through the F64 loader
        assert np.array_equal(a["iters"], b["iters"]) and np.array_equal(a["hard"], b["hard"])
        n = d.decode(llr.astype(np.float32), 10, no_early_exit=True)     # iters reports the first success
        assert np.array_equal(a["iters"], n["iters"])
        ok = a["iters"] > 0
        assert np.array_equal(a["hard"][ok], n["hard"][ok])              # a codeword is a fixed point


def test_simulate_matches_decode_on_generated_llr(ldpc, po):
    """simulate() (LLRs generated inside the decoder's first load) == decode(generate_llr()) + error counting."""
    hd, _ = load_code("ref32x16_b")
    with ldpc.Decoder(hd, 256, po.LMS, precision=32) as d:
        sim = d.simulate(2.0, 500, 10, seed=3, stream=2, first_frame=1000, want_per_frame=True)
        llr = d.generate_llr(2.0, 500, seed=3, stream=2, first_frame=1000)
        dec = d.decode(llr, 10)
        split = d.simulate(2.0, 200, 10, seed=3, stream=2, first_frame=1300, want_per_frame=True)
    errs = dec["hard"].sum(axis=1)
    info = dec["hard"][:, d.R:].sum(axis=1)
    assert sim["frames"] == 500
    assert sim["frame_errors"] == int((errs > 0).sum())
    assert sim["bit_errors"] == int(errs.sum())
    assert sim["info_bit_errors"] == int(info[errs > 0].sum())
    assert sim["undetected"] == int(((errs > 0) & (dec["iters"] >= 0)).sum())
    assert sim["iter_sum"] == int(np.abs(dec["iters"]).sum())
    pf = sim["per_frame"]
    assert np.array_equal(pf >> 31, (errs > 0).astype(np.uint32))
    assert np.array_equal(pf & 0xFFFFFF, info.astype(np.uint32))
    assert np.array_equal(split["per_frame"], pf[300:])                 # independent of how frames are batched
    # the oracle agrees with the device decode of the same buffer
    want = po.orc_decode(po.LMS, hd, 256, llr, 10, dtype=np.float32)
    assert np.array_equal(dec["iters"], want["iters"]) and np.array_equal(dec["hard"], want["hard"])


@pytest.mark.parametrize("code,Z,snr", [("c4_wifi_12x24", 81, 1.5), ("ref32x16_a", 126, 2.5)])
def test_runtime_compiled_kernel_bit_exact(ldpc, po, code, Z, snr):
    """use_fast = 2: the code-specialised kernel is generated and compiled (NVRTC) for a matrix that has no
    ahead-of-time instance; same bit-identical results, incl. a lifting size that is not a multiple of 32."""
    hd, llr = _case(code, Z, snr, 300)
    llr = llr.astype(np.float32)
    want = po.orc_decode(po.LMS, hd, Z, llr, 10, dtype=np.float32)
    with ldpc.Decoder(hd, Z, po.LMS, precision=32, use_fast=2) as d:
        assert d.kernel_info()["fast"] == 3, d.kernel_info()
        got = d.decode(llr, 10, want_post=True)
        sim = d.simulate(snr, 200, 10, seed=4, want_per_frame=True)
        dec = d.decode(d.generate_llr(snr, 200, seed=4), 10)
    assert np.array_equal(got["iters"], want["iters"])
    assert np.array_equal(got["hard"], want["hard"])
    assert np.array_equal(got["post"], want["post"])
    assert sim["frame_errors"] == int((dec["hard"].sum(axis=1) > 0).sum())
    with ldpc.Decoder(hd, Z, po.LMS, precision=32, use_fast=1) as d:
        assert d.kernel_info()["fast"] == 1
        assert np.array_equal(d.decode(llr, 10)["iters"], want["iters"])


def test_large_code_variant_c3_qam64(ldpc, po):
    """Config C3 (46 x 68 'BG1-shaped', Z = 384, QAM-64, two punctured block columns) on the second layout variant of
    the code-specialised kernel (single-copy posteriors, sign/position words in shared memory, row weight 19 > 16)."""
    hd, _ = load_code("c3_bg1_46x68")
    Z = 384
    with ldpc.Decoder(hd, Z, po.LMS, precision=32, use_fast=2) as d:
        assert d.kernel_info()["fast"] == 3, d.kernel_info()
        llr = np.concatenate([d.generate_llr(snr, 12, modulation=ldpc.MOD_QAM64, punct=2, seed=8) for snr in (0.0, 1.0, 1.5, 2.0, 2.5, 3.0, 4.0)])
        got = d.decode(llr, 10, want_post=True)
        sim = d.simulate(4.0, 12, 10, modulation=ldpc.MOD_QAM64, punct=2, seed=8)
    want = po.orc_decode(po.LMS, hd, Z, llr, 10, dtype=np.float32)
    assert len(set(want["iters"].tolist())) > 2, want["iters"]      # failures, slow and fast convergence
    assert np.array_equal(got["iters"], want["iters"])
    assert np.array_equal(got["hard"], want["hard"])
    assert np.array_equal(got["post"], want["post"])
    assert sim["frame_errors"] == int((got["hard"][72:].sum(axis=1) > 0).sum())
