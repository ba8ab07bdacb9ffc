"""Edge cases through the C ABI: a tiny ragged code (Z = 7 lanes in a 32-thread CTA, N = 35 not a multiple of 4 or 32,
row weight 2, weight-1 columns), a lifting size > 256, single-frame and odd batch sizes, pipelined host batches larger
than one staging chunk, puncturing, extreme LLRs."""
import numpy as np
import pytest

from codes import load_code, awgn_llr

pytestmark = pytest.mark.gpu

TINY = np.array([[0, -1, 3, 1, -1],
                 [-1, 0, 5, -1, 2]], dtype=np.int16)          # 2 x 5, Z = 7


@pytest.mark.parametrize("dec", ["LMS", "MS", "IMS", "TASP", "ASP", "BP", "SP", "LCHE", "IASP"])
def test_tiny_ragged_code_all_decoders(ldpc, po, dec):
    Z = 7
    llr = awgn_llr(np.random.default_rng(3), 50, 5 * Z, 2, 5, 1.0)
    did = getattr(po, dec)
    want = po.orc_decode(did, TINY, Z, llr, 8)
    with ldpc.Decoder(TINY, Z, did) as d:
        got = d.decode(llr, 8, want_post=True)
    bad = (got["iters"] != want["iters"]) | (got["hard"] != want["hard"]).any(axis=1)
    assert bad.sum() == 0
    if dec in ("LMS", "MS", "IMS", "IASP", "LCHE"):
        assert np.array_equal(got["post"].astype(np.float64), want["post"].astype(np.float64))


@pytest.mark.parametrize("dec,prec", [("LMS", 32), ("MS", 32), ("IMS", 64)])
@pytest.mark.parametrize("use_fast", [1, 2])
def test_tiny_ragged_code_fast_kernels(ldpc, po, dec, prec, use_fast):
    """The shared-memory / code-specialised kernels with 7 active lanes of 32, and the N % 4 != 0 tail of the fused
    BPSK first load (simulate == decode(generate_llr))."""
    Z = 7
    did = getattr(po, dec)
    llr = awgn_llr(np.random.default_rng(4), 40, 5 * Z, 2, 5, 0.5).astype(np.float32)
    want = po.orc_decode(did, TINY, Z, llr.astype(np.float64) if dec == "IMS" else llr, 8, dtype=np.float64 if dec == "IMS" else np.float32)
    with ldpc.Decoder(TINY, Z, did, precision=prec, use_fast=use_fast) as d:
        info = d.kernel_info()
        got = d.decode(llr, 8, want_post=True)
        sim = d.simulate(0.5, 64, 8, seed=9, want_per_frame=True)
        dec2 = d.decode(d.generate_llr(0.5, 64, seed=9), 8)
    if use_fast == 2:
        assert info["fast"] == 3, info
    assert np.array_equal(got["iters"], want["iters"]) and np.array_equal(got["hard"], want["hard"])
    assert np.array_equal(got["post"].astype(np.float64), want["post"].astype(np.float64))
    assert np.array_equal(sim["per_frame"] >> 31, (dec2["hard"].sum(axis=1) > 0).astype(np.uint32))


def test_large_lifting_and_batch_shapes(ldpc, po):
    """Z = 384 > 256 on the 12 x 24 matrix (shifts < 81 stay valid), batches of 1 and of an odd size."""
    hd, _ = load_code("c4_wifi_12x24")
    Z = 384
    llr = awgn_llr(np.random.default_rng(6), 37, 24 * Z, 12, 24, 1.8).astype(np.float32)
    want = po.orc_decode(po.LMS, hd, Z, llr, 10, dtype=np.float32)
    for fast in (0, 1, 2):
        with ldpc.Decoder(hd, Z, po.LMS, precision=32, use_fast=fast) as d:
            got = d.decode(llr, 10)
            one = d.decode(llr[5:6], 10)
        assert np.array_equal(got["iters"], want["iters"]) and np.array_equal(got["hard"], want["hard"]), fast
        assert one["iters"][0] == want["iters"][5] and np.array_equal(one["hard"][0], want["hard"][5])


def test_host_batch_larger_than_a_staging_chunk(ldpc, po):
    """The H2D | decode | D2H pipeline over several chunks and both slots gives the same answers as one chunk."""
    hd, _ = load_code("ref32x16_b")
    Z = 256
    nf = 9000                                            # > 192 MiB / 32 KiB = 6144 frames per chunk
    llr = awgn_llr(np.random.default_rng(8), nf, 32 * Z, 16, 32, 3.0).astype(np.float32)
    with ldpc.Decoder(hd, Z, po.LMS, precision=32) as d:
        got = d.decode(llr, 10, packed=True)
        head = d.decode(llr[:300], 10, packed=True)
        tail = d.decode(llr[-300:], 10, packed=True)
    assert np.array_equal(got["iters"][:300], head["iters"]) and np.array_equal(got["hard"][:300], head["hard"])
    assert np.array_equal(got["iters"][-300:], tail["iters"]) and np.array_equal(got["hard"][-300:], tail["hard"])
    want = po.orc_decode(po.LMS, hd, Z, llr[6100:6200], 10, dtype=np.float32)      # frames straddling the chunk boundary
    assert np.array_equal(got["iters"][6100:6200], want["iters"])


def test_extreme_and_degenerate_llrs(ldpc, po):
    hd, _ = load_code("ref32x16_b")
    Z = 126
    N = 32 * Z
    llr = np.zeros((6, N), np.float32)
    llr[0] = 1e30                                        # huge confidence, all-zero codeword
    llr[1] = -1e30                                       # huge confidence in the all-ones word (not a codeword here)
    llr[2] = 0.0                                         # erasures everywhere
    llr[3] = awgn_llr(np.random.default_rng(1), 1, N, 16, 32, 2.0)[0] * 1e4      # beyond the 32767 ceiling of the minima
    llr[4, ::2] = 5.0; llr[4, 1::2] = -5.0
    llr[5] = 40000.0
    want = po.orc_decode(po.LMS, hd, Z, llr, 6, dtype=np.float32)
    with ldpc.Decoder(hd, Z, po.LMS, precision=32) as d:
        got = d.decode(llr, 6, want_post=True)
    assert np.array_equal(got["iters"], want["iters"]) and np.array_equal(got["hard"], want["hard"])
    assert np.array_equal(got["post"], want["post"])
    wantd = po.orc_decode(po.IMS, hd, Z, llr[[0, 1, 3, 4, 5]].astype(np.float64), 6)
    with ldpc.Decoder(hd, Z, po.IMS, use_fast=2) as d:
        gotd = d.decode(llr[[0, 1, 3, 4, 5]].astype(np.float64), 6, want_aux=True)
    assert np.array_equal(gotd["iters"], wantd["iters"]) and np.array_equal(gotd["aux"], wantd["aux"])
