"""Random base matrices through every throughput kernel against the oracle: the generators of the code-specialised
kernels (rotation tables, TMEM column packing, row-weight chunks) and the table-driven tensor-memory kernels must hold
for shapes nobody hand-picked -- lifting sizes that are not multiples of 32, weight-1 columns, rows of weight 2 and of
weight > 16, all-zero shifts, one warp per frame and twelve."""
import numpy as np
import pytest

from codes import awgn_llr

pytestmark = pytest.mark.gpu


def random_code(seed):
    rng = np.random.default_rng(1000 + seed)
    b = int(rng.integers(2, 11))
    c = int(rng.integers(b + 2, 2 * b + 6))
    Z = int(rng.choice([17, 32, 40, 64, 81, 96, 127, 200, 256, 330]))
    hd = np.full((b, c), -1, np.int16)
    for j in range(b):                                   # row weights 2 .. min(c, 18)
        w = int(rng.integers(2, min(c, 18) + 1))
        cols = rng.choice(c, size=w, replace=False)
        hd[j, cols] = rng.integers(0, Z, size=w)
    for i in range(c):                                   # every column takes part in at least one check
        if (hd[:, i] == -1).all():
            hd[int(rng.integers(0, b)), i] = int(rng.integers(0, Z))
    if seed % 4 == 0:
        hd[hd != -1] = 0                                 # all shifts zero: every rotation delta is zero
    return hd, Z


SEEDS = list(range(8))


@pytest.mark.parametrize("seed", SEEDS)
@pytest.mark.parametrize("dec,prec", [("LMS", 32), ("MS", 32), ("IMS", 64)])
def test_code_specialised_kernels_on_random_codes(ldpc, po, seed, dec, prec):
    hd, Z = random_code(seed)
    b, c = hd.shape
    llr = awgn_llr(np.random.default_rng(seed), 48, c * Z, b, c, 1.0).astype(np.float32)
    did = getattr(po, dec)
    want = po.orc_decode(did, hd, Z, llr.astype(np.float64) if dec == "IMS" else llr, 8, dtype=np.float64 if dec == "IMS" else np.float32)
    with ldpc.Decoder(hd, Z, did, precision=prec, use_fast=2) as d:
        info = d.kernel_info()
        # (normally a run-time compiled instance, tensor-memory variant when the messages fit; a table-driven kernel otherwise)
        got = d.decode(llr, 8, want_post=True)
        sim = d.simulate(1.0, 96, 8, seed=seed, want_per_frame=True)
        again = d.decode(d.generate_llr(1.0, 96, seed=seed), 8)
    assert np.array_equal(got["iters"], want["iters"]), (info, hd.shape, Z)
    assert np.array_equal(got["hard"], want["hard"])
    assert np.array_equal(got["post"].astype(np.float64), want["post"].astype(np.float64))
    assert np.array_equal(sim["per_frame"] >> 31, (again["hard"].sum(axis=1) > 0).astype(np.uint32))


@pytest.mark.parametrize("seed", SEEDS)
@pytest.mark.parametrize("dec", ["TASP", "ASP", "LCHE", "IASP", "MS"])
def test_table_driven_tmem_kernels_on_random_codes(ldpc, po, seed, dec):
    hd, Z = random_code(seed)
    b, c = hd.shape
    llr = awgn_llr(np.random.default_rng(seed), 48, c * Z, b, c, 1.0)
    did = getattr(po, dec)
    want = po.orc_decode(did, hd, Z, llr, 8)
    with ldpc.Decoder(hd, Z, did) as d:
        info = d.kernel_info()
        got = d.decode(llr, 8, want_post=True)
    bad = (got["iters"] != want["iters"]) | (got["hard"] != want["hard"]).any(axis=1)
    assert bad.sum() == 0, (info, hd.shape, Z, int(bad.sum()))
    if dec in ("LCHE", "IASP", "MS"):
        assert np.array_equal(got["post"].astype(np.float64), want["post"].astype(np.float64))
