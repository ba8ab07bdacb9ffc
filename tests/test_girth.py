"""Girth / ACE / cycle spectrum (ldpcb200_girth_spectrum, csrc/girth.cpp) against the reference's trace_bound_pol_mon_pm as its
driver calls it (main_simulation.cpp:148-205): golden values recorded from the compiled reference, and a live comparison on
random base matrices when oracle/_ref is built."""
import numpy as np
import pytest

from codes import load_code

# (code, Z) -> (girth, ACE[4], spectrum[4]) from oracle/_ref (pyoracle.ref_girth), 2026-10
GOLDEN = {
    ("ref32x16_b", 126): (6, [16, 14, 17, 15], [93, 2828, 90894, 3109874]),
    ("ref32x16_a", 126): (6, [17, 15, 17, 17], [58, 1856, 52692, 1616777]),
    ("c4_wifi_12x24", 81): (6, [16, 18, 14, 15], [41, 871, 18091, 410773]),
    ("ref32x16_b", 256): (6, [16, 14, 18, 19], [53, 1453, 45857, 1542873]),
}


@pytest.mark.parametrize("code,Z", sorted(GOLDEN))
def test_girth_spectrum_matches_the_reference_golden(ldpc, code, Z):
    hd, _ = load_code(code)
    assert ldpc.girth_spectrum(hd, Z) == GOLDEN[(code, Z)]


def test_girth_spectrum_matches_the_compiled_reference_on_random_matrices(ldpc, po):
    if not po.have_ref():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(5)
    for b, c, Z, dens in [(4, 8, 16, 0.6), (5, 10, 31, 0.5), (6, 12, 64, 0.4), (3, 9, 7, 0.9)]:
        hd = np.where(rng.random((b, c)) < dens, rng.integers(0, Z, (b, c)), -1)
        for i in range(c):                                   # every column needs two entries (the reference reads f[0], f[1])
            rows = rng.choice(b, 2, replace=False)
            for r in rows:
                if hd[r, i] < 0:
                    hd[r, i] = rng.integers(0, Z)
        # the reference's edge model takes exactly the first two entries of a column of its split graph: any weight works
        assert ldpc.girth_spectrum(hd, Z) == po.ref_girth(hd, Z), (b, c, Z)


def test_girth_spectrum_rejects_bad_arguments(ldpc):
    with pytest.raises(ldpc.LdpcError):
        ldpc.girth_spectrum(np.zeros((2, 2), np.int16), 0)
