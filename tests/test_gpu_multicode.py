"""ldpcb200_simulate_codes: many candidate matrices of one shape scored in ONE launch -- blockIdx.y of
the table-driven TASP_DEC kernel is the code -- against one ldpcb200_simulate call per candidate on its own handle: the same
counters and per-frame records, bit for bit (each code sees the same noise, as the reference's reset_random() before every
candidate gives it, main_good_code_search.cpp:316)."""
import numpy as np
import pytest

from codes import load_code

pytestmark = pytest.mark.gpu


def candidates(hd, Z, K, seed):
    """K matrices of hd's shape and number of circulants: random shifts on hd's mask, and every third one with a
    circulant moved to another column of its row (a different mask and other column weights, the same E)."""
    rng = np.random.default_rng(seed)
    out = []
    for k in range(K):
        m = hd.copy()
        nz = m >= 0
        m[nz] = rng.integers(0, Z, int(nz.sum()))
        if k % 3 == 2:
            for _ in range(3):
                j = int(rng.integers(0, m.shape[0]))
                a = rng.choice(np.flatnonzero(m[j] >= 0))
                b = rng.choice(np.flatnonzero(m[j] < 0))
                if (m[:, a] >= 0).sum() > 1:                       # keep every column connected
                    m[j, b], m[j, a] = m[j, a], -1
        out.append(m)
    return np.stack(out)


@pytest.mark.parametrize("code,Z,snr,K", [("c4_wifi_12x24", 81, 2.0, 7), ("ref32x16_b", 126, 2.5, 4)])
def test_batched_codes_equal_single_calls(ldpc, po, code, Z, snr, K):
    hd, _ = load_code(code)
    hds = candidates(hd, Z, K, 5)
    nf, it = 160, 12
    with ldpc.Decoder(hd, Z, po.TASP) as d:
        got = d.simulate_codes(hds, snr, nf, it, seed=3, stream=2, first_frame=1000, want_per_frame=True)
        again = d.simulate_codes(hds[::-1].copy(), snr, nf, it, seed=3, stream=2, first_frame=1000)      # the handle is reused
        fixed = d.simulate_codes(hds[:2], snr, nf, it, seed=3, stream=2, first_frame=1000, no_early_exit=True)
    assert len(got) == K
    for k in range(K):
        with ldpc.Decoder(hds[k], Z, po.TASP) as s:
            want = s.simulate(snr, nf, it, seed=3, stream=2, first_frame=1000, want_per_frame=True)
            wfix = s.simulate(snr, nf, it, seed=3, stream=2, first_frame=1000, no_early_exit=True) if k < 2 else None
        assert np.array_equal(got[k].pop("per_frame"), want.pop("per_frame")), k
        assert got[k] == want, (k, got[k], want)
        assert again[K - 1 - k] == want
        if wfix is not None:
            assert fixed[k] == wfix
    assert len({g["frame_errors"] for g in got}) > 1                 # the candidates do differ


def test_batched_codes_refusals(ldpc, po):
    hd, _ = load_code("c4_wifi_12x24")
    fewer = hd.copy()
    fewer[0, np.flatnonzero(fewer[0] >= 0)[0]] = -1                  # one circulant less: another E
    with ldpc.Decoder(hd, 81, po.TASP) as d:
        with pytest.raises(ldpc.LdpcError):
            d.simulate_codes(np.stack([hd, fewer]), 2.0, 10, 5)
        assert d.simulate_codes(np.stack([hd]), 2.0, 0, 5)[0]["frames"] == 0
    with ldpc.Decoder(hd, 81, po.LMS, precision=32) as d:
        with pytest.raises(ldpc.LdpcError):
            d.simulate_codes(np.stack([hd]), 2.0, 10, 5)
