"""Bit interleaver tables (ldpcb200_interleaver_tables, csrc/interleaver.cpp) against the reference's
(direct_inverse_perm.cpp via oracle/_ref): committed golden tables (tests/golden/interleaver.npz, made by
tests/golden/make_golden.py interleaver) and, when the compiled reference is there, a live sweep over modes and modulations."""
import os

import numpy as np
import pytest

from codes import load_code

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_interleaver_tables_match_the_reference_golden(ldpc):
    G = np.load(os.path.join(ROOT, "tests", "golden", "interleaver.npz"))
    for k, (code, (Z, mod, mode, block, inter)) in enumerate(zip(G["codes"], G["cases"])):
        hd, _ = load_code(str(code))
        d, i = ldpc.interleaver_tables(hd, int(Z), int(mod), int(mode), int(block), int(inter))
        assert np.array_equal(d, G["direct_%d" % k]) and np.array_equal(i, G["inverse_%d" % k]), (code, Z, mod, mode)
        assert np.array_equal(d[i], np.arange(d.size))                  # decoder input i <- position inverse[i] carries bit i


def test_interleaver_tables_match_the_compiled_reference(ldpc, po):
    if not po.have_ref():
        pytest.skip("oracle/_ref not built")
    for code, Z in [("ref32x16_b", 42), ("c3_bg1_46x68", 24), ("c4_wifi_12x24", 81)]:
        hd, _ = load_code(code)
        for mod in range(5):
            for mode, block, inter in [(0, 1, 1), (1, 1, 1), (2, 1, 1), (3, 96, 1), (3, 100, 1), (4, 1, 12), (4, 1, 7)]:
                rd, ri = po.ref_permutation(hd, Z, mod, mode, block, inter)
                n = rd.size
                ref_is_perm = np.array_equal(np.sort(rd), np.arange(n)) and np.array_equal(np.sort(ri), np.arange(n))
                try:
                    d, i = ldpc.interleaver_tables(hd, Z, mod, mode, block, inter)
                except ldpc.LdpcError:
                    assert not ref_is_perm, (code, Z, mod, mode)        # refused only where the reference reads stale memory
                    continue
                assert np.array_equal(d, rd) and np.array_equal(i, ri), (code, Z, mod, mode, block, inter)
