#!/usr/bin/env python
"""Writes the frozen base matrices of the benchmark configurations (BASELINE.json:configs)
as `.jsonx` records in the reference's own format (`code = matrix (b c) { ... }`, settings.cpp:236-262).

  ref32x16_a.jsonx   REF-32x16-A    : C5 (IMS_DEC, Z=126)
  ref32x16_b.jsonx   REF-32x16-B   : C1 (Z=126) and C2 (Z=256)
  c4_wifi_12x24.jsonx  802.11n-shaped rate-1/2, Z=81, columns [parity | info]
  c3_bg1_46x68.jsonx   5G-NR-BG1-shaped, Z=384, seeded synthetic, columns [46 parity | 22 info]

The two 16x32 matrices were emitted by the reference's own `search` (seed 1) and are kept as literals;
the other two are synthetic.  Run from the repo root:  python configs/make_configs.py
"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

REF_A = """
0 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 0 -1 0 -1 -1 0 0 -1 0 -1 -1 -1 -1 -1 -1 -1
-1 0 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 0 -1 -1 -1 -1 68 -1 120 0 0 0 -1 -1 -1 -1
-1 -1 0 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 0 -1 -1 -1 0 -1 73 -1 43 106 35 -1 -1 -1 -1 -1
-1 -1 -1 0 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 99 -1 116 71 102 82 -1 -1 -1 -1
-1 -1 -1 -1 0 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 102 -1 -1 0 -1 86 100 2 -1 -1 -1 -1
-1 -1 -1 -1 -1 0 -1 -1 -1 -1 -1 -1 -1 -1 -1 0 -1 42 -1 0 -1 -1 -1 -1 -1 -1 -1 -1 0 0 -1 -1
-1 -1 -1 -1 -1 0 0 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 97 28 60 0 0
-1 -1 -1 -1 -1 -1 0 0 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 91 -1 -1 -1 -1 74 -1 -1 -1 -1 96 45 6 -1
-1 -1 -1 -1 -1 -1 -1 0 0 -1 -1 -1 -1 -1 -1 -1 121 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 39 123 -1 47
-1 -1 -1 -1 -1 -1 -1 -1 0 0 -1 -1 -1 -1 -1 -1 -1 -1 -1 9 -1 -1 73 34 114 52 12 122 71 113 56 19
-1 -1 -1 -1 -1 -1 -1 -1 -1 0 0 -1 -1 -1 -1 125 -1 -1 -1 -1 99 -1 -1 -1 -1 -1 -1 -1 -1 -1 58 13
-1 -1 -1 -1 -1 -1 -1 -1 -1 -1 0 0 -1 -1 -1 -1 -1 18 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 77 100 5
-1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 0 0 -1 -1 -1 -1 -1 113 40 -1 -1 -1 8 -1 -1 -1 -1 -1 -1 114 68
-1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 0 0 -1 -1 -1 -1 -1 -1 -1 -1 -1 10 -1 -1 -1 -1 -1 73 105 65
-1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 0 0 -1 -1 -1 -1 -1 -1 81 51 98 104 19 119 53 87 118 95 21
-1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 0 0 -1 -1 -1 -1 -1 77 111 90 118 111 12 83 115 109 104 100
"""

REF_B = """
0 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 0 -1 0 0 -1 0 -1 -1 0 -1 -1 -1 -1
-1 0 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 0 -1 -1 -1 114 89 0 29 0 0 49 0 0 0 0
-1 -1 0 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 108 18 -1 -1 -1 0 -1 0 4 108 -1 125 -1 -1 67 87 81 110 34
-1 -1 -1 0 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 0 -1 -1 -1 90 67 18 -1 31 -1 -1 -1 -1 -1 -1 -1
-1 -1 -1 -1 0 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 11 103 -1 111 -1 -1 123 35 -1 -1 -1
-1 -1 -1 -1 -1 0 -1 -1 -1 -1 -1 -1 -1 -1 -1 0 -1 -1 -1 -1 -1 115 -1 -1 -1 25 -1 25 2 28 -1 -1
-1 -1 -1 -1 -1 0 0 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 79 -1 108 -1 -1 96 6 21 -1
-1 -1 -1 -1 -1 -1 0 0 -1 -1 -1 -1 -1 -1 -1 -1 95 -1 47 -1 -1 -1 -1 -1 -1 65 58 -1 -1 91 -1 -1
-1 -1 -1 -1 -1 -1 -1 0 0 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 72 76 61 87 110 34 51 46 125 48 15
-1 -1 -1 -1 -1 -1 -1 -1 0 0 -1 -1 -1 -1 -1 -1 -1 -1 122 -1 -1 -1 -1 115 -1 -1 -1 -1 -1 30 120 84
-1 -1 -1 -1 -1 -1 -1 -1 -1 0 0 -1 -1 -1 -1 98 -1 110 -1 10 51 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1
-1 -1 -1 -1 -1 -1 -1 -1 -1 -1 0 0 -1 -1 -1 -1 101 -1 -1 -1 -1 -1 -1 17 -1 -1 5 -1 -1 -1 -1 64
-1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 0 0 -1 -1 -1 -1 55 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 107 66 111
-1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 0 0 -1 -1 -1 -1 -1 14 -1 -1 6 7 64 19 120 5 84 38 22 41
-1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 0 0 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 104 62 -1 -1 -1 114 85
-1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 0 0 -1 -1 -1 -1 -1 -1 -1 100 -1 -1 56 -1 -1 -1 45 94
"""

# 802.11n rate-1/2 Z=81 shaped, as recalled in SURVEY Appendix B, standard order [info | parity]
C4_INFO_PARITY = """
57 -1 -1 -1 50 -1 11 -1 50 -1 79 -1 1 0 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1
3 -1 28 -1 0 -1 -1 -1 55 7 -1 -1 -1 0 0 -1 -1 -1 -1 -1 -1 -1 -1 -1
30 -1 -1 -1 24 37 -1 -1 56 14 -1 -1 -1 -1 0 0 -1 -1 -1 -1 -1 -1 -1 -1
62 53 -1 -1 53 -1 -1 3 35 -1 -1 -1 -1 -1 -1 0 0 -1 -1 -1 -1 -1 -1 -1
40 -1 -1 20 66 -1 -1 22 28 -1 -1 -1 -1 -1 -1 -1 0 0 -1 -1 -1 -1 -1 -1
0 -1 -1 -1 8 -1 42 -1 50 -1 -1 8 -1 -1 -1 -1 -1 0 0 -1 -1 -1 -1 -1
69 79 79 -1 -1 -1 56 -1 52 -1 -1 -1 0 -1 -1 -1 -1 -1 0 0 -1 -1 -1 -1
65 -1 -1 -1 38 57 -1 -1 72 -1 27 -1 -1 -1 -1 -1 -1 -1 -1 0 0 -1 -1 -1
64 -1 -1 -1 14 52 -1 -1 30 -1 -1 32 -1 -1 -1 -1 -1 -1 -1 -1 0 0 -1 -1
-1 45 -1 70 0 -1 -1 -1 77 9 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 0 0 -1
2 56 -1 57 35 -1 -1 -1 -1 -1 12 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 0 0
24 -1 61 -1 60 -1 -1 27 51 -1 -1 16 1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 0
"""


def parse(txt):
    return np.array([[int(x) for x in line.split()] for line in txt.strip().split("\n")], dtype=np.int64)


def make_c3(seed=1, Z=384):
    """46x68 matrix with 5G NR BG1's shape: 4 core rows of weight 19 with a weight-3 + dual-diagonal core
    parity, 42 extension rows (weights 3..10, 240 entries) each owning one degree-1 parity column.
    Column order [46 parity | 22 info]; the two high-degree (punctured in 5G) info columns are LAST."""
    rng = np.random.RandomState(seed)
    b, c, npar = 46, 68, 46
    info = list(range(npar, c))
    hi = [c - 2, c - 1]
    while True:
        H = -np.ones((b, c), dtype=np.int64)
        # core parity: weight-3 column + dual diagonal
        H[0, 0], H[1, 0], H[3, 0] = 1, 0, 1
        H[0, 1] = H[1, 1] = 0
        H[1, 2] = H[2, 2] = 0
        H[2, 3] = H[3, 3] = 0
        for r in range(4):
            need = 19 - int((H[r] >= 0).sum())
            cols = set(hi)
            cols.update(rng.choice(info[:-2], need - 2, replace=False).tolist())
            for k in cols:
                H[r, k] = rng.randint(0, Z)
        # extension rows: weights 3..10 summing to 240 (identity column included), heavier rows first
        w = np.sort(rng.randint(3, 11, size=42))[::-1].copy()
        while w.sum() != 240:
            i = rng.randint(0, 42)
            if w.sum() < 240 and w[i] < 10:
                w[i] += 1
            elif w.sum() > 240 and w[i] > 3:
                w[i] -= 1
        w = np.sort(w)[::-1]
        for t, r in enumerate(range(4, b)):
            H[r, r] = 0                                   # degree-1 parity column (identity)
            need = int(w[t]) - 1
            cols = {hi[rng.randint(0, 2)]}
            pool = info[:-2] + [0, 1, 2, 3] + hi
            while len(cols) < need:
                cols.add(int(pool[rng.randint(0, len(pool))]))
            for k in cols:
                H[r, k] = rng.randint(0, Z)
        ok = (H >= 0).sum() == 316 and all((H[:, k] >= 0).sum() >= 2 for k in info)
        colset = {tuple(H[:, k]) for k in range(c)}
        if ok and len(colset) == c:
            return H


def write_jsonx(path, name, H, Z, extra=""):
    b, c = H.shape
    wid = max(len(str(int(v))) for v in H.flatten())
    with open(path, "w") as f:
        f.write("{\n")
        f.write('    name = "%s"\n' % name)
        f.write("    _lifting = %d\n" % Z)
        f.write(extra)
        f.write("    code = matrix (%d %d) {\n" % (b, c))
        for r in range(b):
            f.write("        " + " ".join(str(int(v)).rjust(wid) for v in H[r]) + "\n")
        f.write("    }\n}\n")


def main():
    A, B = parse(REF_A), parse(REF_B)
    assert A.shape == (16, 32) and (A >= 0).sum() == 122
    assert B.shape == (16, 32) and (B >= 0).sum() == 128
    C4s = parse(C4_INFO_PARITY)
    C4 = np.concatenate([C4s[:, 12:], C4s[:, :12]], axis=1)          # -> [parity | info]
    assert C4.shape == (12, 24) and (C4 >= 0).sum() == 86
    C3 = make_c3(seed=1)
    assert C3.shape == (46, 68) and (C3 >= 0).sum() == 316
    write_jsonx(os.path.join(HERE, "ref32x16_a.jsonx"), "REF-32x16-A", A, 126)
    write_jsonx(os.path.join(HERE, "ref32x16_b.jsonx"), "REF-32x16-B", B, 126)
    write_jsonx(os.path.join(HERE, "c4_wifi_12x24.jsonx"), "C4-80211n-shaped-12x24", C4, 81)
    write_jsonx(os.path.join(HERE, "c3_bg1_46x68.jsonx"), "C3-BG1-shaped-46x68-seed1", C3, 384)


if __name__ == "__main__":
    main()
