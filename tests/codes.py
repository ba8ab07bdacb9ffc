"""Base matrices of the benchmark configurations, read from configs/*.jsonx (see configs/make_configs.py)."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_code(name):
    """-> (hd int16[b, c], default lifting Z) from configs/<name>.jsonx"""
    txt = open(os.path.join(ROOT, "configs", name + ".jsonx")).read()
    m = re.search(r"code\s*=\s*matrix\s*\((\d+)\s+(\d+)\)\s*\{(.*?)\}", txt, re.S)
    b, c = int(m.group(1)), int(m.group(2))
    hd = np.array(m.group(3).split(), dtype=np.int16).reshape(b, c)
    Z = int(re.search(r"_lifting\s*=\s*(\d+)", txt).group(1))
    return hd, Z


def awgn_llr(rng, nf, N, b, c, snr_db, punct=0, dtype=np.float64):
    """All-zero codeword over BPSK/AWGN, LLR = -2(sigma*n - 1)/sigma^2 (bp_simulation.cpp:444-445,603)."""
    rate = (c - b) / (c - punct)
    sigma = np.sqrt(10 ** (-snr_db / 10) / 2 / rate)
    noise = rng.standard_normal((nf, N))
    return (-2.0 * (sigma * noise - 1.0) / (sigma * sigma)).astype(dtype)
