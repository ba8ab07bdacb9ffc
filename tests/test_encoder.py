"""The host-side QC encoder (ldpc-lib_b200/host/encoder.cpp, row A16) against the reference's random_codeword() golden
vectors: same exit codes ("bad encoding" for parity structures the encoder does not understand), same codewords for the
same seed, and every codeword satisfies H c = 0."""
import os
import subprocess

import numpy as np
import pytest

from codes import load_code
from conftest import ROOT

PKG = os.path.join(ROOT, "ldpc-lib_b200")
G = np.load(os.path.join(ROOT, "tests", "golden", "encoder.npz"))


@pytest.fixture(scope="module")
def tool(tmp_path_factory):
    if not os.path.exists(os.path.join(PKG, "libldpcb200_host.a")):
        subprocess.check_call(["make", "-s", "-C", PKG, "host"])
    out = tmp_path_factory.mktemp("enc") / "encoder_main"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(PKG, "host"), os.path.join(ROOT, "tests", "cpp", "encoder_main.cpp"),
                           os.path.join(PKG, "libldpcb200_host.a"), "-o", str(out)])
    return str(out)


CASES = [("ref32x16_a", 126), ("ref32x16_b", 126), ("ref32x16_b", 256), ("c4_wifi_12x24", 27), ("c3_bg1_46x68", 16)]


@pytest.mark.parametrize("name,Z", CASES)
@pytest.mark.parametrize("seed", [1, 5])
def test_random_codeword_matches_reference(tool, tmp_path, name, Z, seed):
    hd, _ = load_code(name)
    hd = np.where(hd > 0, hd % Z, hd).astype(np.int32)
    b, c = hd.shape
    hd.tofile(tmp_path / "hd.bin")
    subprocess.check_call([tool, str(b), str(c), str(Z), str(seed), str(tmp_path / "hd.bin"), str(tmp_path / "out.bin")])
    raw = np.fromfile(tmp_path / "out.bin", np.uint8)
    rc = int(raw[:4].view(np.int32)[0])
    key = "%s_Z%d_s%d" % (name, Z, seed)
    assert rc == int(G[key + "_rc"])
    if rc == 0:
        cw = raw[4:]
        assert cw.size == c * Z and cw[b * Z:].any()                      # random information bits
        assert np.array_equal(np.packbits(cw), G[key + "_cw"])
        for j in range(b):                                                # H c = 0, lane formulation
            s = np.zeros(Z, np.uint8)
            for i in range(c):
                if hd[j, i] >= 0:
                    s ^= np.roll(cw[i * Z:(i + 1) * Z], -int(hd[j, i]))
            assert not s.any()
