"""The C++ drop-in boundary on the GPU: the decoders.h interface used the way the reference's callers use it,
bin/main simulation against the reference's own result file, Demodulate()/QAM_modulator() at the function
boundary, and the sharded frame loop on a real decoder."""
import os
import re
import subprocess

import numpy as np
import pytest

from codes import load_code, awgn_llr
from conftest import ROOT
from test_frame_loop import load_simhost

pytestmark = pytest.mark.gpu
PKG = os.path.join(ROOT, "ldpc-lib_b200")


@pytest.fixture(scope="module")
def compat_bin(tmp_path_factory):
    out = tmp_path_factory.mktemp("cpp") / "decoders_compat"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(PKG, "host"), os.path.join(ROOT, "tests", "cpp", "decoders_compat_main.cpp"),
                           os.path.join(PKG, "libldpcb200_host.a"), "-L", PKG, "-lldpcb200", "-lpthread", "-Wl,-rpath," + PKG, "-o", str(out)])
    return str(out)


@pytest.mark.parametrize("dec", ["BP", "SP", "ASP", "MS", "IMS", "IASP", "TASP", "LMS", "LCHE"])
def test_decoders_h_interface_matches_reference_golden(po, compat_bin, tmp_path, dec):
    """Same buffers, same call sequence as bp_simulation.cpp -> the reference's recorded decisions and return values
    (frame after frame on ONE state, so BP_DEC's stale-syndrome carry-over applies, as in the reference)."""
    G = np.load(os.path.join(ROOT, "tests", "golden", "decoders_c4_z27.npz"))
    hd, Z, llr = G["hd"], int(G["Z"]), G["llr"]
    did = getattr(po, dec)
    maxiter = 3 if dec == "BP" else int(G["maxiter"])
    hd.astype(np.int16).tofile(tmp_path / "hd.bin")
    llr.tofile(tmp_path / "llr.bin")
    subprocess.check_call([compat_bin, str(did), str(hd.shape[0]), str(hd.shape[1]), str(Z), str(maxiter),
                           str(tmp_path / "hd.bin"), str(tmp_path / "llr.bin"), str(tmp_path / "out.bin")])
    raw = np.fromfile(tmp_path / "out.bin", np.uint8).reshape(llr.shape[0], 4 + llr.shape[1])
    iters = raw[:, :4].copy().view(np.int32).ravel()
    hard = raw[:, 4:]
    key = "BP_chain" if dec == "BP" else dec
    assert np.array_equal(iters, G[key + "_iters"])
    assert np.array_equal(np.packbits(hard, axis=1), G[key + "_hard"])


@pytest.mark.parametrize("dec", ["BP", "SP", "ASP", "MS", "IMS", "IASP", "TASP", "LMS", "LCHE"])
def test_decision_argument_and_input_side_effects_match_the_reference(po, compat_bin, tmp_path, dec):
    """decision != 0: decword carries what the reference puts there (its soft output; TASP_DEC and LCHE_DEC ignore the
    argument), and the input array is left as the reference leaves it (posteriors for BP / SP, channel probabilities for
    ASP / TASP / IASP, untouched for the min-sum family and LCHE) -- against the compiled reference, call by call."""
    if not po.have_ref():
        pytest.skip("oracle/_ref not built")
    G = np.load(os.path.join(ROOT, "tests", "golden", "decoders_c4_z27.npz"))
    hd, Z, llr = G["hd"], int(G["Z"]), G["llr"]
    did = getattr(po, dec)
    maxiter = int(G["maxiter"])
    N = llr.shape[1]
    hd.astype(np.int16).tofile(tmp_path / "hd.bin")
    for f in (0, 5):
        llr[f:f + 1].tofile(tmp_path / "llr.bin")
        subprocess.check_call([compat_bin, str(did), str(hd.shape[0]), str(hd.shape[1]), str(Z), str(maxiter),
                               str(tmp_path / "hd.bin"), str(tmp_path / "llr.bin"), str(tmp_path / "out.bin"), "1"])
        raw = np.fromfile(tmp_path / "out.bin", np.uint8)
        it = int(raw[:4].view(np.int32)[0])
        decword, after = raw[4:4 + 8 * N].view(np.float64), raw[4 + 8 * N:].view(np.float64)
        want = po.ref_decode_arrays(did, hd, Z, llr[f], maxiter, 1)
        assert it == want["iter"]
        if dec in ("MS", "IMS", "IASP", "LMS", "TASP", "LCHE"):        # exact arithmetic, or 0 / 1
            assert np.array_equal(decword, want["decword"]), dec
        else:                                                            # exp / log: libm vs CUDA; BP / SP: regrouped expressions (bpsp4.cu), whose
            tol = 1e-5 if dec in ("BP", "SP") else 1e-9                  # cancellation in (AA - 1) / (AA + 1) amplifies last-bit differences; the bar is 1e-4
            assert np.allclose(decword, want["decword"], rtol=tol, atol=1e-12), dec
        if dec in ("BP", "SP"):
            assert np.allclose(after, want["soft_after"], rtol=1e-5, atol=1e-12)
        else:
            assert np.array_equal(after, want["soft_after"]), dec


def wilson(k, n, z=1.96):
    p = k / n
    d = 1 + z * z / n
    c = (p + z * z / (2 * n)) / d
    h = z * np.sqrt(p * (1 - p) / n + z * z / (4 * n * n)) / d
    return c - h, c + h


def parse_result(path):
    txt = open(path).read()
    fer = [float(x) for x in re.search(r"FER = array \{([^}]*)\}", txt).group(1).split()]
    ber = [float(x) for x in re.search(r"BER = array \{([^}]*)\}", txt).group(1).split()]
    return fer, ber, txt


@pytest.mark.parametrize("name,errors", [("sim_c1_lms", 100), ("sim_c1_tasp", 100)])
def test_main_simulation_curve_inside_reference_interval(tmp_path, name, errors):
    """`main simulation` on the same input file as the reference: every FER point must fall inside the 95 % (Wilson)
    interval of the reference's estimate (100 frame errors per point), widened by our own sampling error."""
    out = tmp_path / "out.jsonx"
    r = subprocess.run([os.path.join(PKG, "bin", "main"), "simulation", os.path.join(ROOT, "configs", name + ".jsonx"), str(out)],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr
    fer, ber, txt = parse_result(out)
    rfer, rber, rtxt = parse_result(os.path.join(ROOT, "tests", "golden", "ref_%s.jsonx" % name))
    assert len(fer) == len(rfer) == 3
    for p, q in zip(fer, rfer):
        n_ref = errors / q
        lo, hi = wilson(errors, n_ref)
        mylo, myhi = wilson(errors, errors / p)
        assert myhi >= lo and mylo <= hi, (p, q, lo, hi)
    for a, bb in zip(ber, rber):
        assert 0.3 * bb <= a <= 3.0 * bb + 1e-6, (a, bb)
    # same descriptor keys, same matrix text
    keys = lambda t: sorted(set(re.findall(r"^  (\w+) =", t, re.M)))
    assert keys(txt) == keys(rtxt)
    mat = lambda t: re.search(r"code = matrix \(16 32\) \{.*?\n  \}", t, re.S).group(0)
    assert mat(txt) == mat(rtxt)


def test_main_simulation_interleaver_modes(tmp_path):
    """permutation_type through the CLI (settings/permutation_type, _block, _inter -> bp_simulation -> ldpcb200_set_interleaver).
    BPSK: the reference de-interleaves i.i.d. LLRs of the all-zero codeword (bp_simulation.cpp:573, :684), so a mode changes
    which noise sample a bit sees but not the statistics.  QAM-64: the bit positions of a symbol differ in reliability, the
    deterministic mode 2 (heaviest block columns on one fixed bit position) and the random mode 1 both run and give curves."""
    base = open(os.path.join(ROOT, "configs", "sim_c1_lms.jsonx")).read().replace("error_blocks = 100", "error_blocks = 40")
    (tmp_path / "sim_c1_lms_codes.jsonx").write_text(open(os.path.join(ROOT, "configs", "sim_c1_lms_codes.jsonx")).read())   # @"..." is relative
    outs = {}
    for perm, mod in [(0, 0), (1, 0), (0, 3), (1, 3), (2, 3)]:
        cfg = tmp_path / ("in_%d_%d.jsonx" % (perm, mod))
        cfg.write_text(base.replace("permutation_type = 0", "permutation_type = %d" % perm).replace("modulation_type = 0", "modulation_type = %d" % mod))
        out = tmp_path / ("out_%d_%d.jsonx" % (perm, mod))
        r = subprocess.run([os.path.join(PKG, "bin", "main"), "simulation", str(cfg), str(out)], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, (perm, mod, r.stderr)
        outs[(perm, mod)] = parse_result(out)
    fer0, fer1 = outs[(0, 0)][0], outs[(1, 0)][0]
    for p, q in zip(fer0, fer1):                                     # same statistics: 40 errors per point each
        lo, hi = wilson(40, 40 / q)
        mylo, myhi = wilson(40, 40 / p)
        assert myhi >= lo and mylo <= hi, (p, q)
    for key in [(0, 3), (1, 3), (2, 3)]:
        assert all(0.0 < f <= 1.0 for f in outs[key][0]), (key, outs[key][0])


def test_one_process_all_gpus_gives_the_single_gpu_result(tmp_path):
    """LDPCB200_DEVICES=all: one handle and one host thread per GPU inside one process (host/bp_simulation.cpp); frames are a
    pure function of (seed, frame index) and the stop rules run in frame order, so the result file is the single-GPU one."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least two GPUs")
    base = open(os.path.join(ROOT, "configs", "sim_c1_lms.jsonx")).read().replace("error_blocks = 100", "error_blocks = 30")
    (tmp_path / "sim_c1_lms_codes.jsonx").write_text(open(os.path.join(ROOT, "configs", "sim_c1_lms_codes.jsonx")).read())
    cfg = tmp_path / "in.jsonx"
    cfg.write_text(base)
    res = {}
    for devs in ("", "all", "0,1"):
        out = tmp_path / ("out_%s.jsonx" % (devs.replace(",", "_") or "one"))
        env = dict(os.environ)
        env.pop("LDPCB200_DEVICES", None)
        if devs:
            env["LDPCB200_DEVICES"] = devs
        r = subprocess.run([os.path.join(PKG, "bin", "main"), "simulation", str(cfg), str(out)], capture_output=True, text=True, timeout=900, env=env)
        assert r.returncode == 0, (devs, r.stderr)
        res[devs] = parse_result(out)[:2]
    assert res["all"] == res[""] and res["0,1"] == res[""]


def test_interrupted_point_returns_minus_one_like_the_reference(tmp_path):
    """bp_simulation_request_interrupt(): the reference's 'x' console hook ends the current point with (-1, -1)
    (bp_simulation.cpp:590, :825-829); the call after it runs normally."""
    exe = tmp_path / "interrupt_main"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(PKG, "host"), os.path.join(ROOT, "tests", "cpp", "interrupt_main.cpp"),
                           os.path.join(PKG, "libldpcb200_host.a"), "-L", PKG, "-lldpcb200", "-lpthread", "-Wl,-rpath," + PKG, "-o", str(exe)])
    hd, _ = load_code("c4_wifi_12x24")
    hd.astype(np.int16).tofile(tmp_path / "hd.bin")
    out = subprocess.run([str(exe), "12", "24", "81", str(tmp_path / "hd.bin")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    r = [float(x) for x in out.stdout.split()]
    assert r[0] == -1.0 and r[1] == -1.0
    assert 0.0 <= r[2] < 1.0 and 0.0 < r[3] <= 1.0


def test_search_output_without_q_mod_is_accepted(tmp_path):
    """The reference's `search` leaves `_q_mod` out of the records it writes (main_good_code_search.cpp:383-399) and its
    `simulation` then refuses them; the drop-in driver reads a missing field as 2 (binary), so search -> simulation round-trips."""
    base = open(os.path.join(ROOT, "configs", "sim_c1_lms.jsonx")).read().replace("error_blocks = 100", "error_blocks = 10")
    codes = open(os.path.join(ROOT, "configs", "sim_c1_lms_codes.jsonx")).read()
    assert "_q_mod" in codes
    (tmp_path / "sim_c1_lms_codes.jsonx").write_text(re.sub(r"^\s*_q_mod\s*=.*\n", "", codes, flags=re.M))
    cfg = tmp_path / "in.jsonx"
    cfg.write_text(base)
    out = tmp_path / "out.jsonx"
    r = subprocess.run([os.path.join(PKG, "bin", "main"), "simulation", str(cfg), str(out)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    assert len(parse_result(out)[0]) == 3


@pytest.mark.parametrize("Q", [16, 64, 256])
def test_demodulate_function_boundary(ldpc, Q):
    """Demodulate() with m = log2(Q), LLR and P1 outputs, against the reference's recorded outputs: only exp/log may
    differ (last ulp), the +-T clip decisions and the NaN case (all points beyond T) are identical."""
    D = np.load(os.path.join(ROOT, "tests", "golden", "demod.npz"))
    x, sigma = D["x%d" % Q], float(D["sigma%d" % Q])
    ns = x.size // 2
    for out_type, key in ((0, "llr%d"), (1, "p1_%d")):
        got = ldpc.demodulate(Q, ns, sigma, x, 26.0, out_type)
        want = D[key % Q]
        assert np.array_equal(np.isnan(got), np.isnan(want))
        ok = ~np.isnan(want)
        clip = np.abs(want) == 26.0 if out_type == 0 else (want == 0.0) | (want == 1.0)
        assert np.array_equal(got[ok & clip], want[ok & clip])
        assert np.allclose(got[ok], want[ok], rtol=1e-12, atol=1e-15)
    assert np.array_equal(ldpc.modulate(Q, D["bits%d" % Q]), D["mod%d" % Q])


def test_demodulate_qam4(ldpc):
    D = np.load(os.path.join(ROOT, "tests", "golden", "demod.npz"))
    assert np.array_equal(ldpc.demodulate(4, 100, 0.7, D["x4"]), D["llr4"])


def test_generated_llr_statistics_and_determinism(ldpc, po):
    """The counter-based channel: bit-identical across calls and batch splits, N(0,1) noise after un-scaling, and the
    QAM-64 LLRs it produces equal Demodulate() of the received symbols it implies."""
    hd, _ = load_code("ref32x16_b")
    with ldpc.Decoder(hd, 126, po.LMS, precision=32) as d:
        a = d.generate_llr(2.0, 64, seed=9, stream=3, first_frame=100)
        b = np.concatenate([d.generate_llr(2.0, 40, seed=9, stream=3, first_frame=100),
                            d.generate_llr(2.0, 24, seed=9, stream=3, first_frame=140)])
        assert np.array_equal(a, b)
        assert not np.array_equal(a, d.generate_llr(2.0, 64, seed=10, stream=3, first_frame=100))
        sigma = ldpc.sigma(16, 32, 0, 2.0)
        noise = (1.0 - a.astype(np.float64) * sigma * sigma / 2.0) / sigma          # invert llr = (1 - sigma n) 2 / sigma^2
        assert abs(noise.mean()) < 0.01 and abs(noise.std() - 1.0) < 0.01
        assert abs(((noise ** 4).mean()) - 3.0) < 0.1
        p = d.generate_llr(2.0, 4, punct=2, seed=9)
        assert np.all(p[:, -2 * 126:] == 0.5) and not np.any(p[:, :-2 * 126] == 0.5)   # bp_simulation.cpp:700-709
        f64 = d.generate_llr(2.0, 8, seed=9, dtype=np.float64)
        assert np.array_equal(f64, d.generate_llr(2.0, 8, seed=9).astype(np.float64))
    hd3, Z3 = load_code("c3_bg1_46x68")
    with ldpc.Decoder(hd3, 96, po.TASP) as d:                                     # N = 6528, a multiple of 6
        q = d.generate_llr(6.0, 3, modulation=ldpc.MOD_QAM64, punct=2, seed=4)
        assert np.all(q[:, -2 * 96:] == 0.0)                                       # probability-domain decoders get 0
        body = q[:, :-2 * 96].astype(np.float64)
        assert np.isfinite(body).all() and (body > 0).mean() > 0.8                 # all-zero codeword: mostly positive LLRs


def test_sharded_frame_loop_on_device(ldpc, po):
    """simhost.bp_simulation on a real decoder equals the frame-ordered scan of one big simulate() call."""
    sh = load_simhost()
    hd, _ = load_code("ref32x16_b")
    with ldpc.Decoder(hd, 126, po.LMS, precision=32) as d:
        ber, fer, r = sh.bp_simulation(d, 10, 40, 50000, 2.0, 1.0, seed=2, round_frames=500)
        pf = d.simulate(2.0, r.experiment, 10, seed=2, want_per_frame=True)["per_frame"]
    err = pf >> 31
    assert int(err.sum()) == r.nde == 40 and err[-1] == 1
    assert r.nse == int((pf[err == 1] & 0xFFFFFF).sum())
    assert fer == 40 / r.experiment


LEVEL2 = os.path.join(ROOT, "oracle", "_ref", "main_level2")


@pytest.mark.skipif(not os.path.exists(LEVEL2), reason="oracle/_ref/main_level2 not built (make -C oracle level2, needs /root/reference)")
def test_level2_reference_driver_on_the_engine(tmp_path):
    """INTEGRATION.md level 2: the reference's OWN main.cpp / main_simulation.cpp / settings.cpp (compiled from
    /root/reference) linked against this repository's bp_simulation + decoders over libldpcb200.so.  Same input file, and
    the FER points must agree with the reference binary's result file like the native driver's do."""
    out = tmp_path / "out.jsonx"
    r = subprocess.run([LEVEL2, "simulation", os.path.join(ROOT, "configs", "sim_c1_lms.jsonx"), str(out)],
                       capture_output=True, text=True, timeout=900, env=dict(os.environ, LDPCB200_JIT="0"))
    assert r.returncode == 0, r.stderr[-2000:]
    assert "girth:" in r.stdout                                   # the reference's own girth / ACE print is back in this build
    fer, ber, txt = parse_result(out)
    rfer, rber, rtxt = parse_result(os.path.join(ROOT, "tests", "golden", "ref_sim_c1_lms.jsonx"))
    for p, q in zip(fer, rfer):
        lo, hi = wilson(100, 100 / q)
        mylo, myhi = wilson(100, 100 / p)
        assert myhi >= lo and mylo <= hi, (p, q)
    assert "girth_ACE = array { 16 14 17 15 }" in txt             # computed by the reference's trace_pm, unchanged
