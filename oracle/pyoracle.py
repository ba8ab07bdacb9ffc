"""TEST INFRASTRUCTURE ONLY: ctypes bindings for the CPU oracle and the compiled reference.

* ``oracle/_build/libldpc_oracle.so``  -- this repo's plain-C restatement (ldpc_oracle.c)
* ``oracle/_ref/libldpcref.so``        -- the unmodified reference built from /root/reference
                                           by oracle/Makefile (travels to the GPU box prebuilt)

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "libldpc_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libldpcref.so")

BP, SP, ASP, MS, IMS, IASP, FHT, TASP, LMS, LCHE = range(10)
NAMES = {BP: "BP", SP: "SP", ASP: "ASP", MS: "MS", IMS: "IMS", IASP: "IASP", TASP: "TASP", LMS: "LMS", LCHE: "LCHE"}

# decoder parameters the reference hard-wires at its call site, bp_simulation.cpp:721-726 / decoders.h:43-48
MS_ALPHA, MS_BETA, MS_THR, MS_QBITS, MS_DBITS = 0.8, 0.4, 1.4, 6, 8

_p = np.ctypeslib.ndpointer


def build(ref=True):
    """(Re)build the oracle, and the reference shim when /root/reference is present."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if ref:
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


_oracle = None
_ref = None


def oracle():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        _oracle = C.CDLL(ORACLE_SO)
        _oracle.orc_sigma_bpsk.restype = C.c_double
        _oracle.orc_sigma_bpsk.argtypes = [C.c_double, C.c_int, C.c_int, C.c_int]
        _oracle.orc_sigma_qam.restype = C.c_double
        _oracle.orc_sigma_qam.argtypes = [C.c_double, C.c_int, C.c_int, C.c_int, C.c_int]
    return _oracle


def have_ref():
    return os.path.exists(REF_SO)


def ref():
    global _ref
    if _ref is None:
        if not have_ref():
            build(ref=True)
        _ref = C.CDLL(REF_SO)
    return _ref


def _hd16(hd):
    hd = np.ascontiguousarray(hd, dtype=np.int16)
    assert hd.ndim == 2
    return hd


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def ref_decode(dec, hd, Z, llr, maxiter, fresh=True, want_post=True):
    """Run the compiled reference decoder `dec` on llr[nf, N] (float64).
    Returns dict(hard uint8[nf,N], iters int32[nf], post float64[nf,N], aux (IMS: ims_y))."""
    hd = _hd16(hd)
    b, c = hd.shape
    llr = np.ascontiguousarray(llr, dtype=np.float64)
    nf, N = llr.shape
    assert N == c * Z
    hard = np.zeros((nf, N), np.uint8)
    iters = np.zeros(nf, np.int32)
    post = np.zeros((nf, N), np.float64) if want_post else None
    aux = np.zeros((nf, N), np.float64) if dec == IMS else None
    rc = ref().ref_decode(C.c_int(dec), _ptr(hd, C.c_short), b, c, Z, _ptr(llr, C.c_double), nf, maxiter,
                          int(bool(fresh)), _ptr(hard, C.c_ubyte), _ptr(iters, C.c_int),
                          _ptr(post, C.c_double), _ptr(aux, C.c_double))
    assert rc == 0, rc
    return dict(hard=hard, iters=iters, post=post, aux=aux)


def ref_decode_arrays(dec, hd, Z, llr1, maxiter, decision):
    """One call of the compiled reference decoder with the given `decision` on one frame: what it leaves in decword[] and
    in its input array.  Returns dict(decword float64[N], soft_after float64[N], iter)."""
    hd = _hd16(hd)
    b, c = hd.shape
    llr1 = np.ascontiguousarray(llr1, dtype=np.float64).reshape(-1)
    N = c * Z
    assert llr1.size == N
    decword, after, it = np.zeros(N), np.zeros(N), C.c_int(0)
    rc = ref().ref_decode_arrays(C.c_int(dec), _ptr(hd, C.c_short), b, c, Z, _ptr(llr1, C.c_double), maxiter, int(decision),
                                 _ptr(decword, C.c_double), _ptr(after, C.c_double), C.byref(it))
    assert rc == 0, rc
    return dict(decword=decword, soft_after=after, iter=it.value)


def ref_permutation(hd, Z, modulation, mode, block=1, inter=1):
    """The compiled reference's interleaver tables (Permutations_Open / Permutation_Init / Permutation on index ramps):
    (direct[N], inverse[N]); modulation as in bp_simulation (0 none, 1 QAM-4, 2 QAM-16, 3 QAM-64, 4 QAM-256)."""
    hd = _hd16(hd)
    b, c = hd.shape
    QAM, half = [(1, 1), (4, 1), (16, 2), (64, 3), (256, 4)][modulation]
    d, i = np.zeros(c * Z, np.int32), np.zeros(c * Z, np.int32)
    rc = ref().ref_permutation(_ptr(hd, C.c_short), b, c, Z, QAM, half, mode, block, inter, _ptr(d, C.c_int), _ptr(i, C.c_int))
    assert rc == 0
    return d, i


def ref_girth(hd, Z, gtarget=4):
    """The compiled reference's trace_bound_pol_mon_pm() as its driver calls it -> (girth, ACE[gtarget], spectrum[gtarget])
    exactly as main_simulation.cpp:148-205 derives them."""
    H = np.ascontiguousarray(hd, dtype=np.int32)
    b, c = H.shape
    S, SA = np.zeros(20, np.int32), np.zeros(20, np.int32)
    rc = ref().ref_girth(_ptr(H, C.c_int), b, c, Z, gtarget, _ptr(S, C.c_int), _ptr(SA, C.c_int))
    assert rc == 0
    girth = next((g for g in range(1, 21) if S[g - 1]), 21)
    ace = [int(x) for x in SA[girth - 1:] if x][:gtarget]
    spec = [int(x) for x in S[girth - 1:] if x][:gtarget]
    return girth, ace + [0] * (gtarget - len(ace)), spec + [0] * (gtarget - len(spec))


def orc_decode(dec, hd, Z, llr, maxiter, dtype=np.float64, chain=False, alpha=MS_ALPHA):
    """Run the C oracle's restatement of decoder `dec`.  dtype=float32 is available for LMS/MS."""
    hd = _hd16(hd)
    b, c = hd.shape
    o = oracle()
    llr = np.ascontiguousarray(llr, dtype=dtype)
    nf, N = llr.shape
    assert N == c * Z
    hard = np.zeros((nf, N), np.uint8)
    iters = np.zeros(nf, np.int32)
    ct = C.c_double if dtype == np.float64 else C.c_float
    aux = None
    if dec in (LMS, MS):
        post = np.zeros((nf, N), dtype)
        if dec == LMS:
            fn = o.orc_lms_f64 if dtype == np.float64 else o.orc_lms_f32
            rc = fn(_ptr(hd, C.c_short), b, c, Z, _ptr(llr, ct), nf, maxiter, _ptr(hard, C.c_ubyte),
                    _ptr(iters, C.c_int), _ptr(post, ct))
        else:
            fn = o.orc_ms_f64 if dtype == np.float64 else o.orc_ms_f32
            rc = fn(_ptr(hd, C.c_short), b, c, Z, _ptr(llr, ct), nf, maxiter, ct(alpha), _ptr(hard, C.c_ubyte),
                    _ptr(iters, C.c_int), _ptr(post, ct))
    elif dec == IMS:
        assert dtype == np.float64
        post = np.zeros((nf, N), np.int16)
        aux = np.zeros((nf, N), np.int16)
        rc = o.orc_ims(_ptr(hd, C.c_short), b, c, Z, _ptr(llr, ct), nf, maxiter, C.c_double(alpha),
                       C.c_double(MS_THR), MS_QBITS, MS_DBITS, _ptr(hard, C.c_ubyte), _ptr(iters, C.c_int),
                       _ptr(post, C.c_short), _ptr(aux, C.c_short))
    elif dec == IASP:
        assert dtype == np.float64
        post = np.zeros((nf, N), np.uint16)
        rc = o.orc_iasp(_ptr(hd, C.c_short), b, c, Z, _ptr(llr, ct), nf, maxiter, _ptr(hard, C.c_ubyte),
                        _ptr(iters, C.c_int), _ptr(post, C.c_ushort))
    else:
        assert dtype == np.float64
        post = np.zeros((nf, N), np.float64)
        if dec == BP:
            rc = o.orc_bp(_ptr(hd, C.c_short), b, c, Z, _ptr(llr, ct), nf, maxiter, int(bool(chain)),
                          _ptr(hard, C.c_ubyte), _ptr(iters, C.c_int), _ptr(post, ct))
        else:
            fn = {TASP: o.orc_tasp, ASP: o.orc_asp, SP: o.orc_sp, LCHE: o.orc_lche}[dec]
            rc = fn(_ptr(hd, C.c_short), b, c, Z, _ptr(llr, ct), nf, maxiter, _ptr(hard, C.c_ubyte),
                    _ptr(iters, C.c_int), _ptr(post, ct))
    assert rc == 0, rc
    return dict(hard=hard, iters=iters, post=post, aux=aux)


def ref_demodulate(Q, ns, sigma, x, T=26.0, out_type=0):
    m = int(np.log2(Q))
    x = np.ascontiguousarray(x, np.float64)
    assert x.size == 2 * ns
    res = np.zeros(ns * m, np.float64)
    rc = ref().ref_demodulate(Q, m, ns * m, ns, C.c_double(sigma), C.c_double(T), out_type,
                              _ptr(x, C.c_double), _ptr(res, C.c_double))
    assert rc == 0
    return res


def orc_demodulate(Q, ns, sigma, x, T=26.0, out_type=0):
    m = int(np.log2(Q))
    x = np.ascontiguousarray(x, np.float64)
    assert x.size == 2 * ns
    res = np.zeros(ns * m, np.float64)
    rc = oracle().orc_demodulate(m, ns, C.c_double(sigma), C.c_double(T), out_type,
                                 _ptr(x, C.c_double), _ptr(res, C.c_double))
    assert rc == 0
    return res


def ref_modulate(Q, bits):
    m = int(np.log2(Q))
    bits = np.ascontiguousarray(bits, np.float64)
    L = bits.size
    ns = (L + m - 1) // m
    out = np.zeros(2 * ns, np.float64)
    got = ref().ref_modulate(Q, L, m, _ptr(bits, C.c_double), _ptr(out, C.c_double))
    assert got == ns
    return out


def orc_modulate(Q, bits):
    m = int(np.log2(Q))
    bits = np.ascontiguousarray(bits, np.uint8)
    ns = bits.size // m
    out = np.zeros(2 * ns, np.float64)
    oracle().orc_modulate(m, ns, _ptr(bits, C.c_ubyte), _ptr(out, C.c_double))
    return out


def ref_bp_simulation(hd, Z, maxiter, n_frame_errors, n_experiments, snr, ref_fer, dec, modulation=0,
                      punct=0, seed=1):
    H = np.ascontiguousarray(hd, dtype=np.int32)
    b, c = H.shape
    ber, fer = C.c_double(), C.c_double()
    ref().ref_bp_simulation(_ptr(H, C.c_int), b, c, Z, maxiter, n_frame_errors, n_experiments, C.c_double(snr),
                            C.c_double(ref_fer), dec, modulation, punct, seed, C.byref(ber), C.byref(fer))
    return ber.value, fer.value


def ref_random_codeword(hd, Z, seed):
    """The reference's random_codeword() right after reset_random(seed): (exit code, codeword bits or None)."""
    H = np.ascontiguousarray(hd, dtype=np.int32)
    b, c = H.shape
    out = np.zeros(c * Z, np.uint8)
    rc = ref().ref_random_codeword(_ptr(H, C.c_int), b, c, Z, seed, _ptr(out, C.c_ubyte))
    return rc, (out if rc == 0 else None)


def sigma_bpsk(snr_db, b, c, punct=0):
    return oracle().orc_sigma_bpsk(snr_db, b, c, punct)


def sigma_qam(snr_db, b, c, punct, Q):
    return oracle().orc_sigma_qam(snr_db, b, c, punct, Q)
