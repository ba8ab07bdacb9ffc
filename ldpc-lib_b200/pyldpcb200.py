"""ctypes binding of the C ABI in include/ldpcb200.h (libldpcb200.so).

This is the Python face of the drop-in boundary: thin, no arithmetic.  numpy arrays are host buffers;
objects with ``data_ptr()`` (torch CUDA tensors) are passed as device pointers.  The library is loaded
from this directory (it is built in-tree by ``make -C ldpc-lib_b200`` / ``__graft_entry__.build()``);
there is no fallback when it is missing or when there is no CUDA device.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libldpcb200.so")

# enum DEC_ID, decoders.h:16-28
BP_DEC, SP_DEC, ASP_DEC, MS_DEC, IMS_DEC, IASP_DEC, FHT_DEC, TASP_DEC, LMS_DEC, LCHE_DEC = range(10)
DECODER_NAMES = {BP_DEC: "BP_DEC", SP_DEC: "SP_DEC", ASP_DEC: "ASP_DEC", MS_DEC: "MS_DEC", IMS_DEC: "IMS_DEC",
                 IASP_DEC: "IASP_DEC", TASP_DEC: "TASP_DEC", LMS_DEC: "LMS_DEC", LCHE_DEC: "LCHE_DEC"}
# enum MODULATION_TYPE, modulation.h:4-11
MOD_BPSK, MOD_QAM4, MOD_QAM16, MOD_QAM64, MOD_QAM256 = range(5)
F64, F32, I16, U16 = range(4)
LLR_ON_DEVICE, OUT_ON_DEVICE, HARD_PACKED, NO_EARLY_EXIT, BP_CHAIN_SYNDROME = 1, 2, 4, 8, 16
OK, EINVAL, ENODEV, ECUDA, ENOMEM, EUNSUPPORTED = 0, -1, -2, -3, -4, -6


class Params(C.Structure):
    _fields_ = [("alpha", C.c_double), ("beta", C.c_double), ("thr", C.c_double), ("qbits", C.c_int),
                ("dbits", C.c_int), ("precision", C.c_int), ("device", C.c_int), ("use_fast", C.c_int),
                ("reserved", C.c_int * 8)]


class SimParams(C.Structure):
    _fields_ = [("snr_db", C.c_double), ("modulation", C.c_int), ("punctured_blocks", C.c_int),
                ("max_iterations", C.c_int), ("seed", C.c_uint64), ("stream", C.c_uint32),
                ("first_frame", C.c_uint64), ("n_frames", C.c_uint32), ("flags", C.c_uint32),
                ("qam_T", C.c_double)]


class Counters(C.Structure):
    _fields_ = [("frames", C.c_uint64), ("frame_errors", C.c_uint64), ("info_bit_errors", C.c_uint64),
                ("undetected", C.c_uint64), ("iter_sum", C.c_uint64), ("bit_errors", C.c_uint64)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


class LdpcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("ldpcb200 error %d: %s" % (code, msg))
        self.code = code


_lib = None


def lib():
    """Load libldpcb200.so (raises if it has not been built -- there is nothing to fall back to)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: run `make -C ldpc-lib_b200` (or __graft_entry__.build())" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.ldpcb200_last_error.restype = C.c_char_p
        L.ldpcb200_sigma.restype = C.c_double
        L.ldpcb200_sigma.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_int]
        L.ldpcb200_stream.restype = C.c_void_p
        L.ldpcb200_stream.argtypes = [C.c_void_p]
        L.ldpcb200_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Params), C.POINTER(C.c_void_p)]
        L.ldpcb200_destroy.argtypes = [C.c_void_p]
        L.ldpcb200_info.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 4
        L.ldpcb200_kernel_info.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 5
        L.ldpcb200_decode_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.ldpcb200_simulate.argtypes = [C.c_void_p, C.POINTER(SimParams), C.POINTER(Counters), C.c_void_p]
        L.ldpcb200_generate_llr.argtypes = [C.c_void_p, C.POINTER(SimParams), C.c_void_p, C.c_int]
        L.ldpcb200_demodulate.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.ldpcb200_modulate.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.ldpcb200_jit_check.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.ldpcb200_last_kernel_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int)]
        L.ldpcb200_girth_spectrum.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ldpcb200_interleaver_tables.argtypes = [C.c_void_p] + [C.c_int] * 7 + [C.c_void_p, C.c_void_p]
        L.ldpcb200_set_interleaver.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ldpcb200_set_codeword.argtypes = [C.c_void_p, C.c_void_p]
        L.ldpcb200_generate_noise.argtypes = [C.c_void_p, C.POINTER(SimParams), C.c_int, C.c_void_p]
        L.ldpcb200_simulate_codes.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(SimParams), C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise LdpcError(rc, lib().ldpcb200_last_error().decode())


def _is_dev(x):
    return hasattr(x, "data_ptr")


def _ptr(x):
    if x is None:
        return None
    if _is_dev(x):
        return C.c_void_p(x.data_ptr())
    return C.c_void_p(x.ctypes.data)


def sigma(b, c, punctured_blocks, snr_db, modulation=MOD_BPSK):
    return lib().ldpcb200_sigma(b, c, punctured_blocks, snr_db, modulation)


class Decoder:
    """One (GPU, code, decoder) handle -- decod_open + hd fill + decod_init of the reference
    (decoders.h:293-294, bp_simulation.cpp:353-382); close() is decod_close."""

    def __init__(self, hd, Z, decoder_id, precision=64, alpha=0.8, thr=1.4, qbits=6, dbits=8, device=-1, use_fast=True):
        hd = np.ascontiguousarray(hd, dtype=np.int16)
        assert hd.ndim == 2
        self.b, self.c = hd.shape
        self.Z, self.decoder_id, self.precision = Z, decoder_id, precision
        p = Params()
        lib().ldpcb200_default_params(C.byref(p))
        p.alpha, p.thr, p.qbits, p.dbits, p.precision, p.device, p.use_fast = alpha, thr, qbits, dbits, precision, device, int(use_fast)
        self._h = C.c_void_p()
        _check(lib().ldpcb200_create(_ptr(hd), self.b, self.c, Z, decoder_id, C.byref(p), C.byref(self._h)))
        n, r, e, d = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _check(lib().ldpcb200_info(self._h, C.byref(n), C.byref(r), C.byref(e), C.byref(d)))
        self.N, self.R, self.E, self.device = n.value, r.value, e.value, d.value
        self.K = self.N - self.R
        self.nwords = (self.N + 31) // 32

    def close(self):
        if self._h:
            lib().ldpcb200_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def kernel_info(self):
        v = [C.c_int() for _ in range(5)]
        _check(lib().ldpcb200_kernel_info(self._h, *[C.byref(x) for x in v]))
        d = dict(zip(("fast", "threads", "frames_per_cta", "ctas_per_sm", "smem_bytes"), [x.value for x in v]))
        d["tmem"] = bool(d["fast"] & 16)
        d["two_frames"] = bool(d["fast"] & 32)
        d["fast"] &= 15
        d["name"] = {0: "generic (table-driven, state in %s)" % ("shared memory" if d["smem_bytes"] else "an L2-resident workspace"), 1: "lms_fast_kernel (table-driven, shared memory)",
                     2: "%s (code-specialised, ahead of time)", 3: "%s (code-specialised, NVRTC)"}.get(d["fast"], "?")
        if d["fast"] == 1 and self.decoder_id == TASP_DEC:
            d["name"] = "tasp_fast_kernel (table-driven, double, lambda messages in tensor memory)"
        if d["fast"] == 1 and self.decoder_id == LCHE_DEC:
            d["name"] = "tasp_fast_kernel<LCHE> (table-driven, double, messages in tensor memory)"
        if d["fast"] == 1 and self.decoder_id == LMS_DEC and self.precision == 64:
            d["name"] = "tasp_fast_kernel<LMS double> (table-driven, messages in tensor memory)"
        if d["fast"] == 1 and self.decoder_id == MS_DEC and self.precision == 64:
            d["name"] = "ms64_fast_kernel (table-driven, double, messages in tensor memory)"
        if d["fast"] == 1 and self.decoder_id == IASP_DEC:
            d["name"] = "iasp_fast_kernel (table-driven, 12-bit fixed point, messages in tensor memory)"
        if d["fast"] == 1 and self.decoder_id in (BP_DEC, SP_DEC):
            four = d["threads"] > (self.Z + 31) // 32 * 32
            d["name"] = "%s<%s> (table-driven, double, messages in tensor memory%s)" % ("bpsp4_kernel" if four else "bpsp_fast_kernel", "SP" if self.decoder_id == SP_DEC else "BP",
                                                                                         ", four threads per check row" if four else "")
        if d["fast"] == 1 and self.decoder_id == ASP_DEC:
            d["name"] = "asp_fast_kernel (table-driven, double, messages in tensor memory)"
        if "%s" in d["name"]:
            d["name"] %= {LMS_DEC: ("lms_tmem2" if d["two_frames"] else "lms_tmem") if d["tmem"] else "lms_spec", MS_DEC: "ms_tmem<float>" if d["tmem"] else "ms_spec<float>", IMS_DEC: ("ims_h2 (frames as fp16 pairs, %d in flight per CTA)" % d["frames_per_cta"] if d["frames_per_cta"] >= 2 else "ms_tmem<int>") if d["tmem"] else "ms_spec<int>"}.get(self.decoder_id, "spec")
        return d

    def post_dtype(self):
        if self.decoder_id == IMS_DEC:
            return I16, np.int16
        if self.decoder_id == IASP_DEC:
            return U16, np.uint16
        return (F32, np.float32) if self.precision == 32 else (F64, np.float64)

    def decode(self, llr, maxiter, want_hard=True, want_post=False, want_aux=False, packed=False, no_early_exit=False,
               chain=False):
        """Decode llr[nf, N] (numpy float64/float32 host array).  Returns dict(hard, iters, post, aux)."""
        llr = np.ascontiguousarray(llr)
        assert llr.dtype in (np.float64, np.float32) and llr.ndim == 2 and llr.shape[1] == self.N, (llr.dtype, llr.shape)
        nf = llr.shape[0]
        flags = (HARD_PACKED if packed else 0) | (NO_EARLY_EXIT if no_early_exit else 0) | (BP_CHAIN_SYNDROME if chain else 0)
        hard = None
        if want_hard:
            hard = np.zeros((nf, self.nwords), np.uint32) if packed else np.zeros((nf, self.N), np.uint8)
        iters = np.zeros(nf, np.int32)
        pcode, ptype = self.post_dtype()
        post = np.zeros((nf, self.N), ptype) if want_post else None
        aux = np.zeros((nf, self.N), np.int16) if want_aux else None
        _check(lib().ldpcb200_decode_batch(self._h, _ptr(llr), F64 if llr.dtype == np.float64 else F32, nf, maxiter, flags,
                                           _ptr(hard), _ptr(iters), _ptr(post), pcode, _ptr(aux)))
        return dict(hard=hard, iters=iters, post=post, aux=aux)

    def decode_device(self, llr, maxiter, hard_words=None, iters=None, no_early_exit=False):
        """Decode device-resident LLRs (torch CUDA tensor [nf, N], float32/float64) into device outputs
        (optional torch tensors: hard_words int32 [nf, nwords], iters int32 [nf])."""
        import torch
        assert llr.is_cuda and llr.is_contiguous() and llr.dim() == 2 and llr.shape[1] == self.N
        dt = F64 if llr.dtype == torch.float64 else F32
        flags = LLR_ON_DEVICE | OUT_ON_DEVICE | HARD_PACKED | (NO_EARLY_EXIT if no_early_exit else 0)
        _check(lib().ldpcb200_decode_batch(self._h, _ptr(llr), dt, llr.shape[0], maxiter, flags, _ptr(hard_words),
                                           _ptr(iters), None, self.post_dtype()[0], None))

    def _sim(self, snr_db, n_frames, maxiter, modulation, punct, seed, stream, first_frame, flags, qam_T):
        sp = SimParams()
        sp.snr_db, sp.modulation, sp.punctured_blocks, sp.max_iterations = snr_db, modulation, punct, maxiter
        sp.seed, sp.stream, sp.first_frame, sp.n_frames, sp.flags, sp.qam_T = seed, stream, first_frame, n_frames, flags, qam_T
        return sp

    def simulate(self, snr_db, n_frames, maxiter, modulation=MOD_BPSK, punct=0, seed=1, stream=0, first_frame=0,
                 no_early_exit=False, want_per_frame=False, qam_T=26.0):
        """One round of bp_simulation's frame loop (bp_simulation.cpp:591-824) on device-generated noise."""
        sp = self._sim(snr_db, n_frames, maxiter, modulation, punct, seed, stream, first_frame,
                       NO_EARLY_EXIT if no_early_exit else 0, qam_T)
        out = Counters()
        pf = np.zeros(n_frames, np.uint32) if want_per_frame else None
        _check(lib().ldpcb200_simulate(self._h, C.byref(sp), C.byref(out), _ptr(pf)))
        d = out.as_dict()
        if want_per_frame:
            d["per_frame"] = pf
        return d

    def generate_llr(self, snr_db, n_frames, modulation=MOD_BPSK, punct=0, seed=1, stream=0, first_frame=0,
                     dtype=np.float32, out=None, qam_T=26.0):
        """The channel LLRs simulate() feeds the decoder (host array, or into a torch CUDA tensor `out`)."""
        flags = 0
        if out is None:
            out = np.zeros((n_frames, self.N), dtype)
            dt = F64 if dtype == np.float64 else F32
        else:
            import torch
            assert out.is_cuda and out.is_contiguous() and out.numel() == n_frames * self.N
            dt = F64 if out.dtype == torch.float64 else F32
            flags = OUT_ON_DEVICE
        sp = self._sim(snr_db, n_frames, 0, modulation, punct, seed, stream, first_frame, flags, qam_T)
        _check(lib().ldpcb200_generate_llr(self._h, C.byref(sp), _ptr(out), dt))
        return out

    def simulate_codes(self, hds, snr_db, n_frames, maxiter, modulation=MOD_BPSK, punct=0, seed=1, stream=0, first_frame=0,
                       no_early_exit=False, want_per_frame=False):
        """One simulate() round for many candidate matrices of this handle's shape in ONE launch (TASP_DEC; the search
        caller's loop, main_good_code_search.cpp:267-411).  hds: [K, b, c] -> list of K counter dicts."""
        hds = np.ascontiguousarray(hds, dtype=np.int16)
        assert hds.ndim == 3 and hds.shape[1:] == (self.b, self.c), hds.shape
        K = hds.shape[0]
        sp = self._sim(snr_db, n_frames, maxiter, modulation, punct, seed, stream, first_frame, NO_EARLY_EXIT if no_early_exit else 0, 26.0)
        out = (Counters * K)()
        pf = np.zeros((K, n_frames), np.uint32) if want_per_frame else None
        _check(lib().ldpcb200_simulate_codes(self._h, K, _ptr(hds), C.byref(sp), out, _ptr(pf)))
        res = [o.as_dict() for o in out]
        if want_per_frame:
            for k, d in enumerate(res):
                d["per_frame"] = pf[k]
        return res

    def set_codeword(self, bits=None):
        """The transmitted codeword of simulate() / generate_llr() (N bits 0 / 1); None = all-zero, as the reference sends."""
        if bits is None:
            _check(lib().ldpcb200_set_codeword(self._h, None))
            return
        b = np.ascontiguousarray(bits, dtype=np.uint8)
        assert b.shape == (self.N,)
        _check(lib().ldpcb200_set_codeword(self._h, _ptr(b)))

    def generate_noise(self, snr_db, n_frames, n_samples, modulation=MOD_BPSK, punct=0, seed=1, stream=0, first_frame=0):
        """The unit-variance noise samples the channel adds (fp32 [n_frames, n_samples])."""
        sp = self._sim(snr_db, n_frames, 0, modulation, punct, seed, stream, first_frame, 0, 26.0)
        out = np.zeros((n_frames, n_samples), np.float32)
        _check(lib().ldpcb200_generate_noise(self._h, C.byref(sp), n_samples, _ptr(out)))
        return out

    def set_interleaver(self, direct=None, inverse=None):
        """Attach (or with no arguments remove) a bit interleaver: simulate() / generate_llr() then feed decoder input i with the
        LLR received at transmitted position inverse[i] (bp_simulation.cpp:684)."""
        if direct is None and inverse is None:
            _check(lib().ldpcb200_set_interleaver(self._h, None, None))
            return
        d, i = np.ascontiguousarray(direct, dtype=np.int32), np.ascontiguousarray(inverse, dtype=np.int32)
        assert d.shape == (self.N,) and i.shape == (self.N,)
        _check(lib().ldpcb200_set_interleaver(self._h, _ptr(d), _ptr(i)))

    def last_kernel_ms(self):
        ms, n = C.c_float(), C.c_int()
        _check(lib().ldpcb200_last_kernel_ms(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def stream(self):
        return lib().ldpcb200_stream(self._h)


def interleaver_tables(hd, Z, modulation, mode, block=1, inter=1):
    """Index tables of the reference's bit interleaver modes 0-4 (direct_inverse_perm.cpp) -> (direct[N], inverse[N])."""
    hd = np.ascontiguousarray(hd, dtype=np.int16)
    N = hd.shape[1] * Z
    d, i = np.zeros(N, np.int32), np.zeros(N, np.int32)
    _check(lib().ldpcb200_interleaver_tables(_ptr(hd), hd.shape[0], hd.shape[1], Z, modulation, mode, block, inter, _ptr(d), _ptr(i)))
    return d, i


def girth_spectrum(hd, Z, gtarget=4):
    """Girth, ACE spectrum and cycle spectrum of a base matrix (host code; main_simulation.cpp:148-205) -> (girth, ace, spectrum)."""
    hd = np.ascontiguousarray(hd, dtype=np.int16)
    g = C.c_int()
    ace, spec = np.zeros(gtarget, np.int32), np.zeros(gtarget, np.int32)
    _check(lib().ldpcb200_girth_spectrum(_ptr(hd), hd.shape[0], hd.shape[1], Z, gtarget, C.byref(g), _ptr(ace), _ptr(spec)))
    return g.value, [int(x) for x in ace], [int(x) for x in spec]


def jit_check(hd, Z, sm=(10, 0)):
    """Generate + NVRTC-compile the code-specialised LMS_DEC kernel for a matrix (needs no GPU); -> cubin size."""
    hd = np.ascontiguousarray(hd, dtype=np.int16)
    n = C.c_int()
    _check(lib().ldpcb200_jit_check(_ptr(hd), hd.shape[0], hd.shape[1], Z, sm[0], sm[1], C.byref(n)))
    return n.value


def demodulate(Q, ns, sigma_, x, T=26.0, out_type=0, device=-1):
    """Demodulate() at the function boundary (QAM_demodulator.cpp:99-566), m = log2(Q)."""
    m = int(np.log2(Q))
    x = np.ascontiguousarray(x, np.float64)
    assert x.size == 2 * ns
    res = np.zeros(ns * m, np.float64)
    _check(lib().ldpcb200_demodulate(Q, ns, sigma_, T, out_type, _ptr(x), _ptr(res), device))
    return res


def modulate(Q, bits, device=-1):
    """QAM_modulator() (QAM_modulator.cpp:142-194): bits -> 2*ns lattice coordinates."""
    m = int(np.log2(Q))
    bits = np.ascontiguousarray(bits, np.uint8)
    ns = bits.size // m
    out = np.zeros(2 * ns, np.float64)
    _check(lib().ldpcb200_modulate(Q, ns, _ptr(bits), _ptr(out), device))
    return out
