"""The algebra behind the message computation of the tensor-memory min-sum kernels (csrc/lms_tmem.cuh min_of_others,
csrc/ms_tmem.cuh pass C), checked in numpy fp32 on the CPU: the kernels compute a message as g(min over the OTHER edges)
with the ceiling folded into a constant, offset / floor / sign as two fused multiply-adds, and IMS_DEC's shift as a
round-down fma -- each must equal the reference's two-smallest form (decoders.cpp:5163-5198, :5653) value for value."""
import numpy as np

f32 = np.float32
CAP = f32(32767.0) + f32(205.0) / f32(512.0)          # lms_tmem.cuh: 32767.400390625f


def lms_reference_form(c):
    """min(max(c - 0.4, 0), 32767) in fp32 (offset, floor, ceiling of lmin_sum_decod_qc_lm)."""
    return np.minimum(np.maximum((c - f32(0.4)).astype(f32), f32(0)), f32(32767))


def fma32(a, b, c):
    """fp32 fused multiply-add through float64: the product of two floats is exact in double, and so is the sum
    for the magnitudes used here (< 2^17 with steps >= 2^-30); one rounding to fp32."""
    return (a.astype(np.float64) * np.float64(b) + np.float64(c)).astype(f32)


def lms_kernel_form(c, s):
    """s * max(min(c, CAP) - 0.4, 0) as the kernel computes it: two FFMAs, s = +-1."""
    m = np.minimum(c, CAP)
    rhalf = f32(s) * f32(0.5)
    nhalf = rhalf * f32(-0.4)
    th = fma32(m, rhalf, nhalf)
    return fma32(np.abs(th), f32(s), th)


def test_folded_ceiling_is_the_smallest_float_that_reaches_32767():
    assert float(CAP) == 32767.400390625
    assert (CAP - f32(0.4)).astype(f32) == f32(32767.0)
    below = np.nextafter(CAP, f32(0))
    assert (below - f32(0.4)).astype(f32) < f32(32767.0)


def test_lms_message_magnitude_forms_agree():
    rng = np.random.default_rng(5)
    # every float from 32766 to 32769 (step 2^-9), the region around the offset, and a broad random sample
    around_cap = np.arange(32766 * 512, 32769 * 512 + 1, dtype=np.int64).astype(np.float64) / 512.0
    c = np.concatenate([around_cap.astype(f32),
                        np.linspace(0, 1, 200001).astype(f32),
                        np.abs(rng.normal(0, 30, 400000)).astype(f32),
                        np.abs(rng.normal(0, 1e5, 100000)).astype(f32),
                        np.array([0.0, 0.4, np.nextafter(f32(0.4), f32(1)), np.nextafter(f32(0.4), f32(0)), 1e30, 3.4e38], f32)])
    want = lms_reference_form(c)
    for s in (1.0, -1.0):
        got = lms_kernel_form(c, s)
        assert np.array_equal(np.abs(got), want)
        assert np.all((got == 0) | (np.sign(got) == s))


def test_min_over_others_equals_two_smallest_selection():
    rng = np.random.default_rng(6)
    for deg in (1, 2, 3, 4, 7, 8, 13, 19):
        v = rng.normal(0, 3, (2000, deg)).astype(f32)
        v[::7, 0] = v[::7, -1]                                      # ties between the two smallest
        a = np.abs(v)
        order = np.sort(a, axis=1)
        c1 = order[:, 0]
        c2 = order[:, 1] if deg > 1 else np.full(len(v), np.inf, f32)
        first = np.argmin(a, axis=1)                                 # the reference's position of the first minimum
        want = np.where(np.arange(deg)[None, :] == first[:, None], c2[:, None], c1[:, None])
        got = np.stack([np.min(np.delete(a, q, axis=1), axis=1, initial=np.inf) for q in range(deg)], axis=1)
        assert np.array_equal(want, got)


def test_ims_shift_as_round_down_fma():
    """(min(c, max_data) * ialpha) >> 4 on integers carried as floats: fma(c, ialpha / 16, 1.5 * 2^23) rounded down."""
    magic = np.float64(12582912.0)
    for ialpha in (1, 7, 12, 13, 16, 31):
        c = np.arange(0, 4096, dtype=np.int64)
        want = (c * ialpha) >> 4
        exact = c.astype(np.float64) * (ialpha * 0.0625) + magic     # exact in double
        t = np.floor(exact)                                          # fp32 grid up there is the integers: round-down = floor
        assert np.all(t < 2 ** 24)
        got = (t.astype(f32) - f32(magic)).astype(np.int64)
        assert np.array_equal(got, want)


def grouped_min_of_others(a, cap):
    """The grouping csrc/lms_tmem.cuh min_of_others uses (groups of three, per-group rest term with the ceiling folded
    in, prefix / suffix minima from four groups on), restated on numpy columns."""
    deg = a.shape[1]
    G = (deg + 2) // 3
    g = [np.min(a[:, 3 * k:min(3 * k + 3, deg)], axis=1) for k in range(G)]
    capv = np.full(len(a), cap, f32)
    if G == 1:
        rest = [capv]
    elif G == 2:
        rest = [np.minimum(g[1], capv), np.minimum(g[0], capv)]
    elif G == 3:
        rest = [np.minimum(np.minimum(g[1], g[2]), capv), np.minimum(np.minimum(g[0], g[2]), capv), np.minimum(np.minimum(g[0], g[1]), capv)]
    else:
        pre, suf = [capv], [None] * G
        for k in range(1, G):
            pre.append(np.minimum(pre[k - 1], g[k - 1]))
        suf[G - 2] = g[G - 1]
        for k in range(G - 3, -1, -1):
            suf[k] = np.minimum(suf[k + 1], g[k + 1])
        rest = [np.minimum(pre[k], suf[k]) for k in range(G - 1)] + [pre[G - 1]]
    m = np.empty_like(a)
    for k in range(G):
        lo, hi = 3 * k, min(3 * k + 3, deg)
        for q in range(lo, hi):
            mates = [p for p in range(lo, hi) if p != q]
            v = rest[k]
            for p in mates:
                v = np.minimum(v, a[:, p])
            m[:, q] = v
    return m


def test_grouped_minima_cover_every_other_edge_exactly_once():
    rng = np.random.default_rng(8)
    for deg in range(1, 25):
        a = np.abs(rng.normal(0, 3, (500, deg))).astype(f32)
        a[::5, deg // 2] = a[::5, 0]                                 # ties
        a[::11] *= f32(2e4)                                          # rows beyond the ceiling
        for cap in (f32(32767.0), CAP, f32(127.0)):
            want = np.stack([np.minimum(np.min(np.delete(a, q, axis=1), axis=1, initial=np.inf), cap) for q in range(deg)], axis=1).astype(f32)
            assert np.array_equal(grouped_min_of_others(a, cap), want), (deg, cap)


def test_ims_shift_as_two_binary16_fmas():
    """csrc/ims_h2.cuh: floor(m * a / 16) = fma(m + d, a / 16, 1025) - 1025 in binary16 round-to-nearest, d = -k / 16 with
    112 / a < k < 128 / a (none for a = 0, 16), for every m <= 127 and every ialpha 0 .. 16; the sign step
    fma(t, +-1, -+1025) then gives +-floor exactly (decoders.cpp:5554, :5640)."""
    f16 = np.float16
    m = np.arange(0, 128, dtype=np.float64)
    for a in range(0, 17):
        k = 0 if a in (0, 16) else 112 // a + 1
        if k:
            assert 112 / a < k < 128 / a
        d = -k / 16.0
        md = (m + d).astype(f16)
        assert np.array_equal(md.astype(np.float64), m + d)                    # m + d is exact on the half grid
        t = (md.astype(np.float64) * (a / 16.0) + 1025.0).astype(f16)          # one rounding: the fma
        want = np.floor(m * a / 16.0)
        assert np.array_equal(t.astype(np.float64) - 1025.0, want), a
        for s in (1.0, -1.0):
            q = (t.astype(np.float64) * s - 1025.0 * s).astype(f16)
            assert np.array_equal(q.astype(np.float64), s * want)
    # every quantity of IMS_DEC with dbits <= 8 is an integer binary16 holds exactly
    ints = np.arange(-2048, 2049)
    assert np.array_equal(ints.astype(f16).astype(np.int64), ints)
