// Table-driven parity kernels of the sum-product family, all in double like the reference:
//
//   TASP_DEC  tdmp_sum_prod_gf2_decod_qc_lm   decoders.cpp:2584-2744   layered, probability domain
//   ASP_DEC   sum_prod_gf2_decod_qc_lm        decoders.cpp:2324-2581   flooding, probability domain
//   BP_DEC    bp_decod_qc_lm                  decoders.cpp:1708-1920   flooding, log-tanh domain
//   SP_DEC    sum_prod_decod_qc_lm            decoders.cpp:1923-2185   flooding, likelihood ratios
//   LCHE_DEC  lche_decod                      decoders.cpp:2893-3010   layered, table look-ups
//   IASP_DEC  isum_prod_gf2_decod_qc_lm       decoders.cpp:3822-4121   flooding, 12-bit fixed point
//
// One frame per CTA at a time, persistent grid, state in the CTA's global workspace slice.  Messages
// are stored per edge and lane ([e*Z + n], n = check-row lane) instead of the reference's dense
// [row][column] arrays; every floating-point sum/product runs in the reference's order (compiled with
// -fmad=false), so only exp()/log() can differ from the host libm, in the last ulp.
#include "dec_common.cuh"

namespace ldpcb200 {

#define MAXD LDPCB200_MAX_ROW_WEIGHT

// mind / maxd / absd of decoders.cpp:104-109 (the comparison direction matters for NaN operands)
__device__ __forceinline__ double mind(double a, double b) { return a < b ? a : b; }
__device__ __forceinline__ double maxd(double a, double b) { return a < b ? b : a; }
__device__ __forceinline__ double absd(double x) { return x < 0 ? -x : x; }

template <class Pred>
__device__ __forceinline__ int syndrome_pred(const QcDev& g, Pred is_one)
{
    int bad = 0;
    for (int r = threadIdx.x; r < g.R; r += blockDim.x) {
        int j = r / g.Z, n = r - j * g.Z, s = 0;
        for (int e = g.rp[j]; e < g.rp[j + 1]; e++)
            s ^= is_one(g.col[e] * g.Z + wrapz(n + g.sh[e], g.Z));
        bad |= s;
    }
    return __syncthreads_or(bad);
}

// LLR -> P(bit = 1), decoders.cpp:2611-2618 (= :2351-2358)
__device__ __forceinline__ double llr_to_p1(double llr)
{
    double x = llr * 0.5;
    double yv = maxd(mind(x, 20.0), -20.0);
    double e0 = exp(yv);
    double e1 = exp(-yv);
    return e1 / (e0 + e1);
}

// map_bin, decoders.cpp:2191-2228, on rw >= 2 values
__device__ __forceinline__ void map_bin(double* s, int rw)
{
    double SF[MAXD], SB[MAXD], P[MAXD];
    for (int i = 0; i < rw; i++) P[i] = 1 - 2 * s[i];
    SF[0] = P[0];
    for (int i = 1; i < rw - 1; i++) SF[i] = P[i] * SF[i - 1];
    SB[rw - 1] = P[rw - 1];
    for (int i = rw - 2; i > 0; i--) SB[i] = P[i] * SB[i + 1];
    s[0] = (1 - SB[1]) / 2;
    for (int i = 1; i < rw - 1; i++) {
        double Zv = SF[i - 1] * SB[i + 1];
        s[i] = (1 - Zv) / 2;
    }
    s[rw - 1] = (1 - SF[rw - 2]) / 2;
}

// ------------------------------------------------------------------------------------------------
struct TaspGeneric {
    static size_t ws_bytes(const QcHost& g, int) { return carve_bytes(g.N, 8) + carve_bytes((size_t)g.E * g.Z, 8); }
    static __device__ void frame(const QcDev& g, const DecParams&, const FrameIO& io, int f, char* ws)
    {
        const int Z = g.Z, N = g.N, nt = blockDim.x, tid = threadIdx.x;
        double* gam = carve<double>(ws, N);
        double* lam = carve<double>(ws, (size_t)g.E * Z);
        const double T = 0.0001, TT = 0;                                         // :2597-2598
        for (int i = tid; i < N; i += nt) gam[i] = llr_to_p1(load_llr(io, N, f, i));
        for (int i = tid; i < g.E * Z; i += nt) lam[i] = 0.5;                    // :2620-2641
        __syncthreads();
        int synd = syndrome_pred(g, [&](int i) { return (int)(gam[i] > 0.5); }); // :2653
        int ret = 0, locked = 0, steps = 0;
        const bool noexit = io.flags & LDPCB200_NO_EARLY_EXIT;
        if (!synd) { locked = 1; ret = 0; }                                      // :2654-2660
        if (synd || noexit) {
            while (steps < io.maxiter) {
                for (int j = 0; j < g.b; j++) {
                    const int e0 = g.rp[j], cnt = g.rp[j + 1] - e0;
                    for (int n = tid; n < Z; n += nt) {
                        double rho[MAXD], a[MAXD];
                        for (int q = 0; q < cnt; q++) {
                            const int e = e0 + q;
                            double x = gam[g.col[e] * Z + wrapz(n + g.sh[e], Z)];
                            double av = lam[(size_t)e * Z + n];
                            double r = x * (1.0 - av) / (av + x - 2.0 * av * x);  // :2686
                            if (r < TT) r = TT;                                   // :2692-2697
                            if (r > 1 - TT) r = 1 - TT;
                            rho[q] = r; a[q] = r;
                        }
                        map_bin(a, cnt);                                          // :2699
                        for (int q = 0; q < cnt; q++) {
                            const int e = e0 + q;
                            double av = a[q];
                            if (av < T) av = T;                                   // :2701-2705
                            if (av > 1.0 - T) av = 1.0 - T;
                            gam[g.col[e] * Z + wrapz(n + g.sh[e], Z)] =
                                rho[q] * av / (1.0 - rho[q] - av + 2 * rho[q] * av);   // :2716
                            lam[(size_t)e * Z + n] = av;
                        }
                    }
                    __syncthreads();
                }
                // the reference re-checks after every layer (:2723); only the last verdict is used (:2733)
                synd = syndrome_pred(g, [&](int i) { return (int)(gam[i] > 0.5); });
                steps++;
                if (!synd) { if (!locked) { ret = steps; locked = 1; } if (!noexit) break; }
            }
        }
        if (!locked) ret = synd ? -steps : steps;                                // :2740-2743
        for (int i = tid; i < N; i += nt) store_post(io, N, f, i, gam[i]);
        emit_frame(g, io, f, ret, [&](int i) { return (int)(gam[i] > 0.5); });   // :2738
    }
};

// ------------------------------------------------------------------------------------------------
struct AspGeneric {
    static size_t ws_bytes(const QcHost& g, int) { return 2 * carve_bytes(g.N, 8) + carve_bytes((size_t)g.E * g.Z, 8); }
    static __device__ void frame(const QcDev& g, const DecParams&, const FrameIO& io, int f, char* ws)
    {
        const int Z = g.Z, N = g.N, R = g.R, nt = blockDim.x, tid = threadIdx.x;
        double* p = carve<double>(ws, N);
        double* so = carve<double>(ws, N);
        double* msg = carve<double>(ws, (size_t)g.E * Z);
        for (int i = tid; i < N; i += nt) { p[i] = llr_to_p1(load_llr(io, N, f, i)); so[i] = p[i]; }   // :2351-2358
        __syncthreads();
        for (int x = tid; x < g.E * Z; x += nt) {                                // :2361-2378
            int e = x / Z, n = x - e * Z;
            msg[x] = p[g.col[e] * Z + wrapz(n + g.sh[e], Z)];
        }
        __syncthreads();
        int synd = syndrome_pred(g, [&](int i) { return (int)(so[i] > 0.5); });  // :2392
        int ret = 0, locked = 0, steps = 0;
        const bool noexit = io.flags & LDPCB200_NO_EARLY_EXIT;
        if (!synd) { locked = 1; ret = 0; }
        if (synd || noexit) {
            while (steps < io.maxiter) {
                for (int r = tid; r < R; r += nt) {                              // check nodes, :2406-2428
                    const int j = r / Z, n = r - j * Z, e0 = g.rp[j], cnt = g.rp[j + 1] - e0;
                    double a[MAXD];
                    for (int q = 0; q < cnt; q++) a[q] = msg[(size_t)(e0 + q) * Z + n];
                    map_bin(a, cnt);
                    for (int q = 0; q < cnt; q++) msg[(size_t)(e0 + q) * Z + n] = a[q];
                }
                __syncthreads();
                if (g.all_cw_2) {                                                // :2432-2482
                    for (int v = tid; v < N; v += nt) {
                        const int i = v / Z, k = v - i * Z;
                        const int ea = g.cedge[g.cp[i]], eb = g.cedge[g.cp[i] + 1];
                        const size_t ma = (size_t)ea * Z + wrapz(k - g.sh[ea] + Z, Z);
                        const size_t mb = (size_t)eb * Z + wrapz(k - g.sh[eb] + Z, Z);
                        double d0 = msg[ma], d1 = msg[mb];
                        double p1 = p[v];
                        double q10 = p1, q11 = p1;
                        double q00 = 1.0 - p1, q01 = 1.0 - p1, p0;
                        q10 = q10 * d1;
                        q00 = q00 * (1 - d1);
                        q11 = q11 * d0;
                        q01 = q01 * (1 - d0);
                        p1 = q10 * d0;
                        p0 = q00 * (1 - d0);
                        so[v] = div_normal(p1, p0 + p1);                          // operands are normal numbers: dec_common.cuh
                        msg[ma] = div_normal(q10, q10 + q00);
                        msg[mb] = div_normal(q11, q11 + q01);
                    }
                    __syncthreads();
                } else {
                    for (int v = tid; v < N; v += nt) {                          // overall products, :2489-2522
                        const int i = v / Z, k = v - i * Z;
                        double P1 = p[v];
                        double P0 = 1 - p[v];
                        for (int q = g.cp[i]; q < g.cp[i + 1]; q++) {
                            const int e = g.cedge[q];
                            double d = msg[(size_t)e * Z + wrapz(k - g.sh[e] + Z, Z)];
                            P1 *= d;
                            P0 *= 1 - d;
                        }
                        so[v] = div_normal(P1, P0 + P1);
                    }
                    __syncthreads();
                    for (int x = tid; x < g.E * Z; x += nt) {                    // local data updating, :2525-2558
                        const int e = x / Z, n = x - e * Z;
                        double s1 = so[g.col[e] * Z + wrapz(n + g.sh[e], Z)];
                        double sos = msg[x];
                        double p1 = div_normal(s1, sos);
                        double p0 = div_normal(1 - s1, 1 - sos);
                        double d = div_normal(p1, p1 + p0);
                        msg[x] = maxd(mind(d, 1.0 - 0.000001), 0.000001);        // SP_DEC_MIN/MAX_VAL, :96-97
                    }
                    __syncthreads();
                }
                synd = syndrome_pred(g, [&](int i) { return (int)(so[i] > 0.5); });   // :2566
                steps++;
                if (!synd) { if (!locked) { ret = steps; locked = 1; } if (!noexit) break; }
            }
        }
        if (!locked) ret = -steps;
        for (int i = tid; i < N; i += nt) store_post(io, N, f, i, so[i]);
        emit_frame(g, io, f, ret, [&](int i) { return (int)(so[i] > 0.5); });    // make_output :2308
    }
};

// ------------------------------------------------------------------------------------------------
// ZZ / BB are stored per edge at the VARIABLE position k of the block column, as the reference does;
// the check-row lane of (edge, k) is n = (k - shift) mod Z.
struct BpGeneric {
    static size_t ws_bytes(const QcHost& g, int)
    {
        return 2 * carve_bytes(g.N, 8) + carve_bytes((size_t)g.E * g.Z, 8) + carve_bytes((size_t)g.E * g.Z, 1) +
               carve_bytes(g.R, 8) + 2 * carve_bytes(g.R, 1);
    }
    static __device__ void frame(const QcDev& g, const DecParams&, const FrameIO& io, int f, char* ws)
    {
        const int Z = g.Z, N = g.N, R = g.R, nt = blockDim.x, tid = threadIdx.x;
        double* soft = carve<double>(ws, N);
        double* yd = carve<double>(ws, N);
        double* ZZ = carve<double>(ws, (size_t)g.E * Z);
        uint8_t* BB = carve<uint8_t>(ws, (size_t)g.E * Z);
        double* s = carve<double>(ws, R);
        uint8_t* bs = carve<uint8_t>(ws, R);
        uint8_t* syndr_local = carve<uint8_t>(ws, R);
        const bool chain = (io.flags & LDPCB200_BP_CHAIN_SYNDROME) && io.bp_syndrome;
        uint8_t* syndr = chain ? io.bp_syndrome : syndr_local;

        for (int i = tid; i < g.E * Z; i += nt) ZZ[i] = 0.0;                     // :1732-1734
        for (int i = tid; i < N; i += nt) yd[i] = soft[i] = maxd(mind(load_llr(io, N, f, i), 20.0), -20.0);   // :1738
        if (!chain) for (int i = tid; i < R; i += nt) syndr[i] = 0;
        __syncthreads();
        // pre-iteration check XORs into the syndrome left by the previous frame, :1742-1759
        int bad = 0;
        for (int r = tid; r < R; r += nt) {
            const int j = r / Z, n = r - j * Z;
            int sy = syndr[r];
            for (int e = g.rp[j]; e < g.rp[j + 1]; e++) sy ^= soft[g.col[e] * Z + wrapz(n + g.sh[e], Z)] < 0;
            syndr[r] = (uint8_t)sy;
            bad |= sy;
        }
        int synd = __syncthreads_or(bad);
        int ret = 0, locked = 0, iter = 0;
        const bool noexit = io.flags & LDPCB200_NO_EARLY_EXIT;
        if (!synd) { locked = 1; ret = 0; }                                      // :1765-1779
        if (synd || noexit) {
            while (iter < io.maxiter) {
                // variable-node step, :1790-1812
                for (int x = tid; x < g.E * Z; x += nt) {
                    const int e = x / Z, k = x - e * Z;
                    double A = exp(soft[g.col[e] * Z + k] - ZZ[x]);
                    double xv = log(absd((A - 1) / (A + 1)));
                    BB[x] = A < 1;
                    ZZ[x] = xv;
                }
                __syncthreads();
                // check sums, ascending block column (:1814-1827)
                for (int r = tid; r < R; r += nt) {
                    const int j = r / Z, n = r - j * Z;
                    double sv = 0;
                    int b1 = 0;
                    for (int e = g.rp[j]; e < g.rp[j + 1]; e++) {
                        const size_t x = (size_t)e * Z + wrapz(n + g.sh[e], Z);
                        sv += ZZ[x];
                        b1 ^= BB[x];
                    }
                    s[r] = sv; bs[r] = (uint8_t)b1;
                }
                __syncthreads();
                // check-node step + posterior, :1832-1862, ascending block row per variable
                for (int v = tid; v < N; v += nt) {
                    const int i = v / Z, k = v - i * Z;
                    double acc = yd[v];
                    for (int q = g.cp[i]; q < g.cp[i + 1]; q++) {
                        const int e = g.cedge[q], j = g.row[e];
                        const int n = wrapz(k - g.sh[e] + Z, Z);
                        const size_t x = (size_t)e * Z + k;
                        double A = exp(s[j * Z + n] - ZZ[x]);
                        int bb = bs[j * Z + n] ^ BB[x];
                        A = (1 - 2 * bb) * log((1 + A) / (1 - A));
                        A = maxd(mind(A, 19.07), -19.07);
                        ZZ[x] = A;
                        acc += A;
                    }
                    soft[v] = acc;
                }
                __syncthreads();
                bad = 0;                                                         // :1865-1882
                for (int r = tid; r < R; r += nt) {
                    const int j = r / Z, n = r - j * Z;
                    int sy = 0;
                    for (int e = g.rp[j]; e < g.rp[j + 1]; e++) sy ^= soft[g.col[e] * Z + wrapz(n + g.sh[e], Z)] < 0;
                    syndr[r] = (uint8_t)sy;
                    bad |= sy;
                }
                synd = __syncthreads_or(bad);
                iter++;
                if (!synd) { if (!locked) { ret = iter; locked = 1; } if (!noexit) break; }
            }
        }
        if (!locked) ret = -iter;
        for (int i = tid; i < N; i += nt) store_post(io, N, f, i, soft[i]);
        emit_frame(g, io, f, ret, [&](int i) { return (int)(soft[i] < 0); });
    }
};

// ------------------------------------------------------------------------------------------------
struct SpGeneric {
    static size_t ws_bytes(const QcHost& g, int)
    {
        return 2 * carve_bytes(g.N, 8) + 2 * carve_bytes((size_t)g.E * g.Z, 8) + carve_bytes(g.R, 8);
    }
    static __device__ void frame(const QcDev& g, const DecParams&, const FrameIO& io, int f, char* ws)
    {
        const int Z = g.Z, N = g.N, R = g.R, nt = blockDim.x, tid = threadIdx.x;
        double* soft = carve<double>(ws, N);
        double* yd = carve<double>(ws, N);
        double* ZZ = carve<double>(ws, (size_t)g.E * Z);     // per edge at the variable position k
        double* ZZ0 = carve<double>(ws, (size_t)g.E * Z);
        double* s = carve<double>(ws, R);
        const double SP_THR = 1.0;                                               // :1922
        for (int i = tid; i < N; i += nt) {
            double v = maxd(mind(load_llr(io, N, f, i), 20.0), -20.0);
            yd[i] = soft[i] = exp(v);                                            // :1947-1951
        }
        for (int i = tid; i < g.E * Z; i += nt) ZZ[i] = 1.0;                     // :1957-1959
        __syncthreads();
        int synd = syndrome_pred(g, [&](int i) { return (int)(soft[i] < SP_THR); });   // :1964-1987
        int ret = 0, locked = 0, iter = 0;
        const bool noexit = io.flags & LDPCB200_NO_EARLY_EXIT;
        if (!synd) { locked = 1; ret = 0; }
        if (synd || noexit) {
            while (iter < io.maxiter) {
                // :2013-2061: extrinsic products over the column (O(dv^2)), ascending block row
                for (int v = tid; v < N; v += nt) {
                    const int i = v / Z, k = v - i * Z;
                    for (int q = g.cp[i]; q < g.cp[i + 1]; q++) {
                        double AA = yd[v];
                        for (int q2 = g.cp[i]; q2 < g.cp[i + 1]; q2++) {
                            if (q2 == q) continue;
                            AA *= ZZ[(size_t)g.cedge[q2] * Z + k];
                        }
                        ZZ0[(size_t)g.cedge[q] * Z + k] = (AA - 1) / (AA + 1);
                    }
                }
                __syncthreads();
                for (int r = tid; r < R; r += nt) {
                    const int j = r / Z, n = r - j * Z;
                    double sv = 1.0;
                    for (int e = g.rp[j]; e < g.rp[j + 1]; e++) sv *= ZZ0[(size_t)e * Z + wrapz(n + g.sh[e], Z)];
                    s[r] = sv;
                }
                __syncthreads();
                for (int v = tid; v < N; v += nt) {                              // :2103-2127
                    const int i = v / Z, k = v - i * Z;
                    double acc = yd[v];
                    for (int q = g.cp[i]; q < g.cp[i + 1]; q++) {
                        const int e = g.cedge[q], j = g.row[e];
                        const int n = wrapz(k - g.sh[e] + Z, Z);
                        double A = s[j * Z + n] / ZZ0[(size_t)e * Z + k];
                        A = (1 + A) / (1 - A);
                        A = maxd(mind(A, 1.9e+8), -5.2e-9);
                        ZZ[(size_t)e * Z + k] = A;
                        acc *= A;
                    }
                    soft[v] = acc;
                }
                __syncthreads();
                synd = syndrome_pred(g, [&](int i) { return (int)(soft[i] < SP_THR); });   // :2129-2149
                iter++;
                if (!synd) { if (!locked) { ret = iter; locked = 1; } if (!noexit) break; }
            }
        }
        if (!locked) ret = -iter;
        for (int i = tid; i < N; i += nt) store_post(io, N, f, i, soft[i]);
        emit_frame(g, io, f, ret, [&](int i) { return (int)(soft[i] < 1.0); });
    }
};

// ------------------------------------------------------------------------------------------------
// logexp_int tables, decoders.cpp:2748-2775
__constant__ double LCHE_A[32] = {
    1.41e+00, 7.72e-01, 4.54e-01, 2.72e-01, 1.65e-01, 9.97e-02, 6.04e-02, 3.66e-02,
    2.22e-02, 1.35e-02, 8.17e-03, 4.96e-03, 3.01e-03, 1.82e-03, 1.11e-03, 6.71e-04,
    4.07e-04, 2.47e-04, 1.50e-04, 9.08e-05, 5.51e-05, 3.34e-05, 2.03e-05, 1.23e-05,
    7.45e-06, 4.52e-06, 2.74e-06, 1.66e-06, 1.01e-06, 6.12e-07, 3.71e-07, 2.25e-07 };
__constant__ double LCHE_B[32] = {
    3.47, 2.77, 2.37, 2.08, 1.86, 1.69, 1.54, 1.41, 1.29, 1.19, 1.11, 1.03, 0.95, 0.89, 0.83, 0.77,
    0.72, 0.67, 0.63, 0.59, 0.55, 0.52, 0.48, 0.45, 0.43, 0.40, 0.37, 0.35, 0.33, 0.31, 0.29, 0.27 };
__constant__ double LCHE_C[32] = {
    6.93, 6.24, 5.83, 5.55, 5.32, 5.14, 4.99, 4.85, 4.73, 4.63, 4.53, 4.45, 4.37, 4.29, 4.22, 4.16,
    4.10, 4.04, 3.99, 3.94, 3.89, 3.84, 3.80, 3.75, 3.71, 3.67, 3.64, 3.60, 3.56, 3.53, 3.50, 3.47 };

__device__ __forceinline__ double logexp_int(double x)                           // :2777-2836
{
    if (x <= 0) x = 1.0 / 4096.0;
    if (x > 16.0) x = 16.0;
    if (x >= 2.0) return -LCHE_A[(int)(2 * x + 0.5) - 1];
    else if (x > 1.0 / 16.0) return -LCHE_B[(int)(16 * x + 0.5) - 1];
    else if (x > 1.0 / 512.0) return -LCHE_C[(int)(512 * x + 0.5) - 1];
    else {
        double s = 0;
        while (x < 1.0 / 512.0) { x *= 32; s -= 3.46; }
        return s - LCHE_C[(int)(512 * x + 0.5) - 1];
    }
}

struct LcheGeneric {
    static size_t ws_bytes(const QcHost& g, int) { return carve_bytes(g.N, 8) + carve_bytes((size_t)g.E * g.Z, 8); }
    static __device__ void frame(const QcDev& g, const DecParams&, const FrameIO& io, int f, char* ws)
    {
        const int Z = g.Z, N = g.N, nt = blockDim.x, tid = threadIdx.x;
        double* so = carve<double>(ws, N);
        double* st = carve<double>(ws, (size_t)g.E * Z);
        for (int i = tid; i < g.E * Z; i += nt) st[i] = 0.0;                     // :2913-2915
        for (int i = tid; i < N; i += nt) so[i] = load_llr(io, N, f, i);
        __syncthreads();
        int synd = syndrome_pred(g, [&](int i) { return (int)(so[i] < 0); });    // :2927-2931
        int ret = 0, locked = 0, steps = 0;
        const bool noexit = io.flags & LDPCB200_NO_EARLY_EXIT;
        if (!synd) { locked = 1; ret = 0; }
        if (synd || noexit) {
            while (steps < io.maxiter) {
                for (int j = 0; j < g.b; j++) {
                    const int e0 = g.rp[j], cnt = g.rp[j + 1] - e0;
                    for (int n = tid; n < Z; n += nt) {
                        double u[MAXD], yv[MAXD], alog[MAXD];
                        int sy = 0;
                        double sum = 0;
                        for (int q = 0; q < cnt; q++) {
                            const int e = e0 + q;
                            u[q] = yv[q] = so[g.col[e] * Z + wrapz(n + g.sh[e], Z)] - st[(size_t)e * Z + n];   // :2962
                        }
                        // map_bin_llr, :2837-2890
                        for (int q = 0; q < cnt; q++) sy ^= u[q] < 0;
                        for (int q = 0; q < cnt; q++) {
                            double ay = u[q] < 0.0 ? -u[q] : u[q];
                            alog[q] = logexp_int(ay);
                        }
                        for (int q = 0; q < cnt; q++) sum += alog[q];
                        for (int q = 0; q < cnt; q++) {
                            int hardb = (u[q] < 0) ^ sy;
                            double A = alog[q] - sum;
                            double av = logexp_int(A);
                            u[q] = (2 * hardb - 1) * av;
                        }
                        for (int q = 0; q < cnt; q++) {
                            const int e = e0 + q;
                            so[g.col[e] * Z + wrapz(n + g.sh[e], Z)] = u[q] + yv[q];   // :2979
                            st[(size_t)e * Z + n] = u[q];
                        }
                    }
                    __syncthreads();
                }
                steps++;
                synd = syndrome_pred(g, [&](int i) { return (int)(so[i] < 0); });
                if (!synd) { if (!locked) { ret = steps; locked = 1; } if (!noexit) break; }
            }
        }
        if (!locked) ret = synd ? -steps : steps;                                // :3006-3009
        for (int i = tid; i < N; i += nt) store_post(io, N, f, i, so[i]);
        emit_frame(g, io, f, ret, [&](int i) { return (int)(so[i] < 0.0); });
    }
};

// ------------------------------------------------------------------------------------------------
// IASP_DEC: SOFT_FPP = 12, ONE_SOFT = 4096 (decoders.cpp:82-84); div_power2r (:80)
#define ONE_SOFT 4096
#define MAX_SOFT 4095
#define DIVR(x, n) (((x) + (1 << ((n) - 1))) >> (n))

__device__ __forceinline__ void imap_bin(uint16_t* s, int rw)                    // :2235-2271
{
    int16_t SF[MAXD], SB[MAXD], P[MAXD];
    for (int i = 0; i < rw; i++) P[i] = (int16_t)(ONE_SOFT - 2 * s[i]);
    SF[0] = P[0];
    for (int i = 1; i < rw - 1; i++) SF[i] = (int16_t)DIVR((int)P[i] * SF[i - 1], 12);
    SB[rw - 1] = P[rw - 1];
    for (int i = rw - 2; i > 0; i--) SB[i] = (int16_t)DIVR((int)P[i] * SB[i + 1], 12);
    s[0] = (uint16_t)DIVR(ONE_SOFT - SB[1], 1);
    s[0] = s[0] < 1 ? 1 : s[0];
    for (int i = 1; i < rw - 1; i++) {
        int Zv = DIVR((int)SF[i - 1] * SB[i + 1], 12);
        s[i] = (uint16_t)DIVR(ONE_SOFT - Zv, 1);
        s[i] = s[i] < 1 ? 1 : s[i];
    }
    s[rw - 1] = (uint16_t)DIVR(ONE_SOFT - SF[rw - 2], 1);
    s[rw - 1] = s[rw - 1] < 1 ? 1 : s[rw - 1];
}

struct IaspGeneric {
    static size_t ws_bytes(const QcHost& g, int) { return 2 * carve_bytes(g.N, 2) + carve_bytes((size_t)g.E * g.Z, 2); }
    static __device__ void frame(const QcDev& g, const DecParams&, const FrameIO& io, int f, char* ws)
    {
        const int Z = g.Z, N = g.N, R = g.R, nt = blockDim.x, tid = threadIdx.x;
        uint16_t* yq = carve<uint16_t>(ws, N);
        uint16_t* so = carve<uint16_t>(ws, N);
        uint16_t* msg = carve<uint16_t>(ws, (size_t)g.E * Z);
        for (int i = tid; i < N; i += nt) {                                      // :3849-3861
            double v = maxd(mind(load_llr(io, N, f, i), 20.0), -20.0);
            double pr = 1.0 / (1.0 + exp(v));
            int x = (int)(pr * ONE_SOFT + 0.5);
            x = MAX_SOFT < x ? MAX_SOFT : x;
            yq[i] = (uint16_t)(x < 1 ? 1 : x);
        }
        __syncthreads();
        for (int x = tid; x < g.E * Z; x += nt) {                                // :3869-3886 (12-bit values)
            int e = x / Z, n = x - e * Z;
            msg[x] = yq[g.col[e] * Z + wrapz(n + g.sh[e], Z)];
        }
        __syncthreads();
        for (int i = tid; i < N; i += nt) { yq[i] = (uint16_t)(yq[i] << 4); so[i] = yq[i]; }   // :3867, :3889
        __syncthreads();
        int synd = syndrome_pred(g, [&](int i) { return (int)(so[i] >> 15); });
        int ret = 0, locked = 0, steps = 0;
        const bool noexit = io.flags & LDPCB200_NO_EARLY_EXIT;
        if (!synd) { locked = 1; ret = 0; }
        if (synd || noexit) {
            while (steps < io.maxiter) {
                for (int r = tid; r < R; r += nt) {                              // :3906-3911
                    const int j = r / Z, n = r - j * Z, e0 = g.rp[j], cnt = g.rp[j + 1] - e0;
                    uint16_t a[MAXD];
                    for (int q = 0; q < cnt; q++) a[q] = msg[(size_t)(e0 + q) * Z + n];
                    imap_bin(a, cnt);
                    for (int q = 0; q < cnt; q++) msg[(size_t)(e0 + q) * Z + n] = a[q];
                }
                __syncthreads();
                if (g.all_cw_2) {                                                // :3915-3977
                    for (int v = tid; v < N; v += nt) {
                        const int i = v / Z, k = v - i * Z;
                        const int ea = g.cedge[g.cp[i]], eb = g.cedge[g.cp[i] + 1];
                        const size_t ma = (size_t)ea * Z + wrapz(k - g.sh[ea] + Z, Z);
                        const size_t mb = (size_t)eb * Z + wrapz(k - g.sh[eb] + Z, Z);
                        uint16_t ip1 = yq[v];
                        uint16_t ip0 = (uint16_t)((ONE_SOFT << 4) - ip1);
                        uint16_t d1 = (uint16_t)(msg[mb] << 4);
                        uint16_t d0 = (uint16_t)(msg[ma] << 4);
                        uint16_t t1 = (uint16_t)((ONE_SOFT << 4) - d1);
                        uint16_t t0 = (uint16_t)((ONE_SOFT << 4) - d0);
                        uint16_t q10 = (uint16_t)DIVR((uint32_t)ip1 * d1, 16);
                        uint16_t q11 = (uint16_t)DIVR((uint32_t)ip1 * d0, 16);
                        uint16_t q00 = (uint16_t)DIVR((uint32_t)ip0 * t1, 16);
                        uint16_t q01 = (uint16_t)DIVR((uint32_t)ip0 * t0, 16);
                        uint16_t p1 = (uint16_t)DIVR((uint32_t)q10 * d0, 16);
                        uint16_t p0 = (uint16_t)DIVR((uint32_t)q00 * t0, 16);
                        p0 = (uint16_t)(p1 + p0);
                        p0 = p0 < 1 ? 1 : p0;
                        uint16_t sv = (uint16_t)(((uint32_t)p1 << 16) / p0);
                        so[v] = sv < (1 << 4) ? (1 << 4) : sv;
                        q00 = (uint16_t)(q00 + q10);
                        q01 = (uint16_t)(q01 + q11);
                        q00 = q00 < 1 ? 1 : q00;
                        q01 = q01 < 1 ? 1 : q01;
                        uint16_t n0 = (uint16_t)(((uint32_t)q10 << 12) / q00);
                        uint16_t n1 = (uint16_t)(((uint32_t)q11 << 12) / q01);
                        msg[ma] = n0 < 1 ? 1 : n0;
                        msg[mb] = n1 < 1 ? 1 : n1;
                    }
                    __syncthreads();
                } else {
                    for (int v = tid; v < N; v += nt) {                          // :3984-4052
                        const int i = v / Z, k = v - i * Z;
                        uint32_t P1 = (uint32_t)yq[v] << 16;
                        uint32_t P0 = (uint32_t)((ONE_SOFT << 4) - yq[v]) << 16;
                        for (int q = g.cp[i]; q < g.cp[i + 1]; q++) {
                            const int e = g.cedge[q];
                            uint16_t d = msg[(size_t)e * Z + wrapz(k - g.sh[e] + Z, Z)];
                            uint16_t d1 = (uint16_t)(d << 4);
                            uint16_t d0 = (uint16_t)((MAX_SOFT - d) << 4);       // MAX_SOFT, not ONE_SOFT (:4010)
                            unsigned long long pp1 = (unsigned long long)P1 * d1;
                            unsigned long long pp0 = (unsigned long long)P0 * d0;
                            P1 = (uint32_t)(pp1 >> 16);
                            P0 = (uint32_t)(pp0 >> 16);
                        }
                        uint32_t x = P1 >> 1;
                        uint32_t yv = (P0 >> 1) + x;
                        int flg = yv > (ONE_SOFT << 4);
                        if (flg) yv = yv >> 12; else x = x << 12;
                        yv = yv < 1 ? 1 : yv;
                        int s = (int)(x / yv);
                        s = MAX_SOFT < s ? MAX_SOFT : s;
                        uint16_t sv = (uint16_t)(s < 1 ? 1 : s);
                        so[v] = (uint16_t)(sv << 4);
                    }
                    __syncthreads();
                    for (int x = tid; x < g.E * Z; x += nt) {                    // :4055-4102
                        const int e = x / Z, n = x - e * Z;
                        int sv = so[g.col[e] * Z + wrapz(n + g.sh[e], Z)] << (12 - 4);
                        int sos = msg[x] < 1 ? 1 : msg[x];
                        int p1 = sv / sos;
                        int t = (ONE_SOFT - sos) < 1 ? 1 : (ONE_SOFT - sos);
                        int p0 = (ONE_SOFT * ONE_SOFT - sv) / t;
                        int yy = DIVR(p1 + p0, 6);
                        int y1 = yy < 1 ? 1 : yy;
                        int d = (p1 << 6) / y1;
                        d = d < 1 ? 1 : d;
                        msg[x] = (uint16_t)(MAX_SOFT < d ? MAX_SOFT : d);
                    }
                    __syncthreads();
                }
                synd = syndrome_pred(g, [&](int i) { return (int)(so[i] >> 15); });
                steps++;
                if (!synd) { if (!locked) { ret = steps; locked = 1; } if (!noexit) break; }
            }
        }
        if (!locked) ret = -steps;
        for (int i = tid; i < N; i += nt) store_post(io, N, f, i, so[i]);
        emit_frame(g, io, f, ret, [&](int i) { return (int)(so[i] >> 15); });
    }
};

// ------------------------------------------------------------------------------------------------
template <class Dec>
__global__ void __launch_bounds__(512) generic_sumprod_kernel(QcDev g, DecParams dp, FrameIO io, char* ws, size_t ws_stride)
{
    // the decoder state lives in shared memory when it fits (ws == nullptr), else in this CTA's slice of an
    // L2-resident global workspace
    extern __shared__ __align__(16) char dyn_ws[];
    char* w = ws ? ws + (size_t)blockIdx.x * ws_stride : dyn_ws;
    for (;;) {
        int f = next_frame(io);
        if (f >= io.nf) break;
        Dec::frame(g, dp, io, f, w);
    }
}

template <class Kern>
static void launch_one(Kern kern, const QcDev& g, const DecParams& dp, const FrameIO& io, char* ws, size_t ws_stride, size_t smem_ws,
                       int grid, int nt, cudaStream_t s)
{
    if (smem_ws) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ws);
    kern<<<grid, nt, smem_ws, s>>>(g, dp, io, smem_ws ? nullptr : ws, ws_stride);
}

size_t sumprod_workspace_bytes(int decoder_id, const QcHost& g, int nt)
{
    switch (decoder_id) {
    case LDPCB200_TASP_DEC: return TaspGeneric::ws_bytes(g, nt);
    case LDPCB200_ASP_DEC:  return AspGeneric::ws_bytes(g, nt);
    case LDPCB200_BP_DEC:   return BpGeneric::ws_bytes(g, nt);
    case LDPCB200_SP_DEC:   return SpGeneric::ws_bytes(g, nt);
    case LDPCB200_LCHE_DEC: return LcheGeneric::ws_bytes(g, nt);
    case LDPCB200_IASP_DEC: return IaspGeneric::ws_bytes(g, nt);
    }
    return 0;
}

cudaError_t launch_sumprod_generic(int decoder_id, const QcDev& g, const DecParams& dp, const FrameIO& io,
                                   char* ws, size_t ws_stride, size_t smem_ws, int grid, int nt, cudaStream_t s)
{
    switch (decoder_id) {
    case LDPCB200_TASP_DEC: launch_one(generic_sumprod_kernel<TaspGeneric>, g, dp, io, ws, ws_stride, smem_ws, grid, nt, s); break;
    case LDPCB200_ASP_DEC:  launch_one(generic_sumprod_kernel<AspGeneric>, g, dp, io, ws, ws_stride, smem_ws, grid, nt, s); break;
    case LDPCB200_BP_DEC:   launch_one(generic_sumprod_kernel<BpGeneric>, g, dp, io, ws, ws_stride, smem_ws, grid, nt, s); break;
    case LDPCB200_SP_DEC:   launch_one(generic_sumprod_kernel<SpGeneric>, g, dp, io, ws, ws_stride, smem_ws, grid, nt, s); break;
    case LDPCB200_LCHE_DEC: launch_one(generic_sumprod_kernel<LcheGeneric>, g, dp, io, ws, ws_stride, smem_ws, grid, nt, s); break;
    case LDPCB200_IASP_DEC: launch_one(generic_sumprod_kernel<IaspGeneric>, g, dp, io, ws, ws_stride, smem_ws, grid, nt, s); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

} // namespace ldpcb200
