// Throughput kernels of the sum-product family with the messages in TENSOR MEMORY -- table-driven (any code, no
// compilation step: these are also the kernels the code-search caller gets), one frame per CTA at a time, persistent
// grid, lane n = check row n of every block row:
//
//   tasp_fast_kernel<.., 0>      TASP_DEC  tdmp_sum_prod_gf2_decod_qc_lm   decoders.cpp:2584-2744   layered, probability domain, double
//   tasp_fast_kernel<.., 1>      LCHE_DEC  lche_decod                      decoders.cpp:2893-3010   layered, LLR domain + look-up tables, double
//   tasp_fast_kernel<.., 2>      LMS_DEC   lmin_sum_decod_qc_lm            decoders.cpp:5064-5425   layered offset min-sum in DOUBLE (bit-exact posteriors)
//   asp_fast_kernel              ASP_DEC   sum_prod_gf2_decod_qc_lm        decoders.cpp:2324-2581   flooding, probability domain, double
//   iasp_fast_kernel             IASP_DEC  isum_prod_gf2_decod_qc_lm       decoders.cpp:3822-4121   flooding, 12-bit fixed point
//   ms64_fast_kernel             MS_DEC    min_sum_decod_qc_lm             decoders.cpp:4554-4767   flooding normalised min-sum in DOUBLE
//   bpsp_fast_kernel<.., false>  BP_DEC    bp_decod_qc_lm                  decoders.cpp:1708-1920   flooding, LLR domain (tanh rule), double
//   bpsp_fast_kernel<.., true>   SP_DEC    sum_prod_decod_qc_lm            decoders.cpp:1923-2185   flooding, likelihood-ratio domain, double
//
// All in the reference's arithmetic and operation order (the expressions of the parity kernels in dec_sumprod.cu, which
// they equal bit for bit: tests/test_gpu_tmem.py).  What makes them faster than the parity kernels (which keep
// everything in an L2-resident workspace and their per-row arrays in local memory because the row weight is a run-time
// value):
//   * messages in TMEM: two 32-bit columns per edge and lane for the double decoders, one for IASP_DEC, fetched / put
//     back with one tcgen05.ld / tcgen05.st group per block row or per four edges (run-time column address);
//   * posteriors / per-bit products in shared memory, edge tables (bit offset, shift) in shared memory;
//   * block rows processed by functions templated on the row weight (switch dispatch), so rho[], the forward /
//     backward products of map_bin and the new messages live in registers; passes without row structure are flat
//     loops over the edges (small code: the CTAs of an SM sit at different places of the kernel);
//   * double divisions through div_normal (channel.cuh): correctly rounded, no slow-path branch.
// Codes whose messages do not fit the 512 TMEM columns or whose row weight exceeds TASP_MAXDEG stay on the parity kernels.
#include "dec_common.cuh"
#include "lms_spec.cuh"
#include "lms_tmem.cuh"
#include "fastmath64.cuh"

namespace ldpcb200 {

#define TASP_MAXDEG 20

struct TaspTab {
    int b, c, Z, N, R, E, nwords, tcols;
};

__device__ __forceinline__ double tf_mind(double a, double b) { return a < b ? a : b; }
__device__ __forceinline__ double tf_maxd(double a, double b) { return a < b ? b : a; }

// LLR -> P(bit = 1), decoders.cpp:2611-2618
__device__ __forceinline__ double tf_llr_to_p1(double llr)
{
    double x = llr * 0.5;
    double yv = tf_maxd(tf_mind(x, 20.0), -20.0);
    double e0 = exp(yv);
    double e1 = exp(-yv);
    return e1 / (e0 + e1);
}

// map_bin, decoders.cpp:2191-2228, row weight RW >= 2 known at compile time: everything in registers
template <int RW>
__device__ __forceinline__ void tf_map_bin(double (&s)[RW])
{
    double SF[RW], SB[RW], P[RW];
#pragma unroll
    for (int i = 0; i < RW; i++) P[i] = 1 - 2 * s[i];
    SF[0] = P[0];
#pragma unroll
    for (int i = 1; i < RW - 1; i++) SF[i] = P[i] * SF[i - 1];
    SB[RW - 1] = P[RW - 1];
#pragma unroll
    for (int i = RW - 2; i > 0; i--) SB[i] = P[i] * SB[i + 1];
    s[0] = (1 - SB[1]) / 2;
#pragma unroll
    for (int i = 1; i < RW - 1; i++) {
        double Zv = SF[i - 1] * SB[i + 1];
        s[i] = (1 - Zv) / 2;
    }
    s[RW - 1] = (1 - SF[RW - 2]) / 2;
}

// one block row of weight RW for check row n (clamped to Z-1 for the idle lanes of the last warp, which run the
// warp-collective TMEM instructions but store nothing)
template <int RW>
__device__ __forceinline__ void tf_row(double* gam, const unsigned* etab, int e0, int n, int Z, bool active, unsigned trow)
{
    const double T = 0.0001, TT = 0;                                             // decoders.cpp:2597-2598
    unsigned lw[2 * RW];
    tmem_ld_n<2 * RW>(trow + 2u * (unsigned)e0, lw);
    int idx[RW];
    double x[RW];
#pragma unroll
    for (int q = 0; q < RW; q++) {
        const unsigned pk = etab[e0 + q];                                        // bit offset of the column | shift << 16
        int k = n + (int)(pk >> 16);
        if (k >= Z) k -= Z;
        idx[q] = (int)(pk & 0xffffu) + k;
        x[q] = gam[idx[q]];
    }
    tmem_wait_ld<2 * RW>(lw);
    double rho[RW], a[RW];
#pragma unroll
    for (int q = 0; q < RW; q++) {
        const double av = __hiloint2double((int)lw[2 * q + 1], (int)lw[2 * q]);
        double r = div_normal(x[q] * (1.0 - av), av + x[q] - 2.0 * av * x[q]);       // :2686
        if (r < TT) r = TT;                                                      // :2692-2697
        if (r > 1 - TT) r = 1 - TT;
        rho[q] = r; a[q] = r;
    }
    tf_map_bin<RW>(a);                                                           // :2699
#pragma unroll
    for (int q = 0; q < RW; q++) {
        double av = a[q];
        if (av < T) av = T;                                                      // :2701-2705
        if (av > 1.0 - T) av = 1.0 - T;
        const double g = div_normal(rho[q] * av, 1.0 - rho[q] - av + 2 * rho[q] * av);   // :2716
        if (active) gam[idx[q]] = g;
        lw[2 * q] = (unsigned)__double2loint(av);
        lw[2 * q + 1] = (unsigned)__double2hiint(av);
    }
    tmem_st_n<2 * RW>(trow + 2u * (unsigned)e0, lw);
}

// ---- LCHE_DEC (lche_decod, decoders.cpp:2893-3010): the other layered sum-product decoder, LLR domain with three
// 32-entry look-up tables for log tanh (logexp_int :2777-2836).  Same kernel skeleton as TASP_DEC: messages st[] in
// tensor memory, posteriors so[] in shared memory, block rows templated on their weight; the tables are copied to
// shared memory (a per-lane index into __constant__ memory would serialise).
__constant__ double LF_A[32] = {
    1.41e+00, 7.72e-01, 4.54e-01, 2.72e-01, 1.65e-01, 9.97e-02, 6.04e-02, 3.66e-02,
    2.22e-02, 1.35e-02, 8.17e-03, 4.96e-03, 3.01e-03, 1.82e-03, 1.11e-03, 6.71e-04,
    4.07e-04, 2.47e-04, 1.50e-04, 9.08e-05, 5.51e-05, 3.34e-05, 2.03e-05, 1.23e-05,
    7.45e-06, 4.52e-06, 2.74e-06, 1.66e-06, 1.01e-06, 6.12e-07, 3.71e-07, 2.25e-07 };
__constant__ double LF_B[32] = {
    3.47, 2.77, 2.37, 2.08, 1.86, 1.69, 1.54, 1.41, 1.29, 1.19, 1.11, 1.03, 0.95, 0.89, 0.83, 0.77,
    0.72, 0.67, 0.63, 0.59, 0.55, 0.52, 0.48, 0.45, 0.43, 0.40, 0.37, 0.35, 0.33, 0.31, 0.29, 0.27 };
__constant__ double LF_C[32] = {
    6.93, 6.24, 5.83, 5.55, 5.32, 5.14, 4.99, 4.85, 4.73, 4.63, 4.53, 4.45, 4.37, 4.29, 4.22, 4.16,
    4.10, 4.04, 3.99, 3.94, 3.89, 3.84, 3.80, 3.75, 3.71, 3.67, 3.64, 3.60, 3.56, 3.53, 3.50, 3.47 };

// tab = A | B | C (96 doubles in shared memory)
__device__ __forceinline__ double lf_logexp_int(const double* tab, double x)     // :2777-2836
{
    if (x <= 0) x = 1.0 / 4096.0;
    if (x > 16.0) x = 16.0;
    if (x >= 2.0) return -tab[(int)(2 * x + 0.5) - 1];
    else if (x > 1.0 / 16.0) return -tab[32 + (int)(16 * x + 0.5) - 1];
    else if (x > 1.0 / 512.0) return -tab[64 + (int)(512 * x + 0.5) - 1];
    else {
        double s = 0;
        while (x < 1.0 / 512.0) { x *= 32; s -= 3.46; }
        return s - tab[64 + (int)(512 * x + 0.5) - 1];
    }
}

template <int RW>
__device__ __forceinline__ void lf_row(double* so, const double* tab, const unsigned* etab, int e0, int n, int Z, bool active, unsigned trow)
{
    unsigned lw[2 * RW];
    tmem_ld_n<2 * RW>(trow + 2u * (unsigned)e0, lw);
    int idx[RW];
    double x[RW];
#pragma unroll
    for (int q = 0; q < RW; q++) {
        const unsigned pk = etab[e0 + q];
        int k = n + (int)(pk >> 16);
        if (k >= Z) k -= Z;
        idx[q] = (int)(pk & 0xffffu) + k;
        x[q] = so[idx[q]];
    }
    tmem_wait_ld<2 * RW>(lw);
    double u[RW], yv[RW], alog[RW];
    int sy = 0;
    double sum = 0;
#pragma unroll
    for (int q = 0; q < RW; q++) u[q] = yv[q] = x[q] - __hiloint2double((int)lw[2 * q + 1], (int)lw[2 * q]);   // :2962
    // map_bin_llr, :2837-2890
#pragma unroll
    for (int q = 0; q < RW; q++) sy ^= u[q] < 0;
#pragma unroll
    for (int q = 0; q < RW; q++) alog[q] = lf_logexp_int(tab, u[q] < 0.0 ? -u[q] : u[q]);
#pragma unroll
    for (int q = 0; q < RW; q++) sum += alog[q];
#pragma unroll
    for (int q = 0; q < RW; q++) {
        const int hardb = (u[q] < 0) ^ sy;
        const double av = lf_logexp_int(tab, alog[q] - sum);
        u[q] = (2 * hardb - 1) * av;
    }
#pragma unroll
    for (int q = 0; q < RW; q++) {
        if (active) so[idx[q]] = u[q] + yv[q];                                   // :2979
        lw[2 * q] = (unsigned)__double2loint(u[q]);
        lw[2 * q + 1] = (unsigned)__double2hiint(u[q]);
    }
    tmem_st_n<2 * RW>(trow + 2u * (unsigned)e0, lw);
}

// ---- LMS_DEC in double (lmin_sum_decod_qc_lm, decoders.cpp:5064-5425): the reference's own arithmetic, for callers that
// want its posteriors bit for bit (LDPCB200_PRECISION=64; the fp32 kernels of lms_tmem.cuh agree on decisions and
// iteration counts but round their posteriors to fp32).  Same skeleton: c2v messages explicit in TMEM (two columns per
// edge and lane -- the reference's compressed prev[] {min1, min2, pos, sign} + signs[] reconstruct exactly these
// values, :5152-5158), posteriors in shared memory.  The offset max(|v2c| - 0.4, 0) and the 32767 ceiling are monotone,
// so they are applied to the two smallest |v2c| instead of to every edge; "position of the first minimum" is replaced by
// "magnitude equals the minimum" (ties select equal values).
template <int RW>
__device__ __forceinline__ void lm_row(double* so, const unsigned* etab, int e0, int n, int Z, bool active, unsigned trow)
{
    unsigned lw[2 * RW];
    tmem_ld_n<2 * RW>(trow + 2u * (unsigned)e0, lw);
    int idx[RW];
    double x[RW];
#pragma unroll
    for (int q = 0; q < RW; q++) {
        const unsigned pk = etab[e0 + q];
        int k = n + (int)(pk >> 16);
        if (k >= Z) k -= Z;
        idx[q] = (int)(pk & 0xffffu) + k;
        x[q] = so[idx[q]];
    }
    tmem_wait_ld<2 * RW>(lw);
    double v[RW], u[RW];
    int csign = 0;
    double c1 = __longlong_as_double(0x7ff0000000000000ll), c2 = c1;
#pragma unroll
    for (int q = 0; q < RW; q++) {
        v[q] = x[q] - __hiloint2double((int)lw[2 * q + 1], (int)lw[2 * q]);    // :5158
        csign ^= v[q] < 0;
        u[q] = v[q] < 0.0 ? -v[q] : v[q];
        if (u[q] < c1) { c2 = c1; c1 = u[q]; }                                   // process_check_node :5012-5027 on the raw magnitudes
        else if (u[q] < c2) c2 = u[q];
    }
    double n1 = c1 - 0.4, n2 = c2 - 0.4;                                          // :5166-5168, ceiling MAX_VAL :5131-5137
    n1 = n1 < 0 ? 0.0 : n1; n2 = n2 < 0 ? 0.0 : n2;
    n1 = n1 < 32767.0 ? n1 : 32767.0; n2 = n2 < 32767.0 ? n2 : 32767.0;
#pragma unroll
    for (int q = 0; q < RW; q++) {
        const double cabs = u[q] == c1 ? n2 : n1;                                // :5193
        const double cval = ((v[q] < 0) ^ csign) ? -cabs : cabs;                 // :5194-5198
        if (active) so[idx[q]] = v[q] + cval;                                    // :5199-5204
        lw[2 * q] = (unsigned)__double2loint(cval);
        lw[2 * q + 1] = (unsigned)__double2hiint(cval);
    }
    tmem_st_n<2 * RW>(trow + 2u * (unsigned)e0, lw);
}

template <int RW, int FL>
__device__ __noinline__ void tf_row_call(double* gam, const double* tab, const unsigned* etab, int e0, int n, int Z, bool active, unsigned trow)
{
    if constexpr (FL == 2) lm_row<RW>(gam, etab, e0, n, Z, active, trow);
    else if constexpr (FL == 1) lf_row<RW>(gam, tab, etab, e0, n, Z, active, trow);
    else tf_row<RW>(gam, etab, e0, n, Z, active, trow);
}

template <int FL>
__device__ __forceinline__ void tf_dispatch(int cnt, double* gam, const double* tab, const unsigned* etab, int e0, int n, int Z, bool active, unsigned trow)
{
    switch (cnt) {
#define TF_CASE(k) case k: tf_row_call<k, FL>(gam, tab, etab, e0, n, Z, active, trow); break;
    TF_CASE(2) TF_CASE(3) TF_CASE(4) TF_CASE(5) TF_CASE(6) TF_CASE(7) TF_CASE(8) TF_CASE(9) TF_CASE(10) TF_CASE(11)
    TF_CASE(12) TF_CASE(13) TF_CASE(14) TF_CASE(15) TF_CASE(16) TF_CASE(17) TF_CASE(18) TF_CASE(19) TF_CASE(20)
#undef TF_CASE
    default: break;
    }
}

// syndrome of the decisions gamma > 0.5 (check_syndrome_thr, decoders.cpp:2274) or, for LCHE_DEC, so < 0: every lane XORs its rows
template <bool LCHE = false>
__device__ __forceinline__ int tf_syndrome(const double* gam, const unsigned* etab, const int* rpw, int b, int Z, int n, bool active)
{
    int bad = 0;
    if (active) {
        for (int j = 0; j < b; j++) {
            int s = 0;
            for (int e = rpw[j]; e < rpw[j + 1]; e++) {
                const unsigned pk = etab[e];
                int k = n + (int)((pk >> 16) & 0x7fffu);
                if (k >= Z) k -= Z;
                const double v = gam[(int)(pk & 0xffffu) + k];
                s ^= LCHE ? (int)(v < 0) : (int)(v > 0.5);
            }
            bad |= s;
        }
    }
    return __syncthreads_or(bad);
}

// FL: 0 TASP_DEC, 1 LCHE_DEC, 2 LMS_DEC in double
template <int FL>
__device__ __forceinline__ void tasp_fast_body(const TaspTab& T, const QcDev& g, const FrameIO& io)
{
    extern __shared__ __align__(16) double tf_smem[];
    const int Z = T.Z, N = T.N, E = T.E, b = T.b, nt = blockDim.x, tid = threadIdx.x;
    double* gam = tf_smem;                                      // TASP: posteriors P(bit = 1); LCHE: posterior LLRs so[]
    double* tab = gam + N;                                      // LCHE: the three look-up tables
    unsigned* etab = (unsigned*)(tab + 96);
    int* rpw = (int*)(etab + E);
    unsigned* s_t = (unsigned*)(rpw + b + 1);
    const bool active = tid < Z;
    const int n = active ? tid : Z - 1;
    const bool noexit = io.flags & LDPCB200_NO_EARLY_EXIT;

    for (int e = tid; e < E; e += nt) etab[e] = (unsigned)(g.col[e] * Z) | ((unsigned)g.sh[e] << 16);
    for (int j = tid; j <= b; j += nt) rpw[j] = g.rp[j];
    for (int i = tid; i < 96; i += nt) tab[i] = i < 32 ? LF_A[i] : i < 64 ? LF_B[i - 32] : LF_C[i - 64];
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"((unsigned)__cvta_generic_to_shared(s_t)), "r"((unsigned)T.tcols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tbase = *(volatile unsigned*)s_t;
    // this thread's TMEM lane (bits 31:16) and first column: 2 columns per edge, 2 * E columns per group of 4 warps
    const unsigned trow = __shfl_sync(0xffffffffu, tbase + ((unsigned)(((tid >> 5) & 3) * 32) << 16) + (unsigned)((tid >> 7) * 2 * E), 0);

    for (;;) {
        const int f = next_frame(io);
        if (f >= io.nf) break;
        for (int i = tid; i < N; i += nt) gam[i] = FL ? load_llr(io, N, f, i) : tf_llr_to_p1(load_llr(io, N, f, i));   // :2611-2646 / :2916 / :5088
        {   // TASP: lambda = 0.5 for every edge (:2620-2641), 0x3FE0000000000000; LCHE: st = 0 (:2913-2915)
            unsigned half[2] = { 0u, FL ? 0u : 0x3FE00000u };
            for (int e = 0; e < E; e++) TmemRow<2>::st(trow + 2u * (unsigned)e, half);
            tmem_wait_st();
        }
        __syncthreads();
        int synd = tf_syndrome<FL != 0>(gam, etab, rpw, b, Z, n, active);                        // :2653 / :2927-2931 / :5111-5115
        int ret = 0, locked = 0, steps = 0;
        if (!synd) { locked = 1; ret = FL == 2 ? 1 : 0; }                                        // :2654-2660; LMS_DEC: 0 + 1 (:5424)
        if (synd || noexit) {
            while (steps < io.maxiter) {
                tmem_wait_st();
                for (int j = 0; j < b; j++) {
                    const int e0 = rpw[j];
                    tf_dispatch<FL>(rpw[j + 1] - e0, gam, tab, etab, e0, n, Z, active, trow);
                    __syncthreads();
                }
                // the reference re-checks after every layer (:2723); only the last verdict is used (:2733)
                synd = tf_syndrome<FL != 0>(gam, etab, rpw, b, Z, n, active);
                steps++;
                if (!synd) { if (!locked) { ret = steps; locked = 1; } if (!noexit) break; }
            }
        }
        if (!locked) ret = synd ? -steps : steps;                                                // :2740-2743
        for (int i = tid; i < N; i += nt) store_post(io, N, f, i, gam[i]);
        emit_frame(g, io, f, ret, [&](int i) { return FL ? (int)(gam[i] < 0.0) : (int)(gam[i] > 0.5); });   // :2738 / :3003 / :5421
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase), "r"((unsigned)T.tcols) : "memory");
    }
}

template <int MAXT, int FL>
__global__ void __launch_bounds__(MAXT, 1) tasp_fast_kernel(const TaspTab T, const QcDev g, const FrameIO io)
{
    tasp_fast_body<FL>(T, g, io);
}

// Many candidate codes of one shape in ONE launch (the code-search caller, main_good_code_search.cpp:
// 267-411, scores every candidate with the same short TASP_DEC simulation): blockIdx.y is the code.  Each code has its own
// edge tables, frame counter, error counters and (optional) per-frame records -- gs[k] / ios[k], device arrays -- and the
// codes share the geometry T (b, c, Z, E: same shape, same number of circulants) and the kernel; the table-driven kernel
// needs no compilation step, which is what makes a fresh matrix per grid row possible.
template <int MAXT>
__global__ void __launch_bounds__(MAXT, 1) tasp_multi_kernel(const TaspTab T, const QcDev* __restrict__ gs, const FrameIO* __restrict__ ios)
{
    const QcDev g = gs[blockIdx.y];
    const FrameIO io = ios[blockIdx.y];
    tasp_fast_body<0>(T, g, io);
}

// ------------------------------------------------------------------------------------------------------------------
// ASP_DEC (flooding sum-product in the probability domain, sum_prod_gf2_decod_qc_lm, decoders.cpp:2324-2581; the general
// path -- codes whose columns all have weight 2 take the reference's shortcut :2432-2482 and stay on the parity kernel).
// Same recipe: messages (two TMEM words per edge and lane), block rows templated on their weight, div_normal.  An
// iteration is three sweeps:
//   R  block row by block row (barrier in between): map_bin on the row's messages (:2406-2428), then the row multiplies
//      its new messages into the per-bit products P1 *= d, P0 *= 1 - d (:2489-2520).  The reference multiplies a bit's
//      messages in ascending block-row order starting from the prior; lanes of one block row touch disjoint bits and a
//      bit gets at most one message per block row, so the order is the reference's (the first block row of a column
//      starts from the prior instead of reading the product);
//   V  per bit: so = P1 / (P0 + P1) (:2522);
//   E  per check row, no barriers: msg = clamp(p1 / (p1 + p0)), p1 = so / d, p0 = (1 - so) / (1 - d) (:2525-2558).
template <int RW>
__device__ __noinline__ void af_rowR(double* A, double* Bv, const double* prior, const unsigned* etab, int e0, int n, int Z, bool active, unsigned trow)
{
    unsigned lw[2 * RW];
    tmem_ld_n<2 * RW>(trow + 2u * (unsigned)e0, lw);
    int idx[RW];
    double P1[RW], P0[RW];
#pragma unroll
    for (int q = 0; q < RW; q++) {
        const unsigned pk = etab[e0 + q];                                        // bit offset | shift << 16 | first-of-column << 31
        int k = n + (int)((pk >> 16) & 0x7fffu);
        if (k >= Z) k -= Z;
        idx[q] = (int)(pk & 0xffffu) + k;
        if (pk >> 31) { P1[q] = prior[idx[q]]; P0[q] = 1 - P1[q]; }              // :2489-2490
        else { P1[q] = A[idx[q]]; P0[q] = Bv[idx[q]]; }
    }
    tmem_wait_ld<2 * RW>(lw);
    double a[RW];
#pragma unroll
    for (int q = 0; q < RW; q++) a[q] = __hiloint2double((int)lw[2 * q + 1], (int)lw[2 * q]);
    tf_map_bin<RW>(a);                                                           // :2406-2428
#pragma unroll
    for (int q = 0; q < RW; q++) {
        if (active) { A[idx[q]] = P1[q] * a[q]; Bv[idx[q]] = P0[q] * (1 - a[q]); }   // :2496-2497
        lw[2 * q] = (unsigned)__double2loint(a[q]);
        lw[2 * q + 1] = (unsigned)__double2hiint(a[q]);
    }
    tmem_st_n<2 * RW>(trow + 2u * (unsigned)e0, lw);
}

// sweep E needs no row structure: a flat loop over the edges, four at a time (one tcgen05.ld / st group per four edges),
// keeps the code small -- the row-templated form was 19 unrolled copies of the three divisions per edge, and four
// CTAs at different places of it stalled on instruction fetch
template <int NE>
__device__ __forceinline__ void af_edgesE(const double* A, const unsigned* etab, int e, int n, int Z, unsigned trow)
{
    unsigned lw[2 * NE];
    tmem_ld_n<2 * NE>(trow + 2u * (unsigned)e, lw);
    double s1[NE];
#pragma unroll
    for (int q = 0; q < NE; q++) {
        const unsigned pk = etab[e + q];
        int k = n + (int)((pk >> 16) & 0x7fffu);
        if (k >= Z) k -= Z;
        s1[q] = A[(int)(pk & 0xffffu) + k];
    }
    tmem_wait_ld<2 * NE>(lw);
#pragma unroll
    for (int q = 0; q < NE; q++) {
        const double sos = __hiloint2double((int)lw[2 * q + 1], (int)lw[2 * q]);
        const double p1 = div_normal(s1[q], sos);                                // :2537-2550
        const double p0 = div_normal(1 - s1[q], 1 - sos);
        double d = div_normal(p1, p1 + p0);
        d = tf_maxd(tf_mind(d, 1.0 - 0.000001), 0.000001);                       // SP_DEC_MIN/MAX_VAL, :96-97
        lw[2 * q] = (unsigned)__double2loint(d);
        lw[2 * q + 1] = (unsigned)__double2hiint(d);
    }
    tmem_st_n<2 * NE>(trow + 2u * (unsigned)e, lw);
}
__device__ __noinline__ void af_sweepE(const double* A, const unsigned* etab, int E, int n, int Z, unsigned trow)
{
    int e = 0;
#pragma unroll 1
    for (; e + 4 <= E; e += 4) af_edgesE<4>(A, etab, e, n, Z, trow);
#pragma unroll 1
    for (; e < E; e++) af_edgesE<1>(A, etab, e, n, Z, trow);
}

#define AF_CASES(CALL) \
    CALL(2) CALL(3) CALL(4) CALL(5) CALL(6) CALL(7) CALL(8) CALL(9) CALL(10) CALL(11) CALL(12) CALL(13) CALL(14) CALL(15) CALL(16) \
    CALL(17) CALL(18) CALL(19) CALL(20)

template <int MAXT>
__global__ void __launch_bounds__(MAXT, 1) asp_fast_kernel(const TaspTab T, const QcDev g, const FrameIO io)
{
    extern __shared__ __align__(16) double tf_smem[];
    const int Z = T.Z, N = T.N, E = T.E, b = T.b, nt = blockDim.x, tid = threadIdx.x;
    double* A = tf_smem;                 // per-bit product P1, then the posterior so
    double* Bv = A + N;                  // per-bit product P0
    double* prior = Bv + N;              // p = P(bit = 1 | channel)
    unsigned* etab = (unsigned*)(prior + N);
    int* rpw = (int*)(etab + E);
    unsigned* s_t = (unsigned*)(rpw + b + 1);
    const bool active = tid < Z;
    const int n = active ? tid : Z - 1;
    const bool noexit = io.flags & LDPCB200_NO_EARLY_EXIT;

    for (int e = tid; e < E; e += nt) {
        const int c = g.col[e];
        etab[e] = (unsigned)(c * Z) | ((unsigned)g.sh[e] << 16) | (g.cedge[g.cp[c]] == e ? 0x80000000u : 0u);
    }
    for (int j = tid; j <= b; j += nt) rpw[j] = g.rp[j];
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"((unsigned)__cvta_generic_to_shared(s_t)), "r"((unsigned)T.tcols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tbase = *(volatile unsigned*)s_t;
    const unsigned trow = __shfl_sync(0xffffffffu, tbase + ((unsigned)(((tid >> 5) & 3) * 32) << 16) + (unsigned)((tid >> 7) * 2 * E), 0);

    for (;;) {
        const int f = next_frame(io);
        if (f >= io.nf) break;
        for (int i = tid; i < N; i += nt) { const double pv = tf_llr_to_p1(load_llr(io, N, f, i)); prior[i] = pv; A[i] = pv; }   // :2351-2358
        __syncthreads();
        for (int e = 0; e < E; e++) {                                                            // msg = prior of the edge's bit, :2361-2378
            const unsigned pk = etab[e];
            int k = n + (int)((pk >> 16) & 0x7fffu);
            if (k >= Z) k -= Z;
            const double pv = prior[(int)(pk & 0xffffu) + k];
            unsigned w[2] = { (unsigned)__double2loint(pv), (unsigned)__double2hiint(pv) };
            TmemRow<2>::st(trow + 2u * (unsigned)e, w);
        }
        tmem_wait_st();
        int synd = tf_syndrome(A, etab, rpw, b, Z, n, active);                                   // :2392 (bit 31 of etab is masked there)
        int ret = 0, locked = 0, steps = 0;
        if (!synd) { locked = 1; ret = 0; }
        if (synd || noexit) {
            while (steps < io.maxiter) {
                for (int j = 0; j < b; j++) {                                                    // sweep R
                    const int e0 = rpw[j];
                    switch (rpw[j + 1] - e0) {
#define AF_R(k) case k: af_rowR<k>(A, Bv, prior, etab, e0, n, Z, active, trow); break;
                    AF_CASES(AF_R)
#undef AF_R
                    default: break;
                    }
                    __syncthreads();
                }
                for (int i = tid; i < N; i += nt) A[i] = div_normal(A[i], Bv[i] + A[i]);         // sweep V, :2522
                __syncthreads();
                tmem_wait_st();
                af_sweepE(A, etab, E, n, Z, trow);                                               // sweep E
                tmem_wait_st();
                synd = tf_syndrome(A, etab, rpw, b, Z, n, active);                               // :2566
                steps++;
                if (!synd) { if (!locked) { ret = steps; locked = 1; } if (!noexit) break; }
            }
        }
        if (!locked) ret = -steps;
        for (int i = tid; i < N; i += nt) store_post(io, N, f, i, A[i]);
        emit_frame(g, io, f, ret, [&](int i) { return (int)(A[i] > 0.5); });                     // make_output :2308
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase), "r"((unsigned)T.tcols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------------------------
// IASP_DEC (isum_prod_gf2_decod_qc_lm, decoders.cpp:3822-4121): the 12-bit fixed-point version of ASP_DEC, bit-exact.
// The sweeps of asp_fast_kernel in the reference's integer arithmetic: messages are 12-bit values, ONE TMEM word per
// edge and lane; the per-bit products P1 / P0 are 32-bit words in shared memory, multiplied in ascending block-row order
// from the check side (64-bit products shifted right by 16, :4003-4016); rounding shifts are DIVR (div_power2r :80).
#define IF_ONE 4096
#define IF_MAX 4095
#define IF_DIVR(x, n) (((x) + (1 << ((n) - 1))) >> (n))

template <int RW>
__device__ __forceinline__ void if_imap_bin(int (&s)[RW])                        // imap_bin :2235-2271, s[] are uint16 values
{
    short SF[RW], SB[RW], P[RW];
#pragma unroll
    for (int i = 0; i < RW; i++) P[i] = (short)(IF_ONE - 2 * s[i]);
    SF[0] = P[0];
#pragma unroll
    for (int i = 1; i < RW - 1; i++) SF[i] = (short)IF_DIVR((int)P[i] * SF[i - 1], 12);
    SB[RW - 1] = P[RW - 1];
#pragma unroll
    for (int i = RW - 2; i > 0; i--) SB[i] = (short)IF_DIVR((int)P[i] * SB[i + 1], 12);
    s[0] = (unsigned short)IF_DIVR(IF_ONE - SB[1], 1);
    s[0] = s[0] < 1 ? 1 : s[0];
#pragma unroll
    for (int i = 1; i < RW - 1; i++) {
        const int Zv = IF_DIVR((int)SF[i - 1] * SB[i + 1], 12);
        s[i] = (unsigned short)IF_DIVR(IF_ONE - Zv, 1);
        s[i] = s[i] < 1 ? 1 : s[i];
    }
    s[RW - 1] = (unsigned short)IF_DIVR(IF_ONE - SF[RW - 2], 1);
    s[RW - 1] = s[RW - 1] < 1 ? 1 : s[RW - 1];
}

template <int RW>
__device__ __noinline__ void if_rowR(unsigned* A, unsigned* Bv, const unsigned* yq, const unsigned* etab, int e0, int n, int Z, bool active, unsigned trow)
{
    unsigned lw[RW];
    tmem_ld_n<RW>(trow + (unsigned)e0, lw);
    int idx[RW];
    unsigned P1[RW], P0[RW];
#pragma unroll
    for (int q = 0; q < RW; q++) {
        const unsigned pk = etab[e0 + q];                                        // bit offset | shift << 16 | first-of-column << 31
        int k = n + (int)((pk >> 16) & 0x7fffu);
        if (k >= Z) k -= Z;
        idx[q] = (int)(pk & 0xffffu) + k;
        if (pk >> 31) { const unsigned y = yq[idx[q]]; P1[q] = y << 16; P0[q] = (unsigned)((IF_ONE << 4) - y) << 16; }   // :3987-3988
        else { P1[q] = A[idx[q]]; P0[q] = Bv[idx[q]]; }
    }
    tmem_wait_ld<RW>(lw);
    int a[RW];
#pragma unroll
    for (int q = 0; q < RW; q++) a[q] = (int)lw[q];
    if_imap_bin<RW>(a);                                                          // :3906-3911
#pragma unroll
    for (int q = 0; q < RW; q++) {
        const unsigned d = (unsigned)a[q] & 0xffffu;
        const unsigned d1 = (d << 4) & 0xffffu;                                  // (uint16)(d << 4), :4008
        const unsigned d0 = ((unsigned)(IF_MAX - (int)d) << 4) & 0xffffu;        // MAX_SOFT, not ONE_SOFT (:4010)
        if (active) {
            A[idx[q]] = (unsigned)(((unsigned long long)P1[q] * d1) >> 16);      // :4011-4016
            Bv[idx[q]] = (unsigned)(((unsigned long long)P0[q] * d0) >> 16);
        }
        lw[q] = d;
    }
    tmem_st_n<RW>(trow + (unsigned)e0, lw);
}

template <int NE>
__device__ __forceinline__ void if_edgesE(const unsigned* A, const unsigned* etab, int e, int n, int Z, unsigned trow)
{
    unsigned lw[NE];
    tmem_ld_n<NE>(trow + (unsigned)e, lw);
    int so[NE];
#pragma unroll
    for (int q = 0; q < NE; q++) {
        const unsigned pk = etab[e + q];
        int k = n + (int)((pk >> 16) & 0x7fffu);
        if (k >= Z) k -= Z;
        so[q] = (int)A[(int)(pk & 0xffffu) + k];
    }
    tmem_wait_ld<NE>(lw);
#pragma unroll
    for (int q = 0; q < NE; q++) {                                               // :4055-4102
        const int sv = so[q] << (12 - 4);
        const int m = (int)lw[q];
        const int sos = m < 1 ? 1 : m;
        const int p1 = sv / sos;
        const int t = (IF_ONE - sos) < 1 ? 1 : (IF_ONE - sos);
        const int p0 = (IF_ONE * IF_ONE - sv) / t;
        const int yy = IF_DIVR(p1 + p0, 6);
        const int y1 = yy < 1 ? 1 : yy;
        int d = (p1 << 6) / y1;
        d = d < 1 ? 1 : d;
        lw[q] = (unsigned)(IF_MAX < d ? IF_MAX : d);
    }
    tmem_st_n<NE>(trow + (unsigned)e, lw);
}
__device__ __noinline__ void if_sweepE(const unsigned* A, const unsigned* etab, int E, int n, int Z, unsigned trow)
{
    int e = 0;
#pragma unroll 1
    for (; e + 4 <= E; e += 4) if_edgesE<4>(A, etab, e, n, Z, trow);
#pragma unroll 1
    for (; e < E; e++) if_edgesE<1>(A, etab, e, n, Z, trow);
}

// syndrome of the decisions so >> 15 (icheck_syndrome :3772)
__device__ __forceinline__ int if_syndrome(const unsigned* so, const unsigned* etab, const int* rpw, int b, int Z, int n, bool active)
{
    int bad = 0;
    if (active) {
        for (int j = 0; j < b; j++) {
            unsigned s = 0;
            for (int e = rpw[j]; e < rpw[j + 1]; e++) {
                const unsigned pk = etab[e];
                int k = n + (int)((pk >> 16) & 0x7fffu);
                if (k >= Z) k -= Z;
                s ^= so[(int)(pk & 0xffffu) + k] >> 15;
            }
            bad |= (int)(s & 1u);
        }
    }
    return __syncthreads_or(bad);
}

template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) iasp_fast_kernel(const TaspTab T, const QcDev g, const FrameIO io)
{
    extern __shared__ __align__(16) double tf_smem[];
    const int Z = T.Z, N = T.N, E = T.E, b = T.b, nt = blockDim.x, tid = threadIdx.x;
    unsigned* A = (unsigned*)tf_smem;    // per-bit product P1, then the posterior so (16-bit value)
    unsigned* Bv = A + N;                // per-bit product P0
    unsigned* yq = Bv + N;               // quantised prior << 4 (:3867)
    unsigned* etab = yq + N;
    int* rpw = (int*)(etab + E);
    unsigned* s_t = (unsigned*)(rpw + b + 1);
    const bool active = tid < Z;
    const int n = active ? tid : Z - 1;
    const bool noexit = io.flags & LDPCB200_NO_EARLY_EXIT;

    for (int e = tid; e < E; e += nt) {
        const int c = g.col[e];
        etab[e] = (unsigned)(c * Z) | ((unsigned)g.sh[e] << 16) | (g.cedge[g.cp[c]] == e ? 0x80000000u : 0u);
    }
    for (int j = tid; j <= b; j += nt) rpw[j] = g.rp[j];
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"((unsigned)__cvta_generic_to_shared(s_t)), "r"((unsigned)T.tcols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tbase = *(volatile unsigned*)s_t;
    // one TMEM column per edge, E columns per group of 4 warps
    const unsigned trow = __shfl_sync(0xffffffffu, tbase + ((unsigned)(((tid >> 5) & 3) * 32) << 16) + (unsigned)((tid >> 7) * E), 0);

    for (;;) {
        const int f = next_frame(io);
        if (f >= io.nf) break;
        for (int i = tid; i < N; i += nt) {                                                      // :3849-3861
            const double v = tf_maxd(tf_mind(load_llr(io, N, f, i), 20.0), -20.0);
            const double pr = 1.0 / (1.0 + exp(v));
            int x = (int)(pr * IF_ONE + 0.5);
            x = IF_MAX < x ? IF_MAX : x;
            yq[i] = (unsigned)(x < 1 ? 1 : x);
        }
        __syncthreads();
        for (int e = 0; e < E; e++) {                                                            // msg = 12-bit prior of the edge's bit, :3869-3886
            const unsigned pk = etab[e];
            int k = n + (int)((pk >> 16) & 0x7fffu);
            if (k >= Z) k -= Z;
            unsigned w[1] = { yq[(int)(pk & 0xffffu) + k] };
            TmemRow<1>::st(trow + (unsigned)e, w);
        }
        tmem_wait_st();
        __syncthreads();
        for (int i = tid; i < N; i += nt) { const unsigned y = (yq[i] << 4) & 0xffffu; yq[i] = y; A[i] = y; }   // :3867, :3889
        __syncthreads();
        int synd = if_syndrome(A, etab, rpw, b, Z, n, active);
        int ret = 0, locked = 0, steps = 0;
        if (!synd) { locked = 1; ret = 0; }
        if (synd || noexit) {
            while (steps < io.maxiter) {
                for (int j = 0; j < b; j++) {                                                    // sweep R
                    const int e0 = rpw[j];
                    switch (rpw[j + 1] - e0) {
#define IF_R(k) case k: if_rowR<k>(A, Bv, yq, etab, e0, n, Z, active, trow); break;
                    AF_CASES(IF_R)
#undef IF_R
                    default: break;
                    }
                    __syncthreads();
                }
                for (int i = tid; i < N; i += nt) {                                              // sweep V, :4018-4030
                    unsigned x = A[i] >> 1;
                    unsigned yv = (Bv[i] >> 1) + x;
                    const int flg = yv > (unsigned)(IF_ONE << 4);
                    if (flg) yv = yv >> 12; else x = x << 12;
                    yv = yv < 1 ? 1 : yv;
                    int sv = (int)(x / yv);
                    sv = IF_MAX < sv ? IF_MAX : sv;
                    sv = sv < 1 ? 1 : sv;
                    A[i] = ((unsigned)sv << 4) & 0xffffu;
                }
                __syncthreads();
                tmem_wait_st();
                if_sweepE(A, etab, E, n, Z, trow);                                               // sweep E
                tmem_wait_st();
                synd = if_syndrome(A, etab, rpw, b, Z, n, active);
                steps++;
                if (!synd) { if (!locked) { ret = steps; locked = 1; } if (!noexit) break; }
            }
        }
        if (!locked) ret = -steps;
        for (int i = tid; i < N; i += nt) store_post(io, N, f, i, A[i]);
        emit_frame(g, io, f, ret, [&](int i) { return (int)(A[i] >> 15); });
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase), "r"((unsigned)T.tcols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------------------------
// MS_DEC in double (min_sum_decod_qc_lm, decoders.cpp:4554-4767): flooding normalised min-sum in the reference's own
// arithmetic (LDPCB200_PRECISION=64).  The sweeps of ms_tmem.cuh on this file's table-driven skeleton: messages (+-min,
// unscaled: what the reference's STATE 1 adds, :4649-4658) are two TMEM columns per edge and lane;
//   A  block row by block row (barrier in between): acc[bit] += message -- every bit sees its messages in ascending
//      block-row order starting from 0, as :4633-4658 does;
//   B  per bit: soft = y + acc * alpha (:4682);
//   C  per check row, no barriers: v2c = soft - message * alpha, syndrome of this pass's decisions, the two smallest
//      |v2c| (ceiling 32767 applied to them instead of to every edge, :4730), new messages (:4688-4755).
template <int RW>
__device__ __noinline__ void m6_rowA(double* A, const unsigned* etab, int e0, int n, int Z, bool active, unsigned trow)
{
    unsigned lw[2 * RW];
    tmem_ld_n<2 * RW>(trow + 2u * (unsigned)e0, lw);
    int idx[RW];
    double acc[RW];
#pragma unroll
    for (int q = 0; q < RW; q++) {
        const unsigned pk = etab[e0 + q];                                        // bit offset | shift << 16 | first-of-column << 31
        int k = n + (int)((pk >> 16) & 0x7fffu);
        if (k >= Z) k -= Z;
        idx[q] = (int)(pk & 0xffffu) + k;
        acc[q] = (pk >> 31) ? 0.0 : A[idx[q]];                                   // :4633
    }
    tmem_wait_ld<2 * RW>(lw);
#pragma unroll
    for (int q = 0; q < RW; q++)
        if (active) A[idx[q]] = acc[q] + __hiloint2double((int)lw[2 * q + 1], (int)lw[2 * q]);   // :4658
}

template <int RW>
__device__ __noinline__ int m6_rowC(const double* A, const unsigned* etab, int e0, int n, int Z, double alpha, unsigned trow)
{
    unsigned lw[2 * RW];
    tmem_ld_n<2 * RW>(trow + 2u * (unsigned)e0, lw);
    double rs[RW];
#pragma unroll
    for (int q = 0; q < RW; q++) {
        const unsigned pk = etab[e0 + q];
        int k = n + (int)((pk >> 16) & 0x7fffu);
        if (k >= Z) k -= Z;
        rs[q] = A[(int)(pk & 0xffffu) + k];
    }
    tmem_wait_ld<2 * RW>(lw);
    double tt[RW], u[RW];
    int synd = 0, csign = 0;
    double c1 = __longlong_as_double(0x7ff0000000000000ll), c2 = c1;
#pragma unroll
    for (int q = 0; q < RW; q++) {
        synd ^= rs[q] < 0;                                                       // :4711
        tt[q] = rs[q] - __hiloint2double((int)lw[2 * q + 1], (int)lw[2 * q]) * alpha;   // :4714-4722
        csign ^= tt[q] < 0;
        u[q] = tt[q] < 0.0 ? -tt[q] : tt[q];                                     // :4729
        if (u[q] < c1) { c2 = c1; c1 = u[q]; }                                   // :4732-4746 on the raw magnitudes
        else if (u[q] < c2) c2 = u[q];
    }
    const double n1 = c1 > 32767.0 ? 32767.0 : c1, n2 = c2 > 32767.0 ? 32767.0 : c2;     // :4730, init :4692-4696
#pragma unroll
    for (int q = 0; q < RW; q++) {
        const double cabs = u[q] == c1 ? n2 : n1;                                // :4649 (position of the minimum -> equality, ties select equal values)
        const double cval = ((tt[q] < 0) ^ csign) ? -cabs : cabs;
        lw[2 * q] = (unsigned)__double2loint(cval);
        lw[2 * q + 1] = (unsigned)__double2hiint(cval);
    }
    tmem_st_n<2 * RW>(trow + 2u * (unsigned)e0, lw);
    return synd;
}

template <int MAXT>
__global__ void __launch_bounds__(MAXT, 1) ms64_fast_kernel(const TaspTab T, const QcDev g, const FrameIO io, const double alpha)
{
    extern __shared__ __align__(16) double tf_smem[];
    const int Z = T.Z, N = T.N, E = T.E, b = T.b, nt = blockDim.x, tid = threadIdx.x;
    double* A = tf_smem;                 // accumulator, then the posterior soft
    double* y = A + N;                   // channel values
    unsigned* etab = (unsigned*)(y + N);
    int* rpw = (int*)(etab + E);
    unsigned* s_t = (unsigned*)(rpw + b + 1);
    const bool active = tid < Z;
    const int n = active ? tid : Z - 1;
    const bool noexit = io.flags & LDPCB200_NO_EARLY_EXIT;

    for (int e = tid; e < E; e += nt) {
        const int c = g.col[e];
        etab[e] = (unsigned)(c * Z) | ((unsigned)g.sh[e] << 16) | (g.cedge[g.cp[c]] == e ? 0x80000000u : 0u);
    }
    for (int j = tid; j <= b; j += nt) rpw[j] = g.rp[j];
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"((unsigned)__cvta_generic_to_shared(s_t)), "r"((unsigned)T.tcols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tbase = *(volatile unsigned*)s_t;
    const unsigned trow = __shfl_sync(0xffffffffu, tbase + ((unsigned)(((tid >> 5) & 3) * 32) << 16) + (unsigned)((tid >> 7) * 2 * E), 0);

    for (;;) {
        const int f = next_frame(io);
        if (f >= io.nf) break;
        for (int i = tid; i < N; i += nt) { const double v = load_llr(io, N, f, i); y[i] = v; A[i] = v; }      // :4579-4596
        {
            unsigned zero[2] = { 0u, 0u };
            for (int e = 0; e < E; e++) TmemRow<2>::st(trow + 2u * (unsigned)e, zero);
            tmem_wait_st();
        }
        __syncthreads();
        int parity = 1, ret = 0, locked = 0, iter;
        for (iter = 0; iter < io.maxiter; iter++) {
            for (int j = 0; j < b; j++) {                                                        // sweep A (STATE 1)
                const int e0 = rpw[j];
                switch (rpw[j + 1] - e0) {
#define M6_A(k) case k: m6_rowA<k>(A, etab, e0, n, Z, active, trow); break;
                case 1: m6_rowA<1>(A, etab, e0, n, Z, active, trow); break;
                AF_CASES(M6_A)
#undef M6_A
                default: break;
                }
                __syncthreads();
            }
            for (int i = tid; i < N; i += nt) A[i] = y[i] + A[i] * alpha;                        // sweep B (STATE 2), :4682
            __syncthreads();
            int bad = 0;
            for (int j = 0; j < b; j++) {                                                        // sweep C (STATE 3)
                const int e0 = rpw[j];
                switch (rpw[j + 1] - e0) {
#define M6_C(k) case k: bad |= m6_rowC<k>(A, etab, e0, n, Z, alpha, trow); break;
                case 1: bad |= m6_rowC<1>(A, etab, e0, n, Z, alpha, trow); break;
                AF_CASES(M6_C)
#undef M6_C
                default: break;
                }
            }
            tmem_wait_st();
            parity = __syncthreads_or(active ? bad : 0);
            if (!parity) { if (!locked) { ret = iter + 1; locked = 1; } if (!noexit) break; }    // :4761
        }
        if (!locked) ret = parity ? -iter : iter + 1;                                            // :4766
        for (int i = tid; i < N; i += nt) store_post(io, N, f, i, A[i]);
        emit_frame(g, io, f, ret, [&](int i) { return (int)(A[i] < 0); });
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase), "r"((unsigned)T.tcols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------------------------
// BP_DEC (bp_decod_qc_lm, decoders.cpp:1708-1920) and SP_DEC (sum_prod_decod_qc_lm, :1923-2185): the two flooding
// sum-product decoders, float class of the parity bar (identical decisions and iteration counts on >= 99.99 % of frames,
// posteriors within 1e-4 relative) -- so the expressions may be regrouped as long as they stay the reference's function.
// Both on the ms64 skeleton: check-to-variable messages as two TMEM columns per edge and lane, posteriors and channel
// values as doubles in shared memory, an iteration is
//   S  per check row: syndrome of the current posteriors (flat loop, CTA-wide OR; :1742-1779 / :1865-1893, :1964-1987 / :2129-2149)
//   C  per check row, no barriers: variable-to-check values of the row's edges from posterior and old message, the row's
//      combination, the new messages (row function templated on the weight: everything in registers)
//   A  block row by block row (barrier in between): posterior = channel value combined with the new messages of the bit, in
//      ascending block-row order like :1834-1862 / :2103-2127 (the first edge of a column starts from the channel value).
// BP_DEC.  The reference goes through the log domain: x_e = log|tanh(d_e / 2)| with d_e = soft - old message, s = sum x_e,
// new message = +-log((1 + A) / (1 - A)), A = exp(s - x_e) (:1790-1862): two exp, two log and two divisions per edge.
// exp(s - x_e) is the product of the other edges' |tanh|, so here T_e = |(e^d - 1) / (e^d + 1)| is kept as it is, S = prod T_e,
// A = S / T_e: one exp, one log, three divisions per edge, the same function (a row with a T_e = 0 gives 0 / 0 = NaN for that
// edge and 0 for the others exactly like exp(-inf - (-inf)) and exp(-inf) do, and the clamp turns NaN into +19.07 as the
// reference's min / max expressions do).  The chained stale syndrome (LDPCB200_BP_CHAIN_SYNDROME, :1742-1759) stays on
// the parity kernel, which runs the frames in order.
// SP_DEC (likelihood-ratio domain, no exp / log inside the loop).  The reference forms the variable-to-check value as the
// channel value times every OTHER message of the column (:2013-2040); here the column product P is kept per bit and the
// edge's own message divided out -- with the count of exactly-zero messages kept beside P, because the clamp (:2118) does
// produce zeros (a = -1) and a zero cannot be divided out.
// A message of these two decoders as it sits in tensor memory: a double in two columns (W = 2), or rounded to fp32 in one
// (W = 1: half the columns per frame, twice the frames per SM -- the kernels are latency-bound, and a message rounded to
// 24 bits moves a posterior by up to 1e-5 relative in the tests, one order below the 1e-4 of the parity bar: opt-in,
// LDPCB200_BPSP_MSG32=1; measured C4: BP_DEC 0.73 instead of 0.64 Gbit/s, SP_DEC 1.03 instead of 0.78)
template <int W> __device__ __forceinline__ double bpsp_msg(const unsigned* lw, int q)
{
    if constexpr (W == 2) return __hiloint2double((int)lw[2 * q + 1], (int)lw[2 * q]);
    else return (double)__uint_as_float(lw[q]);
}
template <int W> __device__ __forceinline__ void bpsp_put(unsigned* lw, int q, double m)
{
    if constexpr (W == 2) { lw[2 * q] = (unsigned)__double2loint(m); lw[2 * q + 1] = (unsigned)__double2hiint(m); }
    else lw[q] = __float_as_uint((float)m);
}

template <int RW, int W>
__device__ __noinline__ void bp_rowC(const double* A, const unsigned* etab, int e0, int n, int Z, unsigned trow)
{
    unsigned lw[W * RW];
    tmem_ld_n<W * RW>(trow + (unsigned)(W * e0), lw);
    double d[RW];
#pragma unroll
    for (int q = 0; q < RW; q++) {
        const unsigned pk = etab[e0 + q];
        int k = n + (int)((pk >> 16) & 0x7fffu);
        if (k >= Z) k -= Z;
        d[q] = A[(int)(pk & 0xffffu) + k];
    }
    tmem_wait_ld<W * RW>(lw);
    double S = 1.0;
    int bs = 0, bb[RW];
#pragma unroll
    for (int q = 0; q < RW; q++) {
        const double a = fx_exp(d[q] - bpsp_msg<W>(lw, q));    // :1797 (fastmath64.cuh: straight-line code)
        bb[q] = a < 1;                                                                             // :1800
        bs ^= bb[q];
        const double t = div_normal(a - 1, a + 1);                                                 // :1798, |.| below
        d[q] = t < 0 ? -t : t;
        S *= d[q];                                                                                 // exp(sum of the logs), :1810
    }
#pragma unroll
    for (int q = 0; q < RW; q++) {
        const double a = d[q] == 0.0 ? S / d[q] : div_normal(S, d[q]);                             // exp(s - x_e), :1843 (0 / 0 -> NaN like exp(NaN))
        const double den = 1 - a;
        const double r = den == 0.0 ? (1 + a) / den : div_normal(1 + a, den);                      // x / 0 must stay +inf
        double m = (1 - 2 * (bs ^ bb[q])) * fx_log(r);                                             // :1846
        m = tf_maxd(tf_mind(m, 19.07), -19.07);                                                    // :1847
        bpsp_put<W>(lw, q, m);
    }
    tmem_st_n<W * RW>(trow + (unsigned)(W * e0), lw);
}

// posterior += new message (BP) in ascending block-row order; the first edge of a column starts from the channel value (:1832)
template <int RW, int W>
__device__ __noinline__ void bp_rowA(double* A, const double* y, const unsigned* etab, int e0, int n, int Z, bool active, unsigned trow)
{
    unsigned lw[W * RW];
    tmem_ld_n<W * RW>(trow + (unsigned)(W * e0), lw);
    int idx[RW];
    double acc[RW];
#pragma unroll
    for (int q = 0; q < RW; q++) {
        const unsigned pk = etab[e0 + q];
        int k = n + (int)((pk >> 16) & 0x7fffu);
        if (k >= Z) k -= Z;
        idx[q] = (int)(pk & 0xffffu) + k;
        acc[q] = (pk >> 31) ? y[idx[q]] : A[idx[q]];
    }
    tmem_wait_ld<W * RW>(lw);
#pragma unroll
    for (int q = 0; q < RW; q++)
        if (active) A[idx[q]] = acc[q] + bpsp_msg<W>(lw, q);   // :1857
}

// syndrome of one block row's check rows: XOR over the edges of (posterior < thr) -- thr = 0 (BP_DEC) or 1 (SP_DEC)
__device__ __forceinline__ int bpsp_row_syndrome(const double* A, const unsigned* etab, int e0, int e1, int n, int Z, double thr)
{
    int synd = 0;
    for (int e = e0; e < e1; e++) {
        const unsigned pk = etab[e];
        int k = n + (int)((pk >> 16) & 0x7fffu);
        if (k >= Z) k -= Z;
        synd ^= A[(int)(pk & 0xffffu) + k] < thr;
    }
    return synd;
}

// SP_DEC row: P = column products without exact zeros, zc = number of exact zeros among the column's messages
template <int RW, int W>
__device__ __noinline__ void sp_rowC(const double* P, const unsigned char* zc, const unsigned* etab, int e0, int n, int Z, unsigned trow)
{
    unsigned lw[W * RW];
    tmem_ld_n<W * RW>(trow + (unsigned)(W * e0), lw);
    double pr[RW], z0[RW];
    int nz[RW];
#pragma unroll
    for (int q = 0; q < RW; q++) {
        const unsigned pk = etab[e0 + q];
        int k = n + (int)((pk >> 16) & 0x7fffu);
        if (k >= Z) k -= Z;
        pr[q] = P[(int)(pk & 0xffffu) + k];
        nz[q] = zc[(int)(pk & 0xffffu) + k];
    }
    tmem_wait_ld<W * RW>(lw);
    double S = 1.0;
#pragma unroll
    for (int q = 0; q < RW; q++) {
        const double own = bpsp_msg<W>(lw, q);
        // the channel value times the OTHER messages of the column (:2022-2036)
        double aa;
        if (own == 0.0) aa = nz[q] == 1 ? pr[q] : 0.0;
        else aa = nz[q] ? 0.0 : div_normal(pr[q], own);
        z0[q] = div_normal(aa - 1, aa + 1);                                                        // :2038
        S *= z0[q];                                                                                // :2044
    }
#pragma unroll
    for (int q = 0; q < RW; q++) {
        // :2111-2112 with div_normal; the two places where IEEE semantics matter are kept: z0 = 0 makes S zero or NaN, so the
        // quotient is NaN either way (-> 1.9e8 after the clamp), and x / 0 must be an infinity, not NaN
        double a = z0[q] == 0.0 ? S / z0[q] : div_normal(S, z0[q]);
        const double den = 1 - a;
        a = den == 0.0 ? (1 + a) / den : div_normal(1 + a, den);
        a = tf_maxd(tf_mind(a, 1.9e+8), -5.2e-9);                                                  // :2113
        bpsp_put<W>(lw, q, a);
    }
    tmem_st_n<W * RW>(trow + (unsigned)(W * e0), lw);
}

// posterior *= new message (SP, :2115) in ascending block-row order, as (product of the non-zero factors, number of zeros)
template <int RW, int W>
__device__ __noinline__ void sp_rowA(double* P, unsigned char* zc, const double* y, const unsigned* etab, int e0, int n, int Z, bool active, unsigned trow)
{
    unsigned lw[W * RW];
    tmem_ld_n<W * RW>(trow + (unsigned)(W * e0), lw);
    int idx[RW], cnt[RW];
    double acc[RW];
#pragma unroll
    for (int q = 0; q < RW; q++) {
        const unsigned pk = etab[e0 + q];
        int k = n + (int)((pk >> 16) & 0x7fffu);
        if (k >= Z) k -= Z;
        idx[q] = (int)(pk & 0xffffu) + k;
        acc[q] = (pk >> 31) ? y[idx[q]] : P[idx[q]];
        cnt[q] = (pk >> 31) ? 0 : zc[idx[q]];
    }
    tmem_wait_ld<W * RW>(lw);
#pragma unroll
    for (int q = 0; q < RW; q++) {
        const double m = bpsp_msg<W>(lw, q);
        if (active) {
            if (m == 0.0) zc[idx[q]] = (unsigned char)(cnt[q] + 1);
            else { zc[idx[q]] = (unsigned char)cnt[q]; P[idx[q]] = acc[q] * m; }
            if (m == 0.0 && (etab[e0 + q] >> 31)) P[idx[q]] = acc[q];
        }
    }
}

// SP = false: BP_DEC, true: SP_DEC
template <int MAXT, bool SP, int W>
__global__ void __launch_bounds__(MAXT, 1) bpsp_fast_kernel(const TaspTab T, const QcDev g, const FrameIO io)
{
    extern __shared__ __align__(16) double tf_smem[];
    const int Z = T.Z, N = T.N, E = T.E, b = T.b, nt = blockDim.x, tid = threadIdx.x;
    double* A = tf_smem;                 // posterior: BP_DEC LLR, SP_DEC product of the non-zero factors
    double* y = A + N;                   // channel values: clamped LLR (:1738) / its exponential (:1947-1951)
    unsigned* etab = (unsigned*)(y + N);
    int* rpw = (int*)(etab + E);
    unsigned* s_t = (unsigned*)(rpw + b + 1);
    unsigned char* zc = (unsigned char*)(s_t + 4);       // SP_DEC: exact zeros among a bit's factors
    const bool active = tid < Z;
    const int n = active ? tid : Z - 1;
    const bool noexit = io.flags & LDPCB200_NO_EARLY_EXIT;
    const double thr = SP ? 1.0 : 0.0;

    for (int e = tid; e < E; e += nt) {
        const int c = g.col[e];
        etab[e] = (unsigned)(c * Z) | ((unsigned)g.sh[e] << 16) | (g.cedge[g.cp[c]] == e ? 0x80000000u : 0u);
    }
    for (int j = tid; j <= b; j += nt) rpw[j] = g.rp[j];
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"((unsigned)__cvta_generic_to_shared(s_t)), "r"((unsigned)T.tcols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tbase = *(volatile unsigned*)s_t;
    const unsigned trow = __shfl_sync(0xffffffffu, tbase + ((unsigned)(((tid >> 5) & 3) * 32) << 16) + (unsigned)((tid >> 7) * W * E), 0);

    // the posterior as the reference holds it (SP_DEC: a zero factor makes it zero)
    auto post = [&](int i) -> double { if constexpr (SP) return zc[i] ? 0.0 : A[i]; else return A[i]; };
    auto syndrome = [&]() -> int {
        int bad = 0;
        for (int j = 0; j < b; j++) {
            int synd = 0;
            for (int e = rpw[j]; e < rpw[j + 1]; e++) {
                const unsigned pk = etab[e];
                int k = n + (int)((pk >> 16) & 0x7fffu);
                if (k >= Z) k -= Z;
                synd ^= post((int)(pk & 0xffffu) + k) < thr;
            }
            bad |= synd;
        }
        return __syncthreads_or(active ? bad : 0);
    };

    for (;;) {
        const int f = next_frame(io);
        if (f >= io.nf) break;
        for (int i = tid; i < N; i += nt) {
            double v = tf_maxd(tf_mind(load_llr(io, N, f, i), 20.0), -20.0);                     // :1738 / :1947-1950
            if constexpr (SP) { v = exp(v); zc[i] = 0; }
            y[i] = v; A[i] = v;
        }
        {
            // messages: BP_DEC 0 (:1732-1734), SP_DEC 1 (:1957-1959)
            unsigned init[2] = { 0u, SP ? 0x3ff00000u : 0u };
            if constexpr (W == 1) init[0] = SP ? 0x3f800000u : 0u;
            for (int e = 0; e < E; e++) TmemRow<W>::st(trow + (unsigned)(W * e), init);
            tmem_wait_st();
        }
        __syncthreads();
        int ret = 0, iter = 0;
        int parity = syndrome();                                                                 // :1742-1779 / :1964-1987
        bool locked = !parity;                                                                   // return 0: the channel values are a codeword
        while (iter < io.maxiter && (parity || noexit)) {
            for (int j = 0; j < b; j++) {                                                        // sweep C
                const int e0 = rpw[j];
                switch (rpw[j + 1] - e0) {
#define BP_C(k) case k: if constexpr (SP) sp_rowC<k, W>(A, zc, etab, e0, n, Z, trow); else bp_rowC<k, W>(A, etab, e0, n, Z, trow); break;
                case 1: if constexpr (SP) sp_rowC<1, W>(A, zc, etab, e0, n, Z, trow); else bp_rowC<1, W>(A, etab, e0, n, Z, trow); break;
                AF_CASES(BP_C)
#undef BP_C
                default: break;
                }
            }
            tmem_wait_st();
            __syncthreads();
            for (int j = 0; j < b; j++) {                                                        // sweep A
                const int e0 = rpw[j];
                switch (rpw[j + 1] - e0) {
#define BP_A(k) case k: if constexpr (SP) sp_rowA<k, W>(A, zc, y, etab, e0, n, Z, active, trow); else bp_rowA<k, W>(A, y, etab, e0, n, Z, active, trow); break;
                case 1: if constexpr (SP) sp_rowA<1, W>(A, zc, y, etab, e0, n, Z, active, trow); else bp_rowA<1, W>(A, y, etab, e0, n, Z, active, trow); break;
                AF_CASES(BP_A)
#undef BP_A
                default: break;
                }
                __syncthreads();
            }
            iter++;
            const int par = syndrome();                                                          // :1865-1893 / :2129-2149
            if (!locked) { parity = par; if (!par) { ret = iter; locked = true; } }
        }
        if (!locked) ret = -iter;                                                                // :1919 / :2184
        for (int i = tid; i < N; i += nt) store_post(io, N, f, i, post(i));
        emit_frame(g, io, f, ret, [&](int i) { return (int)(post(i) < thr); });
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase), "r"((unsigned)T.tcols) : "memory");
    }
}

size_t lms_tmem_pad_smem(size_t smem, int minb);

// decoder_id: LDPCB200_TASP_DEC, LDPCB200_LCHE_DEC, LDPCB200_ASP_DEC, LDPCB200_IASP_DEC, or LDPCB200_LMS_DEC / LDPCB200_MS_DEC (double)
FastPlan plan_tasp_fast(const QcHost& g, int decoder_id, int smem_per_sm, int smem_per_block)
{
    FastPlan p;
    const char* off = getenv("LDPCB200_NO_TASP_FAST");
    if (off && *off == '1') return p;
    if (g.maxdeg > TASP_MAXDEG || g.N > 65535 || g.Z > 1024) return p;   // (row weight 1 only for MS_DEC: map_bin needs 2, the other row functions start at 2)
    const bool iasp = decoder_id == LDPCB200_IASP_DEC;
    const bool bpsp = decoder_id == LDPCB200_BP_DEC || decoder_id == LDPCB200_SP_DEC;
    const bool ms = decoder_id == LDPCB200_MS_DEC || bpsp;               // row weight 1 allowed, two doubles per bit, every column needs an edge
    const bool asp = decoder_id == LDPCB200_ASP_DEC || iasp;
    if (!ms && g.mindeg < 2) return p;
    if (asp && g.all_cw_2) { p.note = "all columns have weight 2: the reference's shortcut arithmetic stays on the parity kernel"; return p; }
    for (int i = 0; (asp || ms) && i < g.c; i++)
        if (g.cp[i + 1] == g.cp[i]) return p;                           // sweep R starts a bit's product at its first edge
    const int zp = (g.Z + 31) & ~31;
    int tcols = 32;
    const char* m32 = getenv("LDPCB200_BPSP_MSG32");
    const bool msg32 = bpsp && m32 && *m32 == '1';                                   // BP_DEC / SP_DEC, opt-in: messages rounded to fp32 (see bpsp_msg)
    while (tcols < ((iasp || msg32) ? 1 : 2) * g.E * ((zp / 32 + 3) / 4)) tcols *= 2;   // fp64 messages take two columns, 12-bit and fp32 ones one
    if (tcols > 512) { p.note = "the lambda messages (2 columns per edge) do not fit tensor memory"; return p; }
    const size_t smem = (iasp ? sizeof(unsigned) * 3 * (size_t)g.N : sizeof(double) * ((size_t)g.N * (asp ? 3 : ms ? 2 : 1) + 96))
                        + sizeof(unsigned) * (size_t)(g.E + g.b + 1 + 4) + 16 + (bpsp ? ((size_t)g.N + 15) / 16 * 16 : 0);
    if (smem > (size_t)smem_per_block) return p;
    int m = 512 / tcols;
    m = std::min(m, (int)((size_t)smem_per_sm / (smem + 2048)));
    m = std::min(m, 2048 / zp);
    m = std::min(m, 65536 / (zp * (iasp ? (zp <= 512 ? 128 : 64) : zp <= 256 ? (bpsp ? 192 : 255) : zp <= 512 ? 128 : 64)));     // register budget of the instance (launch_tasp_fast; the BP / SP rows need fewer than 192)
    if (m < 1) m = 1;
    p.ok = 1; p.variant = 0; p.tmem = 1; p.msg32 = msg32;
    p.threads = zp; p.frames_per_cta = 1; p.ctas_per_sm = m;
    p.smem_bytes = std::min(lms_tmem_pad_smem(smem, m), (size_t)smem_per_block);
    p.tab.assign(sizeof(TaspTab), 0);
    TaspTab& T = *reinterpret_cast<TaspTab*>(p.tab.data());
    T.b = g.b; T.c = g.c; T.Z = g.Z; T.N = g.N; T.R = g.R; T.E = g.E; T.nwords = (g.N + 31) / 32; T.tcols = tcols;
    return p;
}

cudaError_t launch_tasp_multi(const FastPlan& p, const QcDev* d_gs, const FrameIO* d_ios, int n_codes, int grid_x, cudaStream_t s)
{
    const TaspTab& T = *reinterpret_cast<const TaspTab*>(p.tab.data());
    void (*kern)(const TaspTab, const QcDev*, const FrameIO*) =
        p.threads <= 128 ? tasp_multi_kernel<128> : p.threads <= 256 ? tasp_multi_kernel<256> : p.threads <= 512 ? tasp_multi_kernel<512> : tasp_multi_kernel<1024>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes);
    if (e != cudaSuccess) return e;
    kern<<<dim3(grid_x, n_codes), p.threads, p.smem_bytes, s>>>(T, d_gs, d_ios);
    return cudaGetLastError();
}

cudaError_t launch_tasp_fast(const FastPlan& p, int decoder_id, const QcDev& g, const FrameIO& io, int grid, cudaStream_t s, double alpha)
{
    const TaspTab& T = *reinterpret_cast<const TaspTab*>(p.tab.data());
    if (decoder_id == LDPCB200_BP_DEC || decoder_id == LDPCB200_SP_DEC) {
        void (*kb)(const TaspTab, const QcDev, const FrameIO);
        const bool sp = decoder_id == LDPCB200_SP_DEC, w1 = p.msg32 != 0;
#define BPSP_PICK(SPV, WV) (p.threads <= 128 ? bpsp_fast_kernel<128, SPV, WV> : p.threads <= 256 ? bpsp_fast_kernel<256, SPV, WV> : p.threads <= 512 ? bpsp_fast_kernel<512, SPV, WV> : bpsp_fast_kernel<1024, SPV, WV>)
        kb = sp ? (w1 ? BPSP_PICK(true, 1) : BPSP_PICK(true, 2)) : (w1 ? BPSP_PICK(false, 1) : BPSP_PICK(false, 2));
#undef BPSP_PICK
        cudaError_t e = cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes);
        if (e != cudaSuccess) return e;
        kb<<<grid, p.threads, p.smem_bytes, s>>>(T, g, io);
        return cudaGetLastError();
    }
    if (decoder_id == LDPCB200_MS_DEC) {
        void (*km)(const TaspTab, const QcDev, const FrameIO, const double) =
            p.threads <= 128 ? ms64_fast_kernel<128> : p.threads <= 256 ? ms64_fast_kernel<256> : p.threads <= 512 ? ms64_fast_kernel<512> : ms64_fast_kernel<1024>;
        cudaError_t e = cudaFuncSetAttribute(km, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes);
        if (e != cudaSuccess) return e;
        km<<<grid, p.threads, p.smem_bytes, s>>>(T, g, io, alpha);
        return cudaGetLastError();
    }
    // the register budget follows the CTA size: 255 registers per thread up to 256 threads
    void (*kern)(const TaspTab, const QcDev, const FrameIO);
    if (decoder_id == LDPCB200_IASP_DEC)
        kern = p.threads <= 128 ? iasp_fast_kernel<128, 4> : p.threads <= 256 ? iasp_fast_kernel<256, 2> : p.threads <= 512 ? iasp_fast_kernel<512, 1> : iasp_fast_kernel<1024, 1>;   // <= 128 registers
    else if (decoder_id == LDPCB200_ASP_DEC)
        kern = p.threads <= 128 ? asp_fast_kernel<128> : p.threads <= 256 ? asp_fast_kernel<256> : p.threads <= 512 ? asp_fast_kernel<512> : asp_fast_kernel<1024>;
    else if (decoder_id == LDPCB200_LMS_DEC)
        kern = p.threads <= 128 ? tasp_fast_kernel<128, 2> : p.threads <= 256 ? tasp_fast_kernel<256, 2> : p.threads <= 512 ? tasp_fast_kernel<512, 2> : tasp_fast_kernel<1024, 2>;
    else if (decoder_id == LDPCB200_LCHE_DEC)
        kern = p.threads <= 128 ? tasp_fast_kernel<128, 1> : p.threads <= 256 ? tasp_fast_kernel<256, 1> : p.threads <= 512 ? tasp_fast_kernel<512, 1> : tasp_fast_kernel<1024, 1>;
    else
        kern = p.threads <= 128 ? tasp_fast_kernel<128, 0> : p.threads <= 256 ? tasp_fast_kernel<256, 0> : p.threads <= 512 ? tasp_fast_kernel<512, 0> : tasp_fast_kernel<1024, 0>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes);
    if (e != cudaSuccess) return e;
    kern<<<grid, p.threads, p.smem_bytes, s>>>(T, g, io);
    return cudaGetLastError();
}

} // namespace ldpcb200
