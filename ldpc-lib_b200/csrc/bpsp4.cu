// BP_DEC (bp_decod_qc_lm, decoders.cpp:1708-1920) and SP_DEC (sum_prod_decod_qc_lm, :1923-2185) with FOUR threads per check
// row.  bpsp_fast_kernel (tasp_fast.cu) gives a check row to one thread: at Z = 81 that is three warps per frame and -- tensor
// memory holding two frames of double messages -- six warps per SM, far too few for chains of double-precision exp / log /
// division (profiles/r02_c4_bp_fast_ncu.txt: issue slots 32 %, `wait` 2.0 stalled warps per issue).  A frame's work is E * Z
// independent edge updates plus cheap per-row combinations, so here lanes 4n .. 4n+3 of a warp share check row n: part p takes
// the row's edges p, p + 4, p + 8, ...; the row's product of |tanh| (BP) / of the variable-to-check values (SP), its sign parity
// and its syndrome bit are combined with two shuffles.  The float class of the parity bar (identical decisions and iteration
// counts on >= 99.99 % of frames, posteriors within 1e-4; measured: 0 of 100 000 frames differ): bpsp_fast_kernel's product forms
// (see there), with the quotient (1 + S / T) / (1 - S / T) taken as (T + S) / (T - S), BP_DEC's logarithm of it as fx_log_ratio
// (no quotient formed: one division instead of three) and its exponential as fx_exp_tab (fastmath64.cuh; table in shared
// memory).  Same sweeps per iteration (S syndrome, C check rows, A posteriors block row by block row), messages as two TMEM
// columns per edge in the lane of the thread that owns the edge.  Four times the threads per frame on the same tensor-memory
// footprint: 22 warps per SM at C4; 130 instructions per edge, 53 of them double-precision.
#include "dec_common.cuh"
#include "lms_spec.cuh"
#include "lms_tmem.cuh"
#include "fastmath64.cuh"

namespace ldpcb200 {

struct Bpsp4Tab {
    int b, c, Z, N, R, E, nwords, tcols, cmax;     // cmax = ceil(max row weight / 4): edges per part and block row
};

__device__ __forceinline__ double b4_mind(double a, double b) { return a < b ? a : b; }
__device__ __forceinline__ double b4_maxd(double a, double b) { return a < b ? b : a; }
__device__ __forceinline__ double b4_get(const unsigned* lw, int k) { return __hiloint2double((int)lw[2 * k + 1], (int)lw[2 * k]); }
__device__ __forceinline__ void b4_put(unsigned* lw, int k, double m) { lw[2 * k] = (unsigned)__double2loint(m); lw[2 * k + 1] = (unsigned)__double2hiint(m); }

// Slot k of a block row (edges 4 k .. 4 k + 3) is empty for all four parts when the row is lighter than the code's heaviest one
// (REF-32x16-B: weights 6 .. 13, so CMAX = 4 and half of the slots of most rows): a branch that is uniform over the CTA skips it.
// Slots 0 and 1 stay unconditional so that their chains share one basic block (a row of weight <= 4 is rare).
#define B4_EMPTY(k, rw) ((k) >= 2 && 4 * (k) >= (rw))

// one block row, sweep C.  e0 .. e0 + rw: the row's edges; this thread owns edges e0 + p + 4 k, k < CMAX (those < e0 + rw)
template <int CMAX, bool SP>
__device__ __forceinline__ void b4_rowC(const double* A, const unsigned char* zc, const double* xt, const unsigned* etab, int e0, int rw, int n, int p, int Z, unsigned tcol)
{
    unsigned lw[2 * CMAX];
    tmem_ld_n<2 * CMAX>(tcol, lw);
    double d[CMAX];
    int nz[CMAX];
    bool have[CMAX];
#pragma unroll
    for (int k = 0; k < CMAX; k++) {
        const int q = p + 4 * k;
        have[k] = q < rw;
        d[k] = 1.0; nz[k] = 0;
        if (B4_EMPTY(k, rw)) continue;
        const unsigned pk = etab[e0 + (have[k] ? q : 0)];
        int pos = n + (int)((pk >> 16) & 0x7fffu);
        if (pos >= Z) pos -= Z;
        d[k] = A[(int)(pk & 0xffffu) + pos];
        nz[k] = SP ? (int)zc[(int)(pk & 0xffffu) + pos] : 0;
    }
    tmem_wait_ld<2 * CMAX>(lw);
    double S = 1.0;
    int bs = 0, bb[CMAX];
#pragma unroll
    for (int k = 0; k < CMAX; k++) {
        bb[k] = 0;
        if (B4_EMPTY(k, rw)) continue;
        if constexpr (SP) {
            const double own = b4_get(lw, k);
            double aa;                                                                   // the channel value times the OTHER messages of the column (:2022-2036)
            if (own == 0.0) aa = nz[k] == 1 ? d[k] : 0.0;
            else aa = nz[k] ? 0.0 : div_normal(d[k], own);
            d[k] = div_normal(aa - 1, aa + 1);                                           // :2038
            bb[k] = 0;
        } else {
            const double a = fx_exp_tab(d[k] - b4_get(lw, k), xt);                       // :1797 (|argument| < 700: plan_bpsp4)
            bb[k] = a < 1;                                                               // :1800
            d[k] = fabs(div_normal(a - 1, a + 1));                                       // :1798
        }
        if (have[k]) { S *= d[k]; bs ^= bb[k]; }                                         // :1810 / :2044
    }
    // the row: four parts
    S *= __shfl_xor_sync(0xffffffffu, S, 1);
    bs ^= __shfl_xor_sync(0xffffffffu, bs, 1);
    S *= __shfl_xor_sync(0xffffffffu, S, 2);
    bs ^= __shfl_xor_sync(0xffffffffu, bs, 2);
#pragma unroll
    for (int k = 0; k < CMAX; k++) {
        double m;
        if (B4_EMPTY(k, rw)) continue;
        // (1 + a) / (1 - a) with a = S / d[k] (:2111-2112, :1843-1846) as (d + S) / (d - S): one division fewer, and BP_DEC's
        // logarithm of it without forming the quotient (fx_log_ratio).  |S| <= |d[k]| (every factor is at most 1 in magnitude
        // and rounding is monotonic).  The reference's special values: d = 0 gives 0 / 0 = NaN, which its min / max turn into
        // the UPPER clamp whatever the sign; d - S = 0 gives an infinity, the clamp again -- as does every quotient beyond
        // e^19.07 = 1.9e8, so a denominator below 4e-9 of the numerator takes the clamp value without dividing.
        const double nn = d[k] + S, qq = d[k] - S;
        const bool sat = fabs(qq) <= fabs(nn) * 4e-9;
        // (no quotient is negative and none is a NaN any more, so one comparison does the reference's two-sided clamp)
        if constexpr (SP) {
            const double a = div_normal(nn, qq);
            m = (sat | !(a < 1.9e+8)) ? 1.9e+8 : a;                                      // :2113
        } else {
            const double lg = fx_log_ratio(nn, qq);
            m = (sat | !(lg < 19.07)) ? 19.07 : lg;                                      // :1847
            const int neg = (bs ^ bb[k]) & (int)(d[k] != 0.0);                           // :1846
            m = __hiloint2double(__double2hiint(m) ^ (neg << 31), __double2loint(m));
        }
        b4_put(lw, k, m);
    }
    tmem_st_n<2 * CMAX>(tcol, lw);
}

// one block row, sweep A: posterior (+)= / (*)= the new messages, the first edge of a column starts from the channel value
template <int CMAX, bool SP>
__device__ __forceinline__ void b4_rowA(double* A, unsigned char* zc, const double* y, const unsigned* etab, int e0, int rw, int n, int p, int Z, bool active, unsigned tcol)
{
    unsigned lw[2 * CMAX];
    tmem_ld_n<2 * CMAX>(tcol, lw);
    int idx[CMAX], cnt[CMAX];
    double acc[CMAX];
    bool have[CMAX], first[CMAX];
#pragma unroll
    for (int k = 0; k < CMAX; k++) {
        const int q = p + 4 * k;
        have[k] = active && q < rw;
        idx[k] = 0; cnt[k] = 0; acc[k] = 0.0; first[k] = false;
        if (B4_EMPTY(k, rw)) continue;
        const unsigned pk = etab[e0 + (q < rw ? q : 0)];
        int pos = n + (int)((pk >> 16) & 0x7fffu);
        if (pos >= Z) pos -= Z;
        idx[k] = (int)(pk & 0xffffu) + pos;
        first[k] = pk >> 31;
        acc[k] = first[k] ? y[idx[k]] : A[idx[k]];
        cnt[k] = (SP && !first[k]) ? (int)zc[idx[k]] : 0;
    }
    tmem_wait_ld<2 * CMAX>(lw);
#pragma unroll
    for (int k = 0; k < CMAX; k++) {
        if (!have[k]) continue;
        const double m = b4_get(lw, k);
        if constexpr (SP) {
            if (m == 0.0) { zc[idx[k]] = (unsigned char)(cnt[k] + 1); if (first[k]) A[idx[k]] = acc[k]; }
            else { zc[idx[k]] = (unsigned char)cnt[k]; A[idx[k]] = acc[k] * m; }        // :2115
        } else
            A[idx[k]] = acc[k] + m;                                                      // :1857
    }
}

template <int MAXT, int MINB, int CMAX, bool SP>
__global__ void __launch_bounds__(MAXT, MINB) bpsp4_kernel(const Bpsp4Tab T, const QcDev g, const FrameIO io)
{
    extern __shared__ __align__(16) double b4_smem[];
    const int Z = T.Z, N = T.N, E = T.E, b = T.b, nt = blockDim.x, tid = threadIdx.x;
    double* A = b4_smem;                 // posterior: BP_DEC LLR, SP_DEC product of the non-zero factors
    double* y = A + N;                   // channel values: clamped LLR (:1738) / its exponential (:1947-1951)
    double* xt = y + N;                  // 2^(j/32), j < 32 (fx_exp_tab)
    unsigned* etab = (unsigned*)(xt + 32);
    int* rpw = (int*)(etab + E);
    unsigned* s_t = (unsigned*)(rpw + b + 1);
    unsigned char* zc = (unsigned char*)(s_t + 4);       // SP_DEC: exact zeros among a bit's factors
    const int row = tid >> 2, p = tid & 3;
    const bool active = row < Z;
    const int n = active ? row : Z - 1;
    const bool noexit = io.flags & LDPCB200_NO_EARLY_EXIT;
    const double thr = SP ? 1.0 : 0.0;

    for (int e = tid; e < E; e += nt) {
        const int c = g.col[e];
        etab[e] = (unsigned)(c * Z) | ((unsigned)g.sh[e] << 16) | (g.cedge[g.cp[c]] == e ? 0x80000000u : 0u);
    }
    for (int j = tid; j <= b; j += nt) rpw[j] = g.rp[j];
    if (tid < 32) xt[tid] = FX_T32_DEV[tid];
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"((unsigned)__cvta_generic_to_shared(s_t)), "r"((unsigned)T.tcols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tbase = *(volatile unsigned*)s_t;
    // this thread's TMEM lane (bits 31:16) and first column: 2 CMAX columns per block row, b of them per group of four warps
    const unsigned trow = __shfl_sync(0xffffffffu, tbase + ((unsigned)(((tid >> 5) & 3) * 32) << 16) + (unsigned)((tid >> 7) * 2 * CMAX * b), 0);

    auto post = [&](int i) -> double { if constexpr (SP) return zc[i] ? 0.0 : A[i]; else return A[i]; };
    auto syndrome = [&]() -> int {
        int bad = 0;
        for (int j = 0; j < b; j++) {
            int synd = 0;
            const int e0 = rpw[j], rw = rpw[j + 1] - e0;
#pragma unroll
            for (int k = 0; k < CMAX; k++) {                                     // straight-line like the sweeps: no loop per row
                const int q = p + 4 * k;
                if (B4_EMPTY(k, rw)) continue;
                const unsigned pk = etab[e0 + (q < rw ? q : 0)];
                int pos = n + (int)((pk >> 16) & 0x7fffu);
                if (pos >= Z) pos -= Z;
                synd ^= (q < rw) & (post((int)(pk & 0xffffu) + pos) < thr);
            }
            bad |= synd << j;                                                    // (b <= 32 checked on the host)
        }
        bad ^= __shfl_xor_sync(0xffffffffu, bad, 1);
        bad ^= __shfl_xor_sync(0xffffffffu, bad, 2);
        return __syncthreads_or(active ? bad : 0);
    };

    for (;;) {
        const int f = next_frame(io);
        if (f >= io.nf) break;
        for (int i = tid; i < N; i += nt) {
            double v = b4_maxd(b4_mind(load_llr(io, N, f, i), 20.0), -20.0);                     // :1738 / :1947-1950
            if constexpr (SP) { v = exp(v); zc[i] = 0; }
            y[i] = v; A[i] = v;
        }
        {
            unsigned init[2] = { 0u, SP ? 0x3ff00000u : 0u };                                    // messages: BP_DEC 0 (:1732-1734), SP_DEC 1 (:1957-1959)
            for (int k = 0; k < CMAX * b; k++) TmemRow<2>::st(trow + 2u * (unsigned)k, init);
            tmem_wait_st();
        }
        __syncthreads();
        int ret = 0, iter = 0;
        int parity = syndrome();                                                                 // :1742-1779 / :1964-1987
        bool locked = !parity;
        while (iter < io.maxiter && (parity || noexit)) {
            for (int j = 0; j < b; j++)                                                          // sweep C
                b4_rowC<CMAX, SP>(A, zc, xt, etab, rpw[j], rpw[j + 1] - rpw[j], n, p, Z, trow + (unsigned)(2 * CMAX * j));
            tmem_wait_st();
            __syncthreads();
            for (int j = 0; j < b; j++) {                                                        // sweep A
                b4_rowA<CMAX, SP>(A, zc, y, etab, rpw[j], rpw[j + 1] - rpw[j], n, p, Z, active, trow + (unsigned)(2 * CMAX * j));
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

// decoder_id: LDPCB200_BP_DEC or LDPCB200_SP_DEC
FastPlan plan_bpsp4(const QcHost& g, int decoder_id, int smem_per_sm, int smem_per_block)
{
    FastPlan p;
    const char* off = getenv("LDPCB200_NO_BPSP4");               // 1: one thread per check row (tasp_fast.cu bpsp_fast_kernel)
    if (off && *off == '1') return p;
    off = getenv("LDPCB200_NO_TASP_FAST");                       // 1: the table-driven parity kernels
    if (off && *off == '1') return p;
    if (g.maxdeg > 20 || g.N > 65535 || g.b > 32) return p;
    for (int i = 0; i < g.c; i++)
        if (g.cp[i + 1] == g.cp[i]) return p;                           // sweep A starts a bit's posterior at its first edge
    for (int i = 0; i < g.c; i++)
        if (g.cp[i + 1] - g.cp[i] > 35) return p;                       // |posterior| <= 20 + 19.07 * weight must stay below fx_exp_tab's 700
    const int cmax = (g.maxdeg + 3) / 4;
    const int threads = ((4 * g.Z + 31) / 32) * 32;
    if (threads > 1024) return p;
    int tcols = 32;
    while (tcols < 2 * cmax * g.b * ((threads / 32 + 3) / 4)) tcols *= 2;
    if (tcols > 512) return p;
    const size_t smem = sizeof(double) * (2 * (size_t)g.N + 48) + sizeof(unsigned) * (size_t)(g.E + g.b + 1 + 4) + 16 + ((size_t)g.N + 15) / 16 * 16;
    if (smem > (size_t)smem_per_block) return p;
    int m = 512 / tcols;
    m = std::min(m, (int)((size_t)smem_per_sm / (smem + 2048)));
    m = std::min(m, 2048 / threads);
    m = std::min(m, threads <= 512 ? 2 : 1);                           // register budget of the instance (b4_launch)
    const char* one = getenv("LDPCB200_BPSP4_ONE_CTA");               // development: one CTA per SM (385 .. 512 threads: the instance with 128 registers)
    if (one && *one == '1') m = 1;
    if (m < 1) m = 1;
    p.ok = 1; p.variant = 0; p.tmem = 1; p.msg32 = 0; p.bpsp4 = 1;
    p.threads = threads; p.frames_per_cta = 1; p.ctas_per_sm = m;
    p.smem_bytes = std::min(lms_tmem_pad_smem(smem, m), (size_t)smem_per_block);
    p.tab.assign(sizeof(Bpsp4Tab), 0);
    Bpsp4Tab& T = *reinterpret_cast<Bpsp4Tab*>(p.tab.data());
    T.b = g.b; T.c = g.c; T.Z = g.Z; T.N = g.N; T.R = g.R; T.E = g.E; T.nwords = (g.N + 31) / 32; T.tcols = tcols; T.cmax = cmax;
    return p;
}

template <int CMAX, bool SP>
static cudaError_t b4_launch(const FastPlan& p, const Bpsp4Tab& T, const QcDev& g, const FrameIO& io, int grid, cudaStream_t s)
{
    // two CTAs per SM up to 512 threads: the launch bound caps the registers accordingly
    void (*k)(const Bpsp4Tab, const QcDev, const FrameIO) =
        p.threads <= 256 ? bpsp4_kernel<256, 2, CMAX, SP> : p.threads <= 384 ? bpsp4_kernel<384, 2, CMAX, SP>
        : p.threads <= 512 ? (p.ctas_per_sm >= 2 ? bpsp4_kernel<512, 2, CMAX, SP> : bpsp4_kernel<512, 1, CMAX, SP>)
        : p.threads <= 768 ? bpsp4_kernel<768, 1, CMAX, SP> : bpsp4_kernel<1024, 1, CMAX, SP>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes);
    if (e != cudaSuccess) return e;
    k<<<grid, p.threads, p.smem_bytes, s>>>(T, g, io);
    return cudaGetLastError();
}

cudaError_t launch_bpsp4(const FastPlan& p, int decoder_id, const QcDev& g, const FrameIO& io, int grid, cudaStream_t s)
{
    const Bpsp4Tab& T = *reinterpret_cast<const Bpsp4Tab*>(p.tab.data());
    const bool sp = decoder_id == LDPCB200_SP_DEC;
    switch (T.cmax) {
    case 1: return sp ? b4_launch<1, true>(p, T, g, io, grid, s) : b4_launch<1, false>(p, T, g, io, grid, s);
    case 2: return sp ? b4_launch<2, true>(p, T, g, io, grid, s) : b4_launch<2, false>(p, T, g, io, grid, s);
    case 3: return sp ? b4_launch<3, true>(p, T, g, io, grid, s) : b4_launch<3, false>(p, T, g, io, grid, s);
    case 4: return sp ? b4_launch<4, true>(p, T, g, io, grid, s) : b4_launch<4, false>(p, T, g, io, grid, s);
    default: return sp ? b4_launch<5, true>(p, T, g, io, grid, s) : b4_launch<5, false>(p, T, g, io, grid, s);
    }
}

} // namespace ldpcb200
