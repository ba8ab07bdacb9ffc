// IMS_DEC (imin_sum_decod_qc_lm, decoders.cpp:5430-5690, fixed point, bit-exact) with TWO frames packed into the two fp16
// halves of every 32-bit word: the low half of a register, a shared-memory word or a tensor-memory word belongs to one
// frame, the high half to another, and every arithmetic instruction (HADD2 / HFMA2 / HMNMX2 / VHMNMX, sign-bit LOP3) works
// on both frames at once.  Schedule, summation order and storage are those of ms_tmem.cuh (pass A: acc += message block row
// by block row; pass B: soft = sat(iy + acc); pass C per check row: v2c, syndrome, new messages as minima over the OTHER
// edges; messages in TENSOR MEMORY, accumulators / posteriors as doubled columns in the rotation of their last writer).
//
// Why this is exact: every IMS_DEC quantity is a small integer -- |iy| <= max_quant, |soft|, |acc| <= max_data = 2^(dbits-1) - 1,
// |v2c| <= 2 max_data, messages <= max_data -- and binary16 represents every integer up to 2048 and adds, subtracts and
// compares them exactly; the host only selects this kernel for dbits <= 8 and 0 <= ialpha <= 16.  The one rounding step of
// the reference, (min * ialpha) >> 4 (:5554, :5640), is floor(m * a / 16):
//     t = fma(m + d, a / 16, 1025)      d = -k / 16 with 112 / a < k < 128 / a  (a = 12: d = -0.625)
// m + d is exact (m < 128: the half grid there is 1/16), the fma's exact value is 1025 + q + (r / 16 + d a / 16) with
// m a = 16 q + r and d a / 16 strictly between -1/2 and -7/16, so the bracket lies strictly inside (-1/2, 1/2) and rounding to
// the integer grid of [1024, 2048) gives 1025 + q -- never a tie; a second fma with the row's sign (+-1) takes the constant
// off: +-(t - 1025) = +-q.  (1025, not 1024: with q = 0 the sum stays above 1024, where the grid is 1.)
// tests/test_message_forms.py checks the identity for every (m, a); tests/test_gpu_ms_spec.py checks ims_y, ims_soft, decisions
// and iteration counts against the oracle and the register-compressed kernel.
//
// Frames are independent and stop at different iterations (:5680-5685), so the halves are SLOTS: a CTA owns 2 G of them (G
// groups of ZP threads -- two groups when Z <= 128, so that a CTA has eight warps that walk the unrolled block rows
// together and share the instruction cache), every iteration runs on all of them, and a slot whose frame stops (syndrome
// zero, or the iteration limit) is written out at that moment and refilled with the next frame of the batch: its half of
// the channel words is rewritten, its half of the message words in tensor memory is cleared, its iteration count starts
// over, while the other slots carry on.  The accumulators / posteriors are rebuilt from the messages in every iteration
// (pass A starts from 0), so nothing else has to be reset.  A slot without a frame keeps computing on bounded values and
// is ignored.
// This header must stay free of #include (NVRTC), and follows ms_tmem.cuh in the translation unit.
#pragma once

namespace ldpcb200 {

// packed binary16 arithmetic on 32-bit registers (PTX, so that NVRTC needs no cuda_fp16.h)
static __device__ __forceinline__ unsigned h2_add(unsigned a, unsigned b) { unsigned d; asm("add.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
static __device__ __forceinline__ unsigned h2_sub(unsigned a, unsigned b) { unsigned d; asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
static __device__ __forceinline__ unsigned h2_mul(unsigned a, unsigned b) { unsigned d; asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
static __device__ __forceinline__ unsigned h2_fma(unsigned a, unsigned b, unsigned c) { unsigned d; asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
static __device__ __forceinline__ unsigned h2_fma_relu(unsigned a, unsigned b, unsigned c) { unsigned d; asm("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
static __device__ __forceinline__ unsigned h2_min(unsigned a, unsigned b) { unsigned d; asm("min.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
static __device__ __forceinline__ unsigned h2_max(unsigned a, unsigned b) { unsigned d; asm("max.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
static __device__ __forceinline__ unsigned h2_abs(unsigned a) { unsigned d; asm("abs.f16x2 %0, %1;" : "=r"(d) : "r"(a)); return d; }
// {hi, lo} as binary16 (both values are integers below 2048 here: exact)
static __device__ __forceinline__ unsigned h2_pack(float lo, float hi) { unsigned d; asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo)); return d; }
static __device__ __forceinline__ float h2_get(unsigned w, int half)
{
    unsigned short h = (unsigned short)(half ? w >> 16 : w & 0xffffu);
    float f;
    asm("cvt.f32.f16 %0, %1;" : "=f"(f) : "h"(h));
    return f;
}

template <class K, int G>
struct ImsH2 {
    typedef LmsTmem<K> T;
    static constexpr int B = K::B, C = K::C, Z = K::Z, N = K::C * K::Z, R = K::B * K::Z, ZP = K::ZP, E = K::E;
    static constexpr int NT = G * ZP;                           // threads per CTA
    static constexpr int S = 2 * G;                             // slots = frames in flight per CTA
    static constexpr int NWORDS = (N + 31) / 32;
    static constexpr bool ALL_ACTIVE = (Z == ZP);
    static constexpr int CS = 2 * Z;
    static constexpr int NWARPS = NT / 32;
    static __host__ __device__ constexpr int pow2_at_least(int n) { int t = 32; while (t < n) t *= 2; return t; }
    static constexpr int TCOLS = pow2_at_least(E * ((NT / 32 + 3) / 4));
    // shared memory (words; one word = the two slots' halves), per group: doubled accumulators / posteriors (between
    // iterations also the staging of a new frame's fp32 LLRs: N of its 2 N words) | quantised channel values (minus
    // max_data, see pass A); then mbarrier | misc
    static constexpr int Y_OFF = C * CS;
    static constexpr int GROUP_WORDS = (Y_OFF + N + 1) & ~1;
    static constexpr int MBAR_OFF = G * GROUP_WORDS;
    static constexpr int MISC_OFF = MBAR_OFF + 2;
    static constexpr int SMEM_WORDS = MISC_OFF + 16 + 2 * (NT / 32 + 2);     // misc | per-warp energy sums + coef (doubles)

    struct Consts { unsigned cap, ncap, cap2, scale, dshift, magic, sone; };

    // ---- pass A, block row J: acc[bit] = sat(acc[bit] + message), ascending block rows per bit (:5540-5576).  The
    // accumulators are kept BIASED by max_data (0 .. 2 max_data), so that the saturation is one fma with a built-in
    // max(., 0) and one minimum instead of an add and two comparisons; pass B takes the bias off (y holds iy - max_data).
    template <int J, int Q>
    static __device__ __forceinline__ void accA_load(const unsigned* softn, unsigned (&acc)[K::RP[J + 1] - K::RP[J]], unsigned cap)
    {
        constexpr int E0 = K::RP[J], DEG = K::RP[J + 1] - K::RP[J];
        if constexpr (Q < DEG) {
            constexpr int off = K::COL[E0 + Q] * CS + K::DELTA[E0 + Q];
            if constexpr (K::FIRST[E0 + Q]) acc[Q] = cap;                                        // 0 + bias, :5536
            else acc[Q] = softn[off];
            accA_load<J, Q + 1>(softn, acc, cap);
        }
    }
    template <int J, int Q>
    static __device__ __forceinline__ void accA_store(unsigned* softn, bool active, const unsigned (&acc)[K::RP[J + 1] - K::RP[J]])
    {
        constexpr int E0 = K::RP[J], DEG = K::RP[J + 1] - K::RP[J];
        if constexpr (Q < DEG) {
            constexpr int off = K::COL[E0 + Q] * CS;
            if (ALL_ACTIVE || active) { softn[off] = acc[Q]; softn[off + Z] = acc[Q]; }
            accA_store<J, Q + 1>(softn, active, acc);
        }
    }
    template <int J>
    static __device__ __forceinline__ void passA(unsigned* softn, unsigned trow, unsigned mbar, unsigned& ph, bool lane0, bool active, const Consts& k)
    {
        if constexpr (J < B) {
            constexpr int E0 = K::RP[J], DEG = K::RP[J + 1] - K::RP[J];
            unsigned msg[DEG], acc[DEG];
            tmem_ld_n<DEG>(trow + E0, msg);
            accA_load<J, 0>(softn, acc, k.cap);
            tmem_wait_ld<DEG>(msg);
#pragma unroll
            for (int q = 0; q < DEG; q++) acc[q] = h2_min(h2_fma_relu(acc[q], k.sone, msg[q]), k.cap2);       // :5567-5568
            T::loads_done(mbar, lane0);
            if constexpr (B % 2 == 0) T::wait_loads(mbar, J & 1);
            else { T::wait_loads(mbar, ph); ph ^= 1u; }
            accA_store<J, 0>(softn, active, acc);
            __syncthreads();
            passA<J + 1>(softn, trow, mbar, ph, lane0, active, k);
        }
    }

    // ---- pass B, position tid of every column: soft = sat(iy + acc) (:5594-5603) = sat((iy - max_data) + biased acc); the column
    // is rotated by K::ROT, the channel values are not
    template <int COL>
    static __device__ __forceinline__ void passB(unsigned* softn, const unsigned* y, int tid, const Consts& k)
    {
        if constexpr (COL < C) {
            constexpr int rot = K::ROT[COL];
            int p = tid + rot;
            if (rot != 0 && p >= Z) p -= Z;
            const unsigned s = h2_max(h2_min(h2_add(y[COL * Z + p], softn[COL * CS]), k.cap), k.ncap);
            softn[COL * CS] = s; softn[COL * CS + Z] = s;
            passB<COL + 1>(softn, y, tid, k);
        }
    }

    // ---- pass C, block row J (:5608-5678)
    template <int J, int Q>
    static __device__ __forceinline__ void softC_load(const unsigned* softn, unsigned (&rs)[K::RP[J + 1] - K::RP[J]])
    {
        constexpr int E0 = K::RP[J], DEG = K::RP[J + 1] - K::RP[J];
        if constexpr (Q < DEG) {
            constexpr int off = K::COL[E0 + Q] * CS + K::SYNSH[E0 + Q];
            rs[Q] = softn[off];
            softC_load<J, Q + 1>(softn, rs);
        }
    }
    // m[q] = min(cap, min over p != q of |v[p]|), both frames at once: groups of three edges, one three-input minimum per group
    // (two dependent min.f16x2 become one VHMNMX), a "rest" term per group (the other groups and the ceiling) and one
    // three-input minimum per edge over its two group mates and the rest -- lms_tmem.cuh min_of_others on fp16 pairs
    static __device__ __forceinline__ unsigned min3(unsigned a, unsigned b, unsigned c) { return h2_min(h2_min(a, b), c); }
    template <int DEG>
    static __device__ __forceinline__ void min_of_others(const unsigned (&v)[DEG], unsigned (&m)[DEG], unsigned cap)
    {
        constexpr int NG = (DEG + 2) / 3;
        unsigned a[DEG], g[NG], rest[NG];
#pragma unroll
        for (int q = 0; q < DEG; q++) a[q] = h2_abs(v[q]);
#pragma unroll
        for (int k = 0; k < NG; k++) {
            const int i = 3 * k;
            g[k] = i + 2 < DEG ? min3(a[i], a[i + 1], a[i + 2]) : i + 1 < DEG ? h2_min(a[i], a[i + 1]) : a[i];
        }
        if constexpr (NG == 1) rest[0] = cap;
        else if constexpr (NG == 2) { rest[0] = h2_min(g[1], cap); rest[1] = h2_min(g[0], cap); }
        else if constexpr (NG == 3) { rest[0] = min3(g[1], g[2], cap); rest[1] = min3(g[0], g[2], cap); rest[2] = min3(g[0], g[1], cap); }
        else {
            unsigned pre[NG], suf[NG];                           // minima of the groups before / after group k
            pre[0] = cap; suf[NG - 2] = g[NG - 1];
#pragma unroll
            for (int k = 1; k < NG; k++) pre[k] = h2_min(pre[k - 1], g[k - 1]);
#pragma unroll
            for (int k = NG - 3; k >= 0; k--) suf[k] = h2_min(suf[k + 1], g[k + 1]);
#pragma unroll
            for (int k = 0; k < NG - 1; k++) rest[k] = h2_min(pre[k], suf[k]);
            rest[NG - 1] = pre[NG - 1];
        }
#pragma unroll
        for (int k = 0; k < NG; k++) {
            const int i = 3 * k;
            if (i + 2 < DEG) {
                m[i] = min3(a[i + 1], a[i + 2], rest[k]);
                m[i + 1] = min3(a[i], a[i + 2], rest[k]);
                m[i + 2] = min3(a[i], a[i + 1], rest[k]);
            } else if (i + 1 < DEG) {
                m[i] = h2_min(a[i + 1], rest[k]);
                m[i + 1] = h2_min(a[i], rest[k]);
            } else m[i] = rest[k];
        }
    }
    template <int J>
    static __device__ __forceinline__ void passC(const unsigned* softn, unsigned trow, const Consts& k, unsigned& bad)
    {
        if constexpr (J < B) {
            constexpr int E0 = K::RP[J], DEG = K::RP[J + 1] - K::RP[J];
            unsigned msg[DEG], rs[DEG], v[DEG], m[DEG];
            tmem_ld_n<DEG>(trow + E0, msg);
            softC_load<J, 0>(softn, rs);
            tmem_wait_ld<DEG>(msg);
            unsigned synd = 0, sacc = 0;
#pragma unroll
            for (int q = 0; q < DEG; q++) {
                synd ^= rs[q];                                                                   // :5631 (sign bits = rs < 0, both frames)
                v[q] = h2_sub(rs[q], msg[q]);                                                    // :5646, the message is already scaled (:5640)
                sacc ^= v[q];
            }
            // the row's sign product as +-1 per frame, and -+1025 to take the rounding constant off again
            const unsigned rone = (sacc & 0x80008000u) | k.sone;
            const unsigned nmag = h2_mul(rone, k.magic ^ 0x80008000u);
            min_of_others<DEG>(v, m, k.cap);                                                     // :5653, :5656-5666
#pragma unroll
            for (int q = 0; q < DEG; q++) {
                const unsigned t = h2_fma(h2_add(m[q], k.dshift), k.scale, k.magic);             // 1025 + ((m * ialpha) >> 4), :5554
                msg[q] = h2_fma(t, rone, nmag) ^ (v[q] & 0x80008000u);                            // row sign, then the edge's own sign
            }
            bad |= synd;
            tmem_st_n<DEG>(trow + E0, msg);                                                      // :5675
            passC<J + 1>(softn, trow, k, bad);
        }
    }

    // channel LLRs of frame f as fp32 into dst[N] (the first load of ms_tmem.cuh; all NT threads)
    static __device__ __forceinline__ void load_frame(const FrameIO& io, int f, float* dst, int tid)
    {
        if (io.ch.enabled) {
            const unsigned long long frame = io.ch.first_frame + (unsigned long long)f;
            if (io.ch.m > 2) {
                const int half = io.ch.m >> 1, ncomp = 2 * (N / io.ch.m);
                for (int c = tid; c < ncomp; c += NT) {
                    float o[4];
                    channel_llr_qam_component(io.ch, frame, c, o);
                    const int i0 = (c >> 1) * io.ch.m + (c & 1) * half;
                    for (int b = 0; b < half; b++) {
                        const int i = channel_dest(io.ch, i0 + b);
                        dst[i] = i >= io.ch.punct_start ? io.ch.punct_value : o[b];
                    }
                }
            } else {
                for (int i4 = tid; i4 < N / 4; i4 += NT) {                                // one Philox block -> four LLRs
                    float o[4];
                    int d[4];
                    channel_llr4_bpsk(io.ch, frame, i4, o, d);
#pragma unroll
                    for (int b = 0; b < 4; b++) dst[d[b]] = o[b];
                }
                for (int j = (N & ~3) + tid; j < N; j += NT) { const int i = channel_dest(io.ch, j); dst[i] = channel_llr(io.ch, frame, i); }
            }
        } else if (io.llr_dtype == 1) {
            const float* src = (const float*)io.llr + (size_t)f * N;
            for (int i = tid; i < N; i += NT) dst[i] = __ldcs(src + i);
        } else {
            const double* src = (const double*)io.llr + (size_t)f * N;
            for (int i = tid; i < N; i += NT) dst[i] = (float)__ldcs(src + i);
        }
    }
    // the quantiser of :5481-5500 for one value, in the reference's double arithmetic.  *near: the value before the floor lies
    // within 1e-9 of an integer (see the energy note in refill())
    static __device__ __noinline__ int quantise_exact(double val, double coef, const MsSpecParams& sp, bool* near)
    {
        int sign = 0;
        if (val < 0) { val = -val; sign = 1; }
        val *= coef;
        if (val > sp.thr) val = sp.thr;
        const double u = div_normal(val * sp.max_quant, sp.thr) + 0.5;   // the correctly rounded quotient without the slow-path branch (channel.cuh)
        const double fl = floor(u);
        if (u - fl < 1e-9 || u - fl > 1.0 - 1e-9) *near = true;
        const int ival = (short)fl;
        return sign ? -ival : ival;
    }
    // The same value from an fp32 estimate whenever that cannot be wrong: t = |y| coef max_quant / thr is at most max_quant
    // (<= 127), the estimate is within 1e-4 of it (a few fp32 roundings), so floor(t + 0.5) is decided unless t + 0.5 lies
    // within 1e-3 of an integer -- one value in 500 -- and only then the double arithmetic runs.  cs = coef max_quant / thr.
    static __device__ __forceinline__ int quantise(double val, float valf, double coef, float cs, const MsSpecParams& sp, bool* near)
    {
        const float u = fminf(fabsf(valf) * cs, (float)sp.max_quant) + 0.5f;
        const float fl = floorf(u), frac = u - fl;
        if (!(frac >= 1e-3f && frac <= 1.0f - 1e-3f)) return quantise_exact(val, coef, sp, near);
        const int ival = (int)fl;
        return valf < 0.0f ? -ival : ival;
    }

    // clear half `keep ^ 0xffffffff` of N consecutive tensor-memory columns, 16 at a time
    template <int NC, int OFF = 0>
    static __device__ __forceinline__ void tmem_mask_n(unsigned t, unsigned keep)
    {
        if constexpr (OFF < NC) {
            constexpr int P = tmem_chunk(NC - OFF);
            unsigned r[P];
            TmemRow<P>::ld(t + OFF, r);
            tmem_wait_ld<P>(r);
#pragma unroll
            for (int i = 0; i < P; i++) r[i] &= keep;
            TmemRow<P>::st(t + OFF, r);
            tmem_mask_n<NC, OFF + P>(t, keep);
        }
    }

    // outputs of frame f = half `h` of the posteriors `gsoft` of its group (the whole CTA calls this)
    static __device__ __noinline__ void emit(const FrameIO& io, const unsigned* gsoft, int* s_misc, int f, int h, int ret, int tid)
    {
        if (io.post) {
            for (int i = tid; i < N; i += NT) {
                const int col = i / Z, p = i - col * Z;
                ((short*)io.post)[(size_t)f * N + i] = (short)(int)h2_get(gsoft[col * CS + p + K::rt_ri()[col]], h);
            }
        }
        const int lane = tid & 31;
        const unsigned sbit = h ? 0x80000000u : 0x8000u;
        int nerr = 0, nerr_info = 0;
        constexpr int NROUND = (N + 31) & ~31;
        if (tid == 0) { s_misc[1] = 0; s_misc[2] = 0; }
        __syncthreads();
        for (int i = tid; i < NROUND; i += NT) {
            int bit = 0;
            if (i < N) {
                const int col = i / Z, p = i - col * Z;
                bit = (gsoft[col * CS + p + K::rt_ri()[col]] & sbit) != 0u;
            }
            const unsigned w = __ballot_sync(0xffffffffu, bit);
            if (lane == 0) {
                if (io.hard_words) io.hard_words[(size_t)f * NWORDS + (i >> 5)] = w;
                nerr += __popc(w);
                const int lo = R - i;
                const unsigned wi = lo <= 0 ? w : (lo >= 32 ? 0u : (w >> lo) << lo);
                nerr_info += __popc(wi);
            }
        }
        if (lane == 0 && nerr) { atomicAdd(&s_misc[1], nerr); atomicAdd(&s_misc[2], nerr_info); }
        __syncthreads();
        if (tid == 0) {
            const int e = s_misc[1], ei = s_misc[2];
            if (io.iters) io.iters[f] = ret;
            if (io.per_frame)
                io.per_frame[f] = (e ? 0x80000000u : 0u) | (ret >= 0 ? 0x40000000u : 0u) | (unsigned)(ei < 0xFFFFFF ? ei : 0xFFFFFF);
            if (io.counters) {
                atomicAdd(&io.counters[0], 1ull);
                atomicAdd(&io.counters[4], (unsigned long long)(ret < 0 ? -ret : ret));
                if (e) {
                    atomicAdd(&io.counters[1], 1ull);
                    atomicAdd(&io.counters[2], (unsigned long long)ei);
                    atomicAdd(&io.counters[5], (unsigned long long)e);
                    if (ret >= 0) atomicAdd(&io.counters[3], 1ull);
                }
            }
        }
        __syncthreads();
    }

    // Put the next frame of the batch into half `h` of group `gs` (the whole CTA calls this, between iterations): fp32 LLRs
    // staged in the group's posterior area, quantised (:5472-5500) into its half of
    // the channel words, its half of the group's message words cleared (dcs[] = 0, :5463-5502).  -> frame index, or -1 when
    // the batch is exhausted (the slot then keeps its old, bounded contents).
    // Tickets are drawn one refill ahead (thread 0 carries the next one): the atomic issued here is only waited for when
    // the function returns, behind the whole refill.  -> (frame, next ticket)
    static __device__ __noinline__ int2 refill(const FrameIO& io, const MsSpecParams& sp, unsigned* soft2, int* s_misc, unsigned trow, int gs, int h, int tid, int ticket)
    {
        __syncthreads();
        int next = 0;
        if (tid == 0) { s_misc[0] = ticket; next = (int)atomicAdd(io.next_frame, 1u); }
        __syncthreads();
        const int f = s_misc[0];
        if (f >= io.nf) return make_int2(-1, next);
        unsigned* gsoft = soft2 + gs * GROUP_WORDS;
        unsigned* gy = gsoft + Y_OFF;
        float* stage = (float*)gsoft;
        load_frame(io, f, stage, tid);
        const unsigned keep = h ? 0x0000ffffu : 0xffff0000u;
        tmem_wait_st();                                                            // the last pass C's messages are in place
        if (tid / ZP == gs) tmem_mask_n<E>(trow, keep);
        __syncthreads();
        const bool f64 = !io.ch.enabled && io.llr_dtype == 0;                       // doubles are quantised as doubles, not through fp32
        const double* src = f64 ? (const double*)io.llr + (size_t)f * N : nullptr;
        // Per-frame energy normalisation coef = sqrt(N / sum y_i^2) (:5472-5479).  The reference adds the squares one after the
        // other in double, and floating-point addition is not associative -- but ANY summation order of N non-negative terms is
        // within 2 N 2^-53 < 1e-12 (relative) of the sequential one, which moves t = |y| coef max_quant / thr (<= 127) by less
        // than 1.3e-10.  So: sum in parallel, quantise, and note whether any t + 0.5 came within 1e-9 of an integer.  If none
        // did (all but one frame in 10^6), every floor() is the reference's whatever the order; otherwise thread 0 adds the
        // squares in the reference's order and the frame is quantised again.  Exact always, sequential almost never -- and the
        // energy pre-pass kernel (a second generation of every noise sample) is gone.
        double* s_red = (double*)(s_misc + 16);
        for (int pass = 0; pass < 2; pass++) {
            if (pass == 0) {
                double part = 0.0;
                for (int i = tid; i < N; i += NT) { const double v = src ? src[i] : (double)stage[i]; part += v * v; }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                if ((tid & 31) == 0) s_red[tid >> 5] = part;
                __syncthreads();
                if (tid == 0) { double en = 0.0; for (int w = 0; w < NWARPS; w++) en += s_red[w]; s_red[NWARPS] = sqrt(N / en); }
            } else {
                if (tid == 0) { double en = 0.0; for (int i = 0; i < N; i++) { const double v = src ? src[i] : (double)stage[i]; en += v * v; } s_red[NWARPS] = sqrt(N / en); }   // :5472-5479
            }
            __syncthreads();
            const double coef = s_red[NWARPS];
            const float cs = (float)(coef * sp.max_quant / sp.thr);
            bool near = false;
            for (int i = tid; i < N; i += NT) {
                const int q = src ? quantise(src[i], (float)src[i], coef, cs, sp, &near) : quantise((double)stage[i], stage[i], coef, cs, sp, &near);
                const unsigned hv = h2_pack((float)(q - sp.max_data), 0.0f) & 0xffffu;    // iy - max_data: pass B
                gy[i] = (gy[i] & keep) | (h ? hv << 16 : hv);
                if (io.aux) io.aux[(size_t)f * N + i] = (short)q;
            }
            if (pass == 1 || !__syncthreads_or(near)) break;
        }
        tmem_wait_st();
        __syncthreads();
        return make_int2(f, next);
    }

    static __device__ __forceinline__ void kernel(const FrameIO& io, const MsSpecParams& sp)
    {
        extern __shared__ __align__(16) unsigned soft2u[];
        unsigned* soft2 = soft2u;
        int* s_misc = (int*)(soft2 + MISC_OFF);                  // [0] frame ticket, [1..2] error counts, [4 .. 4+S) parity flags
        const int tid = threadIdx.x;
        const int g = G == 1 ? 0 : tid / ZP, tg = tid - g * ZP;  // group, thread within the group = check row lane
        const bool active = tg < Z;
        const bool lane0 = (tid & 31) == 0;
        const bool noexit = io.flags & 8u;                       // LDPCB200_NO_EARLY_EXIT
        unsigned* gsoft = soft2 + g * GROUP_WORDS;
        unsigned* gy = gsoft + Y_OFF;
        unsigned* softn = gsoft + tg;
        const unsigned mbar = (unsigned)__cvta_generic_to_shared(soft2 + MBAR_OFF);
        unsigned ph = 0;

        Consts k;
        {
            const float cap = (float)sp.max_data;
            k.cap = h2_pack(cap, cap);
            k.ncap = k.cap ^ 0x80008000u;
            k.cap2 = h2_pack(2.0f * cap, 2.0f * cap);
            const float sc = (float)sp.ialpha * 0.0625f;
            k.scale = h2_pack(sc, sc);
            // d = -kk / 16, 112 / a < kk < 128 / a (none needed for a = 0 and a = 16: the product is an integer)
            const int a = sp.ialpha;
            const int kk = (a <= 0 || a >= 16) ? 0 : 112 / a + 1;
            const float d = -(float)kk * 0.0625f;
            k.dshift = h2_pack(d, d);
            k.magic = h2_pack(1025.0f, 1025.0f);
            k.sone = h2_pack(1.0f, 1.0f);
        }

        if (tid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar), "r"((unsigned)NWARPS) : "memory");
        if (tid < 32) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                         :: "r"((unsigned)__cvta_generic_to_shared(s_misc)), "r"((unsigned)TCOLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        for (int i = tid; i < G * GROUP_WORDS; i += NT) soft2[i] = 0u;
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const unsigned tbase = *(volatile unsigned*)s_misc;
        // this thread's lane (bits 31:16: 32 * (warp mod 4) + lane) and first column (warps 4c .. 4c+3 own columns c E ..)
        const unsigned trow = __shfl_sync(0xffffffffu, tbase + ((unsigned)(((tid >> 5) & 3) * 32) << 16) + (unsigned)((tid >> 7) * E), 0);
        tmem_zero_n<E>(trow);
        tmem_wait_st();
        __syncthreads();
        if (tid < S) s_misc[4 + tid] = 0;
        int ticket = 0;
        if (tid == 0) ticket = (int)atomicAdd(io.next_frame, 1u);

        if (io.maxiter <= 0) {
            // no pass runs: decisions and "posteriors" are the quantised channel values, the return value is 0
            for (;;) {
                const int2 rf = refill(io, sp, soft2, s_misc, trow, 0, 0, tid, ticket);
                const int f = rf.x;
                ticket = rf.y;
                if (f < 0) break;
                for (int i = tid; i < N; i += NT) {
                    const int col = i / Z, p = i - col * Z;
                    int q = p + K::rt_rot()[col];
                    if (q >= Z) q -= Z;
                    const unsigned w = h2_add(soft2[Y_OFF + col * Z + q], k.cap);
                    soft2[col * CS + p] = w; soft2[col * CS + p + Z] = w;
                }
                __syncthreads();
                emit(io, soft2, s_misc, f, 0, 0, tid);
            }
        } else {
            int sf[S], sit[S], sret[S];
            bool sset[S];
            int live = 0;
#pragma unroll
            for (int s = 0; s < S; s++) {
                const int2 rf = refill(io, sp, soft2, s_misc, trow, s >> 1, s & 1, tid, ticket);
                sf[s] = rf.x; ticket = rf.y;
                sit[s] = 0; sret[s] = 0; sset[s] = false;
                live += sf[s] >= 0;
            }
            while (live > 0) {
                tmem_wait_st();
                passA<0>(softn, trow, mbar, ph, lane0, active, k);                               // STATE 1
                if (ALL_ACTIVE || active) passB<0>(softn, gy, tg, k);                            // STATE 2
                __syncthreads();
                unsigned bad = 0;                                                                // STATE 3
                passC<0>(softn, trow, k, bad);
                if (!(ALL_ACTIVE || active)) bad = 0;
                const bool b0 = __any_sync(0xffffffffu, (bad >> 15) & 1u), b1 = __any_sync(0xffffffffu, bad >> 31);
                if (lane0) { if (b0) s_misc[4 + 2 * g] = 1; if (b1) s_misc[4 + 2 * g + 1] = 1; }
                __syncthreads();
                int par[S];
#pragma unroll
                for (int s = 0; s < S; s++) par[s] = s_misc[4 + s];
                __syncthreads();
                if (tid < S) s_misc[4 + tid] = 0;
                bool fin[S];
#pragma unroll
                for (int s = 0; s < S; s++) {
                    fin[s] = false;
                    if (sf[s] < 0) continue;
                    sit[s]++;
                    if (!par[s] && !sset[s]) { sret[s] = sit[s]; sset[s] = true; }               // :5680-5685
                    if ((!par[s] && !noexit) || sit[s] >= io.maxiter) {
                        emit(io, soft2 + (s >> 1) * GROUP_WORDS, s_misc, sf[s], s & 1, sset[s] ? sret[s] : -sit[s], tid);   // :5689
                        fin[s] = true;
                    }
                }
#pragma unroll
                for (int s = 0; s < S; s++)
                    if (fin[s]) {
                        const int2 rf = refill(io, sp, soft2, s_misc, trow, s >> 1, s & 1, tid, ticket);
                        sf[s] = rf.x; ticket = rf.y;
                        sit[s] = 0; sret[s] = 0; sset[s] = false;
                        live -= sf[s] < 0;
                    }
            }
        }

        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid < 32) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase), "r"((unsigned)TCOLS) : "memory");
        }
    }
};

} // namespace ldpcb200
