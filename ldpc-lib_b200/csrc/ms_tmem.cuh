// Flooding min-sum pair with the check-to-variable messages in TENSOR MEMORY: MS_DEC (min_sum_decod_qc_lm,
// decoders.cpp:4554-4767, fp32) and IMS_DEC (imin_sum_decod_qc_lm, decoders.cpp:5430-5690, fixed point, bit-exact).
// Same arithmetic and summation order as ms_spec.cuh (see there for why the bit sums are accumulated from the check
// side, block row by block row); what changes is where the state lives, as in lms_tmem.cuh:
//
//   * the signed message of every edge and lane -- MS: +-min (unscaled, what pass A adds, :4649-4658); IMS: the scaled
//     +-((min * ialpha) >> 4) (:5551-5563) -- is one 32-bit word in TMEM (column RP[J] + q of the thread's lane).
//     Pass A and pass C fetch a block row's messages with one tcgen05.ld group; pass C puts the new ones back with one
//     tcgen05.st.  No per-row {min1, min2, sign word, flags} in registers, no select / shift / sign fix-up per edge;
//   * accumulators / posteriors in shared memory, doubled columns kept in the rotation of their last writer
//     (K::DELTA / ROT / SYNSH, lms_tmem.cuh): pass A reads position n + DELTA and writes positions n and n + Z with plain
//     instructions, split mbarrier between a block row's loads and stores, __syncthreads between block rows;
//   * pass B per position: MS soft = y + acc * alpha (:4682), IMS soft = sat(iy + acc) (:5599-5601); the channel values y
//     stay in bit order (the IMS quantiser sums their squares sequentially), the column rotation is applied here;
//   * pass C per check row, no barriers: v2c = soft - alpha * old (MS) / soft - old (IMS), syndrome of this pass's
//     decisions, two smallest |v2c|, new messages.
// IMS_DEC's fixed-point values are small integers (|x| <= 2^dbits, products with ialpha < 2^24), which fp32 represents
// and adds / subtracts / compares / scales by ialpha/16 EXACTLY; the kernel therefore carries them as integer-valued
// floats and shares MS_DEC's instruction stream (FADD, FMNMX for the saturation, sign-bit logic) -- the results are the
// reference's integers bit for bit (the host checks max_data * ialpha < 2^24 before choosing this kernel).
// This header must stay free of #include (NVRTC), and follows lms_tmem.cuh in the translation unit.
#pragma once

namespace ldpcb200 {

template <class K, bool IS_INT>
struct MsTmem {
    typedef LmsSpec<K> S;
    typedef LmsTmem<K> T;
    static constexpr int B = K::B, C = K::C, Z = K::Z, N = K::C * K::Z, R = K::B * K::Z, ZP = K::ZP, E = K::E;
    static constexpr int NWORDS = (N + 31) / 32;
    static constexpr bool ALL_ACTIVE = (Z == ZP);
    static constexpr int CS = 2 * Z;
    static constexpr int NWARPS = ZP / 32;
    static constexpr int TCOLS = K::TCOLS;
    // shared memory (words): doubled accumulators / posteriors | channel values | mbarrier (8-byte aligned) | misc |
    // staging of y^2 for the IMS energy sum (doubles)
    static constexpr int Y_OFF = C * CS;
    static constexpr int MBAR_OFF = (Y_OFF + N + 1) & ~1;
    static constexpr int MISC_OFF = MBAR_OFF + 2;
    static constexpr int SQ_OFF = (MISC_OFF + 4 + 1) & ~1;
    static constexpr int SMEM_WORDS = SQ_OFF + (IS_INT ? 1024 : 0);

    // ---- pass A, block row J: acc[bit] += message, in ascending block-row order per bit
    template <int J, int Q>
    static __device__ __forceinline__ void accA_load(const float* softn, float (&acc)[K::RP[J + 1] - K::RP[J]])
    {
        constexpr int E0 = K::RP[J], DEG = K::RP[J + 1] - K::RP[J];
        if constexpr (Q < DEG) {
            constexpr int off = K::COL[E0 + Q] * CS + K::DELTA[E0 + Q];
            if constexpr (K::FIRST[E0 + Q]) acc[Q] = 0.0f;                       // :4633 / :5536 (int 0 has the same bits)
            else acc[Q] = softn[off];
            accA_load<J, Q + 1>(softn, acc);
        }
    }
    template <int J, int Q>
    static __device__ __forceinline__ void accA_store(float* softn, bool active, const float (&acc)[K::RP[J + 1] - K::RP[J]])
    {
        constexpr int E0 = K::RP[J], DEG = K::RP[J + 1] - K::RP[J];
        if constexpr (Q < DEG) {
            constexpr int off = K::COL[E0 + Q] * CS;
            if (ALL_ACTIVE || active) { softn[off] = acc[Q]; softn[off + Z] = acc[Q]; }
            accA_store<J, Q + 1>(softn, active, acc);
        }
    }
    template <int J>
    static __device__ __forceinline__ void passA(float* softn, unsigned trow, unsigned mbar, unsigned& ph, bool lane0, bool active,
                                                 const MsSpecParams& sp)
    {
        if constexpr (J < B) {
            constexpr int E0 = K::RP[J], DEG = K::RP[J + 1] - K::RP[J];
            unsigned msg[DEG];
            float acc[DEG];
            const float mx = (float)sp.max_data;
            tmem_ld_n<DEG>(trow + E0, msg);
            accA_load<J, 0>(softn, acc);
            tmem_wait_ld<DEG>(msg);
#pragma unroll
            for (int q = 0; q < DEG; q++) {
                acc[q] = __fadd_rn(acc[q], __uint_as_float(msg[q]));                             // :4658 / :5567
                if constexpr (IS_INT) acc[q] = fminf(fmaxf(acc[q], -mx), mx);                    // :5568 saturate after every add
            }
            T::loads_done(mbar, lane0);
            if constexpr (B % 2 == 0) T::wait_loads(mbar, J & 1);
            else { T::wait_loads(mbar, ph); ph ^= 1u; }
            accA_store<J, 0>(softn, active, acc);
            __syncthreads();
            passA<J + 1>(softn, trow, mbar, ph, lane0, active, sp);
        }
    }

    // ---- pass B, position tid of every column: the column is rotated by K::ROT, the channel values are not
    template <int COL>
    static __device__ __forceinline__ void passB(float* softn, const float* y, int tid, const MsSpecParams& sp)
    {
        if constexpr (COL < C) {
            constexpr int rot = K::ROT[COL];
            int k = tid + rot;
            if (rot != 0 && k >= Z) k -= Z;
            const float yv = y[COL * Z + k];
            const float a = softn[COL * CS];
            float s;
            if constexpr (IS_INT) s = fminf(fmaxf(__fadd_rn(yv, a), -(float)sp.max_data), (float)sp.max_data);   // :5599-5601
            else s = __fadd_rn(yv, __fmul_rn(a, sp.alpha));                                      // :4682
            softn[COL * CS] = s; softn[COL * CS + Z] = s;
            passB<COL + 1>(softn, y, tid, sp);
        }
    }

    // ---- pass C, block row J
    template <int J, int Q>
    static __device__ __forceinline__ void softC_load(const float* softn, float (&rs)[K::RP[J + 1] - K::RP[J]])
    {
        constexpr int E0 = K::RP[J], DEG = K::RP[J + 1] - K::RP[J];
        if constexpr (Q < DEG) {
            constexpr int off = K::COL[E0 + Q] * CS + K::SYNSH[E0 + Q];
            rs[Q] = softn[off];
            softC_load<J, Q + 1>(softn, rs);
        }
    }
    template <int J>
    static __device__ __forceinline__ void passC(const float* softn, unsigned trow, const MsSpecParams& sp, unsigned& bad)
    {
        if constexpr (J < B) {
            constexpr int E0 = K::RP[J], DEG = K::RP[J + 1] - K::RP[J];
            unsigned msg[DEG];
            float rs[DEG], v[DEG];
            tmem_ld_n<DEG>(trow + E0, msg);
            softC_load<J, 0>(softn, rs);
            tmem_wait_ld<DEG>(msg);
            unsigned synd = 0;
#pragma unroll
            for (int q = 0; q < DEG; q++) {
                synd ^= __float_as_uint(rs[q]);                                                  // :4711 / :5631 (sign bit = rs < 0)
                if constexpr (IS_INT) v[q] = __fsub_rn(rs[q], __uint_as_float(msg[q]));          // :5646, the message is already scaled (:5640)
                else v[q] = __fsub_rn(rs[q], __fmul_rn(__uint_as_float(msg[q]), sp.alpha));      // :4714-4722
            }
            // the message of an edge = g(min over the OTHER edges of |v2c|) with the row's sign product and the edge's
            // own sign (lms_tmem.cuh min_of_others: one FMNMX3 per edge instead of two-smallest tracking + a select)
            float m[DEG];
            const float rone = __uint_as_float((T::template sign_xor<DEG, 0, DEG>(v) & 0x80000000u) | 0x3f800000u);
            if constexpr (IS_INT) {
                // g(c) = (min(c, max_data) * ialpha) >> 4 (:5653, :5554, :5656-5666) on integers carried as floats:
                // t = fma(c, ialpha / 16, 1.5 * 2^23) rounded DOWN is 1.5 * 2^23 + floor(c * ialpha / 16) exactly (ulp 1 up
                // there); the second fma takes the constant off again and applies the row sign: +-floor(x), exact.
                const float MAGIC = 12582912.0f;
                const float scale = (float)sp.ialpha * 0.0625f, nmag = __fmul_rn(rone, -MAGIC);
                T::template min_of_others<DEG>(v, m, (float)sp.max_data);
#pragma unroll
                for (int q = 0; q < DEG; q++) {
                    const float t = __fmaf_rd(m[q], scale, MAGIC);
                    msg[q] = __float_as_uint(__fmaf_rn(t, rone, nmag)) ^ (__float_as_uint(v[q]) & 0x80000000u);
                }
            } else {
                T::template min_of_others<DEG>(v, m, 32767.0f);                                  // :4730-4746, init :4692-4696
#pragma unroll
                for (int q = 0; q < DEG; q++)
                    msg[q] = __float_as_uint(__fmul_rn(m[q], rone)) ^ (__float_as_uint(v[q]) & 0x80000000u);
            }
            bad |= synd;
            tmem_st_n<DEG>(trow + E0, msg);                                                      // :4753 / :5675
            passC<J + 1>(softn, trow, sp, bad);
        }
    }

    static __device__ __forceinline__ void kernel(const FrameIO& io, const MsSpecParams& sp)
    {
        extern __shared__ __align__(16) float soft2[];
        float* y = soft2 + Y_OFF;                                // channel values in bit order: fp32 LLRs (MS) or quantised ints (IMS)
        int* s_misc = (int*)(soft2 + MISC_OFF);
        double* s_sq = (double*)(soft2 + SQ_OFF);                // IMS: staging of y^2 for the sequential energy sum
        const int tid = threadIdx.x;
        const bool active = tid < Z;
        const bool lane0 = (tid & 31) == 0;
        const bool noexit = io.flags & 8u;                       // LDPCB200_NO_EARLY_EXIT
        float* softn = soft2 + tid;
        const unsigned mbar = (unsigned)__cvta_generic_to_shared(soft2 + MBAR_OFF);
        unsigned ph = 0;
        if (tid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar), "r"((unsigned)NWARPS) : "memory");
        if (tid < 32) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                         :: "r"((unsigned)__cvta_generic_to_shared(s_misc)), "r"((unsigned)TCOLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const unsigned tbase = *(volatile unsigned*)s_misc;
        const unsigned trow = __shfl_sync(0xffffffffu, tbase + ((unsigned)(((tid >> 5) & 3) * 32) << 16) + (unsigned)((tid >> 7) * E), 0);

        for (;;) {
            __syncthreads();
            if (tid == 0) { s_misc[0] = (int)atomicAdd(io.next_frame, 1u); s_misc[1] = 0; s_misc[2] = 0; }
            __syncthreads();
            const int f = s_misc[0];
            if (f >= io.nf) break;

            // ---- first load: channel LLRs (fp32) into y
            if (io.ch.enabled) {
                const unsigned long long frame = io.ch.first_frame + (unsigned long long)f;
                if (io.ch.m > 2) {
                    const int half = io.ch.m >> 1, ncomp = 2 * (N / io.ch.m);
                    for (int c = tid; c < ncomp; c += ZP) {
                        float o[4];
                        channel_llr_qam_component(io.ch, frame, c, o);
                        const int i0 = (c >> 1) * io.ch.m + (c & 1) * half;
                        for (int b = 0; b < half; b++) {
                            const int i = channel_dest(io.ch, i0 + b);
                            y[i] = i >= io.ch.punct_start ? io.ch.punct_value : o[b];
                        }
                    }
                } else {
                    for (int i4 = tid; i4 < N / 4; i4 += ZP) {                        // one Philox block -> four LLRs
                        float o[4];
                        int d[4];
                        channel_llr4_bpsk(io.ch, frame, i4, o, d);
#pragma unroll
                        for (int b = 0; b < 4; b++) y[d[b]] = o[b];
                    }
                    for (int j = (N & ~3) + tid; j < N; j += ZP) { const int i = channel_dest(io.ch, j); y[i] = channel_llr(io.ch, frame, i); }
                }
            } else if (io.llr_dtype == 1) {
                const float* src = (const float*)io.llr + (size_t)f * N;
                for (int i = tid; i < N; i += ZP) y[i] = __ldcs(src + i);
            } else {
                const double* src = (const double*)io.llr + (size_t)f * N;
                for (int i = tid; i < N; i += ZP) y[i] = (float)__ldcs(src + i);
            }
            tmem_zero_n<E>(trow);                                                    // dcs[] = 0, :4579-4596 / :5463-5502
            tmem_wait_st();
            __syncthreads();
            if constexpr (IS_INT) {
                // per-frame energy normalisation + quantiser (:5472-5500), as in ms_spec.cuh
                const double* src64 = (!io.ch.enabled && io.llr_dtype == 0) ? (const double*)io.llr + (size_t)f * N : nullptr;
                if (io.coef) {
                    if (tid == 0) s_sq[0] = io.coef[f];                                      // from the energy pre-pass (channel.cu)
                } else {
                    double en = 0;
                    for (int base = 0; base < N; base += 512) {
                        __syncthreads();
                        for (int i = tid; i < 512 && base + i < N; i += ZP) {
                            const double val = src64 ? src64[base + i] : (double)y[base + i];
                            s_sq[i] = val * val;
                        }
                        __syncthreads();
                        if (tid == 0) {
                            const int m = N - base < 512 ? N - base : 512;
                            for (int i = 0; i < m; i++) en += s_sq[i];
                        }
                    }
                    if (tid == 0) s_sq[0] = sqrt(N / en);                                    // :5479
                }
                __syncthreads();
                const double coef = s_sq[0];
                __syncthreads();
                for (int i = tid; i < N; i += ZP) {
                    double val = src64 ? src64[i] : (double)y[i];
                    int sign = 0;
                    if (val < 0) { val = -val; sign = 1; }
                    val *= coef;
                    if (val > sp.thr) val = sp.thr;
                    const int ival = (short)floor(val * sp.max_quant / sp.thr + 0.5);
                    const int q = sign ? -ival : ival;
                    y[i] = (float)q;
                    if (io.aux) io.aux[(size_t)f * N + i] = (short)q;
                }
                __syncthreads();
            }

            int parity = 1, ret = 0, locked = 0, iter;
            for (iter = 0; iter < io.maxiter; iter++) {
                tmem_wait_st();
                passA<0>(softn, trow, mbar, ph, lane0, active, sp);                              // STATE 1
                if (ALL_ACTIVE || active) passB<0>(softn, y, tid, sp);                       // STATE 2
                __syncthreads();
                unsigned bad = 0;                                                            // STATE 3
                passC<0>(softn, trow, sp, bad);
                parity = __syncthreads_or((ALL_ACTIVE || active) ? (int)(bad >> 31) : 0);
                if (!parity && !locked) { ret = iter + 1; locked = 1; }
                if (!parity && !noexit) break;                                               // :4761 / :5680-5685
            }
            if (!locked) ret = parity ? -iter : iter + 1;                                    // :4766 / :5689
            if (io.maxiter <= 0) {                                                           // no pass ran: decisions of the channel values
                for (int col = 0; col < C; col++)
                    if (ALL_ACTIVE || active) {
                        int k = tid + K::rt_rot()[col];
                        if (k >= Z) k -= Z;
                        softn[col * CS] = y[col * Z + k]; softn[col * CS + Z] = y[col * Z + k];
                    }
                __syncthreads();
            }

            if (io.post) {
                for (int col = 0; col < C; col++)
                    if (ALL_ACTIVE || active) {
                        const size_t k = (size_t)f * N + col * Z + tid;
                        const float s = soft2[col * CS + tid + K::rt_ri()[col]];
                        if constexpr (IS_INT) ((short*)io.post)[k] = (short)(int)s;
                        else if (io.post_dtype == 1) ((float*)io.post)[k] = s;
                        else ((double*)io.post)[k] = (double)s;
                    }
            }
            {
                const int lane = tid & 31;
                int nerr = 0, nerr_info = 0;
                constexpr int NROUND = (N + 31) & ~31;
                for (int i = tid; i < NROUND; i += ZP) {
                    int bit = 0;
                    if (i < N) {
                        const int col = i / Z, k = i - col * Z;
                        const float s = soft2[col * CS + k + K::rt_ri()[col]];
                        bit = s < 0.0f;
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
