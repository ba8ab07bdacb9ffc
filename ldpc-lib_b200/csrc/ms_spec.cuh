// Code-specialised kernels for the flooding min-sum pair -- MS_DEC (min_sum_decod_qc_lm, decoders.cpp:4554-4767,
// here in fp32) and IMS_DEC (imin_sum_decod_qc_lm, decoders.cpp:5430-5690, fixed point, bit-exact) -- built on the
// same machinery as lms_spec.cuh: one frame per CTA, lane n owns check row (j, n) of every block row j, the check
// state (min1, min2, c2v signs + minimum flags) of all block rows stays in REGISTERS, the base matrix is baked in.
//
// A flooding iteration needs, per bit, the sum of the check-to-variable messages of ALL its checks, accumulated by
// the reference in ascending block-row order (the order matters: float rounding in MS_DEC, saturation after every
// add in IMS_DEC, decoders.cpp:4658 / 5565-5569).  It is done here from the check side:
//   pass A  for j = 0 .. b-1 (barrier between block rows): lane n adds the message of each edge of row (j, n) into
//           the accumulator of its bit -- lanes of one block row touch disjoint bits, and a bit gets at most one
//           message per block row, so every bit sees its messages in ascending block-row order, as in the reference;
//   pass B  per bit: MS: soft = y + acc * alpha (:4682)   IMS: soft = sat(iy + acc) (:5599-5601), written to both
//           copies of the doubled column;
//   pass C  per check row, no barriers (reads soft, writes registers): v2c = soft - scaled old message, signs, the two
//           smallest |v2c|, syndrome of the hard decisions of this pass (:4688-4755 / :5608-5678).
// Shared memory per frame: the doubled posteriors (copy 0 doubles as pass A's accumulator) + the channel values,
// 12N bytes.  This header must stay free of #include (NVRTC), and follows lms_spec.cuh in the translation unit.
#pragma once

namespace ldpcb200 {

template <class K, bool IS_INT>
struct MsSpec {
    typedef LmsSpec<K> Base;
    typedef typename Base::RowAcc RowAcc;
    static constexpr int B = K::B, C = K::C, Z = K::Z, N = K::C * K::Z, R = K::B * K::Z, ZP = K::ZP;
    static constexpr int NWORDS = (N + 31) / 32;
    static constexpr bool ALL_ACTIVE = (Z == ZP);
    static constexpr int CS = 2 * Z;

    static __device__ __forceinline__ int sat(int x, int mx) { return x > mx ? mx : (x < -mx ? -mx : x); }   // limit_val :4308

    // accumulator of the bit an edge points at (copy 0 of the doubled column; wrapped lanes use the second address)
    template <int OFF, int SH>
    static __device__ __forceinline__ float acc_load(unsigned saddr, int n)
    {
        if constexpr (SH == 0) return Base::template load_plain<OFF>(saddr);
        else return Base::template load_wrapped<OFF, Z - SH>(saddr, n);
    }
    template <int OFF, int SH>
    static __device__ __forceinline__ void acc_store(unsigned saddr, int n, float x)
    {
        if constexpr (SH == 0) Base::template store_plain<OFF>(saddr, x);
        else Base::template store_wrapped<OFF, Z - SH>(saddr, n, x);
    }

    // ---- pass A: the (unscaled) old message of every edge of row J into the accumulators, copy 0 of each column.
    // Three sweeps over the row's edges -- all loads, then the adds, then all stores -- so that the (ordered, inline
    // assembly) shared-memory loads are in flight together; the edges of one row hit distinct block columns.
    template <int J, int Q>
    static __device__ __forceinline__ void scatter_load(unsigned saddr, int n, float (&acc)[K::RP[J + 1] - K::RP[J]])
    {
        constexpr int E0 = K::RP[J], DEG = K::RP[J + 1] - K::RP[J];
        if constexpr (Q < DEG) {
            constexpr int sh = K::SH[E0 + Q];
            constexpr int off = K::COL[E0 + Q] * CS + sh;
            if constexpr (K::FIRST[E0 + Q]) acc[Q] = 0.0f;                                   // :4633 / :5536 (int 0 has the same bits)
            else acc[Q] = acc_load<off, sh>(saddr, n);
            scatter_load<J, Q + 1>(saddr, n, acc);
        }
    }

    template <int J, int Q>
    static __device__ __forceinline__ void scatter_add(float m1, float m2, unsigned ps, const MsSpecParams& sp, float (&acc)[K::RP[J + 1] - K::RP[J]])
    {
        constexpr int DEG = K::RP[J + 1] - K::RP[J];
        if constexpr (Q < DEG) {
            const bool flag = ps & (0x10000u << Q);
            const bool neg = ps & (1u << Q);
            if constexpr (IS_INT) {
                const int t = __float_as_int(flag ? m2 : m1);                                // :5551, already scaled (:5554) by the caller
                const int val = neg ? -t : t;
                acc[Q] = __int_as_float(sat(__float_as_int(acc[Q]) + val, sp.max_data));     // :5567-5568
            } else {
                const float tmp = flag ? m2 : m1;                                            // :4649
                acc[Q] = __fadd_rn(acc[Q], neg ? -tmp : tmp);                                // :4658
            }
            scatter_add<J, Q + 1>(m1, m2, ps, sp, acc);
        }
    }

    template <int J, int Q>
    static __device__ __forceinline__ void scatter_store(unsigned saddr, int n, const float (&acc)[K::RP[J + 1] - K::RP[J]])
    {
        constexpr int E0 = K::RP[J], DEG = K::RP[J + 1] - K::RP[J];
        if constexpr (Q < DEG) {
            constexpr int sh = K::SH[E0 + Q];
            constexpr int off = K::COL[E0 + Q] * CS + sh;
            acc_store<off, sh>(saddr, n, acc[Q]);
            scatter_store<J, Q + 1>(saddr, n, acc);
        }
    }

    template <int J, int Q>
    static __device__ __forceinline__ void scatter(unsigned saddr, int n, float m1, float m2, unsigned ps, const MsSpecParams& sp)
    {
        float acc[K::RP[J + 1] - K::RP[J]];
        asm volatile("" : "+r"(ps), "+f"(m1), "+f"(m2) :: "memory");           // scheduling fence, see passC
        if constexpr (IS_INT) {                                                 // (min * ialpha) >> 4 once per row, not per edge (:5554)
            m1 = __int_as_float((__float_as_int(m1) * sp.ialpha) >> 4);
            m2 = __int_as_float((__float_as_int(m2) * sp.ialpha) >> 4);
        }
        scatter_load<J, 0>(saddr, n, acc);
        scatter_add<J, 0>(m1, m2, ps, sp, acc);
        scatter_store<J, 0>(saddr, n, acc);
    }

    template <int J>
    static __device__ __forceinline__ void passA(unsigned saddr, int n, bool active, const float (&m1)[B], const float (&m2)[B],
                                                 const unsigned (&ps)[B], const MsSpecParams& sp)
    {
        if constexpr (J < B) {
            if (ALL_ACTIVE || active) scatter<J, 0>(saddr, n, m1[J], m2[J], ps[J], sp);
            __syncthreads();
            passA<J + 1>(saddr, n, active, m1, m2, ps, sp);
        }
    }

    // ---- pass C: one check row
    template <int J, int Q>
    static __device__ __forceinline__ void gather(const float* soft2, int n, float pm1, float pm2, unsigned pps, const MsSpecParams& sp,
                                                  float (&v)[K::RP[J + 1] - K::RP[J]], unsigned& synd)
    {
        constexpr int E0 = K::RP[J], DEG = K::RP[J + 1] - K::RP[J];
        if constexpr (Q < DEG) {
            constexpr int off = K::COL[E0 + Q] * CS + K::SH[E0 + Q];
            const float rs = soft2[n + off];
            const bool flag = pps & (0x10000u << Q);
            const bool neg = pps & (1u << Q);
            if constexpr (IS_INT) {
                const int r = __float_as_int(rs);
                synd ^= (unsigned)r;                                                         // :5631 (sign bit = rs < 0)
                const int val = __float_as_int(flag ? pm2 : pm1);                            // scaled by the caller (:5640)
                const int tt = neg ? -val : val;
                v[Q] = __int_as_float(r - tt);                                               // :5646
            } else {
                synd ^= __float_as_uint(rs);                                                 // :4711
                const float val = flag ? pm2 : pm1;                                          // :4714, scaled by the caller (:4719)
                const float tt = neg ? -val : val;
                v[Q] = __fsub_rn(rs, tt);                                                    // :4722
            }
            gather<J, Q + 1>(soft2, n, pm1, pm2, pps, sp, v, synd);
        }
    }

    // integer twin of LmsSpec::reduce: two smallest |v2c| and the XOR of the signs
    template <int J, int Q>
    static __device__ __forceinline__ void ireduce(const float (&v)[K::RP[J + 1] - K::RP[J]], int& c1, int& c2, unsigned& sacc)
    {
        constexpr int DEG = K::RP[J + 1] - K::RP[J];
        if constexpr (Q < DEG) {
            const int x = __float_as_int(v[Q]);
            sacc ^= (unsigned)x;
            const int a = x < 0 ? -x : x;
            c2 = min(c2, max(c1, a));
            c1 = min(c1, a);
            ireduce<J, Q + 1>(v, c1, c2, sacc);
        }
    }

    template <int J, int Q>
    static __device__ __forceinline__ void flags(const float (&v)[K::RP[J + 1] - K::RP[J]], float c1f, int c1i, unsigned rs, unsigned& S, unsigned& MF)
    {
        if constexpr (Q >= 0) {
            bool ismin;
            unsigned bits = __float_as_uint(v[Q]);
            if constexpr (IS_INT) { const int x = (int)bits; ismin = (x < 0 ? -x : x) == c1i; }
            else ismin = fabsf(v[Q]) == c1f;
            if (ismin) MF |= 0x10000u << Q;
            S = __funnelshift_l(bits ^ rs, S, 1);                                            // c2v sign = sign(v2c) ^ row sign
            flags<J, Q - 1>(v, c1f, c1i, rs, S, MF);
        }
    }

    template <int J>
    static __device__ __forceinline__ void passC(const float* soft2, int n, float (&m1)[B], float (&m2)[B], unsigned (&ps)[B],
                                                 const MsSpecParams& sp, unsigned& bad)
    {
        if constexpr (J < B) {
            constexpr int DEG = K::RP[J + 1] - K::RP[J];
            float v[DEG];
            unsigned synd = 0, sacc = 0, S = 0, MF = 0;
            // scheduling fence: the old messages of ALL block rows depend only on registers, and without this the
            // compiler computes all E of them up front (hundreds of live registers, kilobytes of spills)
            asm volatile("" : "+r"(ps[J]), "+f"(m1[J]), "+f"(m2[J]) :: "memory");
            float a1, a2;                                                                    // old minima, scaled once per row
            if constexpr (IS_INT) {
                a1 = __int_as_float((__float_as_int(m1[J]) * sp.ialpha) >> 4);
                a2 = __int_as_float((__float_as_int(m2[J]) * sp.ialpha) >> 4);
            } else {
                a1 = __fmul_rn(m1[J], sp.alpha);
                a2 = __fmul_rn(m2[J], sp.alpha);
            }
            gather<J, 0>(soft2, n, a1, a2, ps[J], sp, v, synd);
            bad |= synd;
            if constexpr (IS_INT) {
                int c1 = 0x7fffffff, c2 = 0x7fffffff;
                ireduce<J, 0>(v, c1, c2, sacc);
                flags<J, DEG - 1>(v, 0.0f, c1, sacc & 0x80000000u, S, MF);
                m1[J] = __int_as_float(min(c1, sp.max_data));                                // :5653, init :5612-5616
                m2[J] = __int_as_float(min(c2, sp.max_data));
            } else {
                RowAcc a;
                a.c1 = __int_as_float(0x7f800000); a.c2 = a.c1;
                Base::template reduce<J, 0>(v, a, sacc);
                flags<J, DEG - 1>(v, a.c1, 0, sacc & 0x80000000u, S, MF);
                m1[J] = fminf(a.c1, 32767.0f);                                               // :4730, init :4692-4696
                m2[J] = fminf(a.c2, 32767.0f);
            }
            ps[J] = S | MF;                                                                  // :4753 / :5675
            passC<J + 1>(soft2, n, m1, m2, ps, sp, bad);
        }
    }

    static __device__ __forceinline__ void kernel(const FrameIO& io, const MsSpecParams& sp)
    {
        extern __shared__ __align__(16) float soft2[];
        float* y = soft2 + C * CS;                               // channel values: fp32 LLRs (MS) or quantised ints (IMS)
        int* s_misc = (int*)(y + N);
        double* s_sq = (double*)(soft2 + ((C * CS + N + 4 + 1) & ~1));   // IMS: staging of y^2 for the sequential energy sum (8-byte aligned)
        const int tid = threadIdx.x;
        const bool active = tid < Z;
        const bool noexit = io.flags & 8u;                       // LDPCB200_NO_EARLY_EXIT
        const unsigned saddr = (unsigned)__cvta_generic_to_shared(soft2 + tid);
        float m1[B], m2[B];
        unsigned ps[B];

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
            __syncthreads();
            if constexpr (IS_INT) {
                // per-frame energy normalisation + quantiser (:5472-5500).  The reference works on doubles: a double
                // input buffer is re-read from memory so that nothing is rounded through fp32; generated / fp32 LLRs are
                // exactly representable.  en += y*y runs in the reference's order on one thread (products in parallel).
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
                    y[i] = __int_as_float(q);
                    if (io.aux) io.aux[(size_t)f * N + i] = (short)q;
                }
                __syncthreads();
            }
#pragma unroll
            for (int j = 0; j < B; j++) { m1[j] = IS_INT ? __int_as_float(0) : 0.0f; m2[j] = m1[j]; ps[j] = 0u; }   // :4579-4596 / :5463-5502

            int parity = 1, ret = 0, locked = 0, iter;
            for (iter = 0; iter < io.maxiter; iter++) {
                passA<0>(saddr, tid, active, m1, m2, ps, sp);                                // STATE 1
                if (ALL_ACTIVE || active) {                                                  // STATE 2
#pragma unroll 8
                    for (int col = 0; col < C; col++) {
                        float s;
                        if constexpr (IS_INT) s = __int_as_float(sat(__float_as_int(y[col * Z + tid]) + __float_as_int(soft2[col * CS + tid]), sp.max_data));
                        else s = __fadd_rn(y[col * Z + tid], __fmul_rn(soft2[col * CS + tid], sp.alpha));
                        soft2[col * CS + tid] = s; soft2[col * CS + Z + tid] = s;
                    }
                }
                __syncthreads();
                unsigned bad = 0;                                                            // STATE 3
                if (ALL_ACTIVE || active) passC<0>(soft2, tid, m1, m2, ps, sp, bad);
                parity = __syncthreads_or((int)(bad >> 31));
                if (!parity && !locked) { ret = iter + 1; locked = 1; }
                if (!parity && !noexit) break;                                               // :4761 / :5680-5685
            }
            if (!locked) ret = parity ? -iter : iter + 1;                                    // :4766 / :5689
            if (io.maxiter <= 0) {                                                           // no pass ran: decisions of the channel values
                for (int col = 0; col < C; col++)
                    if (ALL_ACTIVE || active) soft2[col * CS + tid] = y[col * Z + tid];
                __syncthreads();
            }

            if (io.post) {
                for (int col = 0; col < C; col++)
                    if (ALL_ACTIVE || active) {
                        const size_t k = (size_t)f * N + col * Z + tid;
                        const float s = soft2[col * CS + tid];
                        if constexpr (IS_INT) ((short*)io.post)[k] = (short)__float_as_int(s);
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
                    if (i < N) { const int col = i / Z, k = i - col * Z; bit = (int)(__float_as_uint(soft2[col * CS + k]) >> 31) & (IS_INT ? 1 : (soft2[col * CS + k] < 0.0f)); }
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
    }
};

} // namespace ldpcb200
