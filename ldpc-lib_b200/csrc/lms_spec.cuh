// Code-specialised LMS_DEC kernel (fp32): the same arithmetic as lms_fast.cu (bit-identical results), with the
// base matrix baked in at compile time.  It is compiled once per code -- ahead of time for the frozen benchmark
// matrices (lms_spec_aot.cu), at run time through NVRTC for any other code (spec_jit.cpp) -- from a generated
//
//   struct Code { static constexpr int B, C, Z, E, ZP, MINB, MAXDEG;         // ZP = threads, MINB = CTAs per SM
//                 static constexpr bool DOUBLED, PS_SMEM;                    // layout variant, see below
//                 static constexpr int RP[B + 1], COL[E], SH[E];             // compile-time copies
//                 static const int* rt_rp() / rt_col() / rt_sh(); };         // __constant__ copies for run-time indexing
//
// What specialisation buys over the table-driven kernel: every layer is unrolled, so (1) the soft-value address
// of an edge is `lane + immediate` (no table loads, no index arithmetic; the mirror store of the doubled column
// is two predicated stores with immediate offsets), (2) the check-row state (min1, min2, signs|pos) of all B
// block rows of a lane lives in REGISTERS across iterations -- shared memory holds only the doubled posteriors,
// 8N bytes per frame (64 KB at N = 8192 -> 3 frames per SM instead of 2), and (3) the position compare is
// against an immediate.  Large codes (the 46 x 68, Z = 384 "BG1-shaped" one: 104 KB of posteriors, 46 block
// rows) use the second layout variant: DOUBLED = false (every block column stored once; the wrapped lanes
// read / write through a predicated second instruction with its own immediate) and PS_SMEM = true (the sign /
// position word of each check row in shared memory, only min1 / min2 in registers).  This header must stay free of #include (NVRTC compiles it as one string together with
// frame_io.h and channel.cuh).
#pragma once

namespace ldpcb200 {

template <class K>
struct LmsSpec {
    static constexpr int B = K::B, C = K::C, Z = K::Z, N = K::C * K::Z, R = K::B * K::Z, ZP = K::ZP;
    static constexpr int HW = ZP / 32;                       // words per block column of the packed decisions
    static constexpr int NB = (Z + 31) / 32;
    static constexpr int NWORDS = (N + 31) / 32;
    static constexpr bool ALL_ACTIVE = (Z == ZP);
    static constexpr bool DOUBLED = K::DOUBLED, PS_SMEM = K::PS_SMEM;
    static constexpr bool USE_POS = K::MAXDEG > 16;          // row word: 16 signs + 16 minimum flags, or 24 signs + position
    static constexpr int CS = DOUBLED ? 2 * Z : Z;           // stride of a block column in shared memory
    static constexpr int SOFT_WORDS = C * CS;
    static constexpr int PS_WORDS = PS_SMEM ? R : 0;

    // ---- one block row J of one lane.  The edge loops are compile-time recursions so that every per-edge
    // constant (K::COL, K::SH) is used in a constant expression: the tables never exist in device memory.
    // Row word ps: bit q = sign of the c2v message of edge q, bit 16 + q = "edge q carried the minimum" (row weight <= 16).
    // With ties several edges carry the flag; their message min2 then equals min1, so the value is the reference's.
    struct RowAcc { float c1, c2; };

    // two smallest of {c1, c2, a, b}: order independent, exact (min / max only)
    static __device__ __forceinline__ void track2(RowAcc& r, float a, float b)
    {
        const float lo = fminf(a, b), hi = fmaxf(a, b);
        const float t = fmaxf(r.c1, lo);
        r.c2 = fminf(fminf(r.c2, hi), t);
        r.c1 = fminf(r.c1, lo);
    }
    static __device__ __forceinline__ void track1(RowAcc& r, float a)
    {
        r.c2 = fminf(r.c2, fmaxf(r.c1, a));                                                  // decoders.cpp:5012-5027
        r.c1 = fminf(r.c1, a);
    }

    template <int J, int Q>
    static __device__ __forceinline__ void phase1(const float* soft2, unsigned saddr, int n, float pm1, float pm2, unsigned pps,
                                                  float (&v)[K::RP[J + 1] - K::RP[J]])
    {
        constexpr int E0 = K::RP[J], DEG = K::RP[J + 1] - K::RP[J];
        if constexpr (Q < DEG) {
            constexpr int sh = K::SH[E0 + Q];
            constexpr int off = K::COL[E0 + Q] * CS + sh;
            float sv;
            if constexpr (DOUBLED || sh == 0) sv = soft2[n + off];
            else sv = load_wrapped<off, Z - sh>(saddr, n);
            bool flag;
            if constexpr (USE_POS) flag = (pps >> 24) == (unsigned)Q;
            else flag = pps & (0x10000u << Q);
            const float pabs = flag ? pm2 : pm1;                                             // :5152
            const float pval = __uint_as_float(__float_as_uint(pabs) ^ ((pps << (31 - Q)) & 0x80000000u));   // :5156
            v[Q] = sv - pval;                                                                // :5158
            phase1<J, Q + 1>(soft2, saddr, n, pm1, pm2, pps, v);
        }
    }

    // minima and sign parity of the row from the v2c values, two edges per step
    template <int J, int Q>
    static __device__ __forceinline__ void reduce(const float (&v)[K::RP[J + 1] - K::RP[J]], RowAcc& a, unsigned& sacc)
    {
        constexpr int DEG = K::RP[J + 1] - K::RP[J];
        if constexpr (Q + 1 < DEG) {
            sacc = sacc ^ __float_as_uint(v[Q]) ^ __float_as_uint(v[Q + 1]);
            track2(a, fabsf(v[Q]), fabsf(v[Q + 1]));
            reduce<J, Q + 2>(v, a, sacc);
        } else if constexpr (Q < DEG) {
            sacc ^= __float_as_uint(v[Q]);
            track1(a, fabsf(v[Q]));
        }
    }

    // single-copy layout: lane n reads / writes bit (n + shift) mod Z of the column; lanes n >= THR have wrapped
    template <int OFF, int THR>
    static __device__ __forceinline__ float load_wrapped(unsigned saddr, int n)
    {
        float x;
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "setp.ge.s32 p, %2, %5;\n\t"
                     "@p ld.shared.f32 %0, [%1+%4];\n\t"
                     "@!p ld.shared.f32 %0, [%1+%3];\n\t}"
                     : "=f"(x) : "r"(saddr), "r"(n), "n"(4 * OFF), "n"(4 * (OFF - Z)), "n"(THR));
        return x;
    }
    template <int OFF, int THR>
    static __device__ __forceinline__ void store_wrapped(unsigned saddr, int n, float nv)
    {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "setp.ge.s32 p, %1, %5;\n\t"
                     "@p st.shared.f32 [%0+%3], %4;\n\t"
                     "@!p st.shared.f32 [%0+%2], %4;\n\t}"
                     :: "r"(saddr), "r"(n), "n"(4 * OFF), "n"(4 * (OFF - Z)), "f"(nv), "n"(THR) : "memory");
    }
    template <int OFF>
    static __device__ __forceinline__ float load_plain(unsigned saddr)
    {
        float x;
        asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(x) : "r"(saddr), "n"(4 * OFF));
        return x;
    }
    template <int OFF>
    static __device__ __forceinline__ void store_plain(unsigned saddr, float nv)
    {
        asm volatile("st.shared.f32 [%0+%1], %2;" :: "r"(saddr), "n"(4 * OFF), "f"(nv) : "memory");
    }

    // store nv at soft2[n + OFF] and at its mirror in the doubled column: lanes that wrapped (n >= THR) mirror
    // downwards, the others upwards -- two predicated stores with immediate offsets, no address arithmetic
    template <int OFF, int THR>
    static __device__ __forceinline__ void store2(unsigned saddr, int n, float nv)
    {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "st.shared.f32 [%0+%2], %5;\n\t"
                     "setp.ge.s32 p, %1, %6;\n\t"
                     "@p st.shared.f32 [%0+%3], %5;\n\t"
                     "@!p st.shared.f32 [%0+%4], %5;\n\t}"
                     :: "r"(saddr), "r"(n), "n"(4 * OFF), "n"(4 * (OFF - Z)), "n"(4 * (OFF + Z)), "f"(nv), "n"(THR) : "memory");
    }
    template <int OFF>
    static __device__ __forceinline__ void store2_noshift(unsigned saddr, float nv)
    {
        asm volatile("st.shared.f32 [%0+%1], %3;\n\tst.shared.f32 [%0+%2], %3;"
                     :: "r"(saddr), "n"(4 * OFF), "n"(4 * (OFF + Z)), "f"(nv) : "memory");
    }

    template <int J, int Q>
    static __device__ __forceinline__ void phase2(unsigned saddr, int n, const float (&v)[K::RP[J + 1] - K::RP[J]], float c1,
                                                  unsigned m1x, unsigned m2x, unsigned& S, unsigned& MF)
    {
        constexpr int E0 = K::RP[J];
        if constexpr (Q >= 0) {
            constexpr int sh = K::SH[E0 + Q];
            constexpr int off = K::COL[E0 + Q] * CS + sh;
            const bool ismin = fabsf(v[Q]) == c1;
            if constexpr (USE_POS) MF = ismin ? ((unsigned)Q << 24) : MF;                    // reverse scan: the first minimum wins
            else { if (ismin) MF |= 0x10000u << Q; }
            const unsigned cv = (ismin ? m2x : m1x) ^ (__float_as_uint(v[Q]) & 0x80000000u); // :5193-5198
            S = __funnelshift_l(cv, S, 1);                                                   // S = S << 1 | sign(cv)
            const float nv = v[Q] + __uint_as_float(cv);                                     // :5199-5204
            if constexpr (DOUBLED) {
                if constexpr (sh == 0) store2_noshift<off>(saddr, nv);                       // a zero shift never wraps
                else store2<off, Z - sh>(saddr, n, nv);
            } else {
                if constexpr (sh == 0) store_plain<off>(saddr, nv);
                else store_wrapped<off, Z - sh>(saddr, n, nv);
            }
            phase2<J, Q - 1>(saddr, n, v, c1, m1x, m2x, S, MF);
        }
    }

    template <int J>
    static __device__ __forceinline__ void layer(float* soft2, unsigned saddr, int n, float& m1, float& m2, unsigned& ps)
    {
        constexpr int DEG = K::RP[J + 1] - K::RP[J];
        float v[DEG];
        phase1<J, 0>(soft2, saddr, n, m1, m2, ps, v);
        RowAcc a;
        a.c1 = __int_as_float(0x7f800000); a.c2 = a.c1;
        unsigned sacc = 0;
        reduce<J, 0>(v, a, sacc);
        const float n1 = fminf(fmaxf(a.c1 - 0.4f, 0.0f), 32767.0f);                          // :5166-5168, :5131-5137
        const float n2 = fminf(fmaxf(a.c2 - 0.4f, 0.0f), 32767.0f);
        const unsigned rs = sacc & 0x80000000u;
        unsigned S = 0, MF = 0;
        phase2<J, DEG - 1>(saddr, n, v, a.c1, __float_as_uint(n1) ^ rs, __float_as_uint(n2) ^ rs, S, MF);
        m1 = n1; m2 = n2; ps = S | MF;                                                       // :5179
    }

    // Block rows that share no block column touch disjoint posteriors, so the barrier between them orders nothing: block row
    // j + 1 may start while other warps still work on block row j (and on every earlier row since the last barrier), with
    // the same result bit for bit.  need_barrier(j): does block row j + 1 share a column with any row of the barrier-free
    // group that block row j closes?  (BG1-shaped C3: 11 of its 46 barriers go; the 16 x 32 and 12 x 24 matrices keep all.)
    static __host__ __device__ constexpr bool rows_share(int a, int b)
    {
        for (int e = K::RP[a]; e < K::RP[a + 1]; e++)
            for (int f = K::RP[b]; f < K::RP[b + 1]; f++)
                if (K::COL[e] == K::COL[f]) return true;
        return false;
    }
    static __host__ __device__ constexpr bool need_barrier(int j)
    {
        if (j + 1 >= B) return true;                             // the iteration ends: syndrome / next iteration
        int start = 0;                                           // first row of the group that row j belongs to
        for (int i = 0; i < j; i++) {
            bool nb = false;
            for (int r = start; r <= i; r++) nb = nb || rows_share(r, i + 1);
            if (nb) start = i + 1;
        }
        for (int r = start; r <= j; r++)
            if (rows_share(r, j + 1)) return true;
        return false;
    }

    template <int J>
    static __device__ __forceinline__ void layers(float* soft2, unsigned saddr, int n, bool active, float (&m1)[B], float (&m2)[B],
                                                  unsigned (&ps)[PS_SMEM ? 1 : B])
    {
        if constexpr (J < B) {
            if (ALL_ACTIVE || active) {
                if constexpr (PS_SMEM) {
                    unsigned* psw = (unsigned*)(soft2 + SOFT_WORDS) + J * Z + n;
                    unsigned w = *psw;
                    layer<J>(soft2, saddr, n, m1[J], m2[J], w);
                    *psw = w;
                } else
                    layer<J>(soft2, saddr, n, m1[J], m2[J], ps[J]);
            }
            constexpr bool BARRIER = need_barrier(J);
            if constexpr (BARRIER) __syncthreads();
            layers<J + 1>(soft2, saddr, n, active, m1, m2, ps);
        }
    }

    // syndrome of the hard decisions on packed bits (see lms_fast.cu)
    static __device__ __forceinline__ int syndrome(const float* soft2, unsigned* hb, int tid)
    {
        const int lane = tid & 31, warp = tid >> 5;
#pragma unroll 4
        for (int col = 0; col < C; col++) {
            const int bit = (ALL_ACTIVE || tid < Z) ? soft2[col * CS + tid] < 0.0f : 0;
            const unsigned w = __ballot_sync(0xffffffffu, bit);
            if (lane == 0) hb[col * HW + warp] = w;
        }
        __syncthreads();
        unsigned bad = 0;
        for (int t = tid; t < B * NB; t += ZP) {
            const int j = t / NB, w = t - j * NB;
            unsigned acc = 0;
            const int* rp = K::rt_rp();
            const int* rcol = K::rt_col();
            const int* rsh = K::rt_sh();
            for (int e = rp[j]; e < rp[j + 1]; e++) {
                const unsigned* hc = hb + rcol[e] * HW;
                int start = 32 * w + rsh[e];
                if (start >= Z) start -= Z;
                const int i0 = start >> 5, i1 = i0 + 1 < HW ? i0 + 1 : HW - 1;
                unsigned win = __funnelshift_r(hc[i0], hc[i1], start & 31);
                const int nvalid = Z - start;
                if (nvalid < 32) win = (win & ((1u << nvalid) - 1u)) | (hc[0] << nvalid);
                acc ^= win;
            }
            const int lanes = Z - 32 * w;
            if (lanes < 32) acc &= (1u << lanes) - 1u;
            bad |= acc;
        }
        return __syncthreads_or(bad != 0);
    }

    static __device__ __forceinline__ void kernel(const FrameIO& io)
    {
        extern __shared__ __align__(16) float soft2[];
        unsigned* hb = (unsigned*)(soft2 + SOFT_WORDS + PS_WORDS);
        int* s_misc = (int*)hb;                                  // aliases hb: only live between frames
        const int tid = threadIdx.x;
        const bool active = tid < Z;
        const bool noexit = io.flags & 8u;                       // LDPCB200_NO_EARLY_EXIT
        const unsigned saddr = (unsigned)__cvta_generic_to_shared(soft2 + tid);     // shared-window byte address of this lane
        float m1[B], m2[B];
        unsigned ps[PS_SMEM ? 1 : B];

        for (;;) {
            __syncthreads();
            if (tid == 0) { s_misc[0] = (int)atomicAdd(io.next_frame, 1u); s_misc[1] = 0; s_misc[2] = 0; }
            __syncthreads();
            const int f = s_misc[0];
            if (f >= io.nf) break;

            if (io.ch.enabled) {
                const unsigned long long frame = io.ch.first_frame + (unsigned long long)f;
                if (io.ch.m > 2) {
                    // QAM-16/64/256: one thread per PAM component, m/2 LLRs from one demodulation
                    const int half = io.ch.m >> 1, ncomp = 2 * (N / io.ch.m);
                    for (int c4 = tid; c4 < ncomp / 4; c4 += ZP) {                    // four components from one Philox block
                        float o4[16];
                        channel_llr_qam_component4(io.ch, frame, c4, o4);
                        for (int q = 0; q < 4; q++) {
                            const int c = 4 * c4 + q, i0 = (c >> 1) * io.ch.m + (c & 1) * half;
                            for (int b = 0; b < half; b++) {
                                const int i = channel_dest(io.ch, i0 + b), col = i / Z, k = i - col * Z;
                                const float x = i >= io.ch.punct_start ? io.ch.punct_value : o4[4 * q + b];
                                soft2[col * CS + k] = x;
                                if constexpr (DOUBLED) soft2[col * CS + Z + k] = x;
                            }
                        }
                    }
                    for (int c = (ncomp & ~3) + tid; c < ncomp; c += ZP) {
                        float o[4];
                        channel_llr_qam_component(io.ch, frame, c, o);
                        const int i0 = (c >> 1) * io.ch.m + (c & 1) * half;
                        for (int b = 0; b < half; b++) {
                            const int i = channel_dest(io.ch, i0 + b), col = i / Z, k = i - col * Z;
                            const float x = i >= io.ch.punct_start ? io.ch.punct_value : o[b];
                            soft2[col * CS + k] = x;
                            if constexpr (DOUBLED) soft2[col * CS + Z + k] = x;
                        }
                    }
                } else {
                    for (int i4 = tid; i4 < N / 4; i4 += ZP) {                        // one Philox block -> four LLRs
                        float o[4];
                        int d[4];
                        channel_llr4_bpsk(io.ch, frame, i4, o, d);
#pragma unroll
                        for (int b = 0; b < 4; b++) {
                            const int i = d[b], col = i / Z, k = i - col * Z;
                            soft2[col * CS + k] = o[b];
                            if constexpr (DOUBLED) soft2[col * CS + Z + k] = o[b];
                        }
                    }
                    for (int j = (N & ~3) + tid; j < N; j += ZP) {                    // tail when N is not a multiple of 4
                        const int i = channel_dest(io.ch, j), col = i / Z, k = i - col * Z;
                        const float x = channel_llr(io.ch, frame, i);
                        soft2[col * CS + k] = x;
                        if constexpr (DOUBLED) soft2[col * CS + Z + k] = x;
                    }
                }
            } else if (io.llr_dtype == 1) {                      // LDPCB200_F32
                const float* y = (const float*)io.llr + (size_t)f * N;
                if (ALL_ACTIVE || active) {
#pragma unroll 8
                    for (int col = 0; col < C; col++) {
                        const float x = __ldcs(y + col * Z + tid);
                        soft2[col * CS + tid] = x;
                        if constexpr (DOUBLED) soft2[col * CS + Z + tid] = x;
                    }
                }
            } else {
                const double* y = (const double*)io.llr + (size_t)f * N;
                if (ALL_ACTIVE || active) {
#pragma unroll 8
                    for (int col = 0; col < C; col++) {
                        const float x = (float)__ldcs(y + col * Z + tid);
                        soft2[col * CS + tid] = x;
                        if constexpr (DOUBLED) soft2[col * CS + Z + tid] = x;
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < B; j++) { m1[j] = 0.0f; m2[j] = 0.0f; }                 // decoders.cpp:5088-5108
            if constexpr (PS_SMEM) {
                unsigned* psw = (unsigned*)(soft2 + SOFT_WORDS);
                for (int i = tid; i < R; i += ZP) psw[i] = 0u;
            } else {
#pragma unroll
                for (int j = 0; j < B; j++) ps[j] = 0u;
            }
            __syncthreads();

            int parity = syndrome(soft2, hb, tid);                                      // :5111-5115
            int ret = 0, locked = 0, iter;
            if (!parity) { ret = 1; locked = 1; }
            for (iter = 0; iter < io.maxiter; iter++) {
                if (!parity && !noexit) break;                                          // :5119
                layers<0>(soft2, saddr, tid, active, m1, m2, ps);
                parity = syndrome(soft2, hb, tid);                                      // :5281-5284
                if (!parity && !locked) { ret = iter + 1; locked = 1; }
                if (!parity && !noexit) break;
            }
            if (!locked) ret = parity ? -iter : iter + 1;                               // :5424

            if (io.post) {
                if (io.post_dtype == 1) {
                    float* p = (float*)io.post + (size_t)f * N;
                    for (int col = 0; col < C; col++)
                        if (ALL_ACTIVE || active) p[col * Z + tid] = soft2[col * CS + tid];
                } else {
                    double* p = (double*)io.post + (size_t)f * N;
                    for (int col = 0; col < C; col++)
                        if (ALL_ACTIVE || active) p[col * Z + tid] = (double)soft2[col * CS + tid];
                }
            }
            __syncthreads();
            if (tid == 0) { s_misc[1] = 0; s_misc[2] = 0; }
            __syncthreads();
            {
                const int lane = tid & 31;
                int nerr = 0, nerr_info = 0;
                constexpr int NROUND = (N + 31) & ~31;
                for (int i = tid; i < NROUND; i += ZP) {
                    int bit = 0;
                    if (i < N) { const int col = i / Z, k = i - col * Z; bit = soft2[col * CS + k] < 0.0f; }   // :5421
                    const unsigned w = __ballot_sync(0xffffffffu, bit);
                    if (lane == 0) {
                        if (io.hard_words) io.hard_words[(size_t)f * NWORDS + (i >> 5)] = w;
                        nerr += __popc(w);
                        const int lo = R - i;                    // bits >= R are information bits (bp_simulation.cpp:738)
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
