// Code-specialised LMS_DEC kernel (fp32), TWO FRAMES PER CTA processed by the same threads in lock step.
//
// Same arithmetic and the same results as lms_tmem.cuh (bit-identical to orc_lms_f32); same mapping (lane n of the CTA =
// check row n of every block row), same PP posterior layout (padded single-copy columns, two buffers used alternately,
// see LmsTmem::PP), c2v messages uncompressed in tensor memory.  What changes: every thread carries the check rows of two
// frames ("slots" x and y) through one instruction stream.
//
//  * Posteriors are float2 (slot x, slot y) per position: one LDS.64 / one STS.64 per edge serve both frames; the
//    messages of an edge sit in adjacent tensor-memory columns (2e, 2e + 1), so one tcgen05.ld brings register PAIRS
//    that feed add / sub / fma .f32x2 directly (FADD2 / FFMA2: one instruction, both frames).
//  * One barrier per block row for two frames; the load / barrier latencies of a block row are amortised over twice
//    the arithmetic, and the two frames' dependency chains interleave in every warp.
//  * Per edge update and frame this is 0.5 LDS + 0.56 STS + 0.3 LDTM/STTM instead of 1 + 1.13 + 0.62 -- the LSU pipe
//    (one warp instruction per clock and SM) was as tight a limit as the issue slots in the one-frame kernel.
//  * TMEM (512 columns) holds exactly the two frames' messages at E = 128, Z = 256: ONE CTA per SM, up to 255 registers.
//
// The two slots run the same block row of the same buffer parity but are otherwise independent: each has its own frame,
// iteration count and syndrome verdict; a slot whose frame is finished (syndrome clean, or maxiter reached) writes its
// results and takes the next frame of the grid's work counter while the other slot carries on.  A new frame is loaded
// into the buffer that the running parity makes "current" for each column ((parity * column weight) mod 2).
//
// Needs Z % 32 == 0 and 2 * E * ceil(Z / 128) <= 512 tensor-memory columns; other codes stay on LmsTmem.
// This header must stay free of #include (NVRTC compiles it as one string after lms_tmem.cuh).
#pragma once

namespace ldpcb200 {

static __device__ __forceinline__ unsigned long long f2_pack(float2 a)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
    return r;
}
static __device__ __forceinline__ float2 f2_unpack(unsigned long long r)
{
    float2 a;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a.x), "=f"(a.y) : "l"(r));
    return a;
}
// each half rounds exactly like the scalar add.rn / sub.rn / fma.rn
static __device__ __forceinline__ float2 f2_add(float2 a, float2 b)
{
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_pack(a)), "l"(f2_pack(b)));
    return f2_unpack(d);
}
static __device__ __forceinline__ float2 f2_sub(float2 a, float2 b)
{
    unsigned long long d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_pack(a)), "l"(f2_pack(b)));
    return f2_unpack(d);
}
static __device__ __forceinline__ float2 f2_fma(float2 a, float2 b, float2 c)
{
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(f2_pack(a)), "l"(f2_pack(b)), "l"(f2_pack(c)));
    return f2_unpack(d);
}

template <class K>
struct LmsTmem2 {
    using T = LmsTmem<K>;
    static constexpr int B = K::B, C = K::C, Z = K::Z, N = K::C * K::Z, ZP = K::ZP, E = K::E;
    static constexpr int HW = ZP / 32, NWARPS = ZP / 32, NWORDS = (N + 31) / 32;
    static constexpr int CS = Z + 32;                           // float2 per column: Z positions + the first 32 repeated
    static constexpr int BUF = C * CS;                          // float2 per buffer
    static constexpr int SOFT_WORDS = 2 * 2 * BUF;              // two buffers of float2
    static constexpr int HB_WORDS = T::HB_WORDS;
    static constexpr int HB_OFF = SOFT_WORDS;                   // packed decisions, one set per slot
    static constexpr int PLAN_OFF = HB_OFF + 2 * HB_WORDS;
    static constexpr int MISC_OFF = (PLAN_OFF + T::PLAN_WORDS + ZP + 1) & ~1;
    static constexpr int CW_OFF = MISC_OFF + 8;
    static constexpr int SMEM_WORDS = CW_OFF + C;
    static __host__ __device__ constexpr int tcols() { int t = 32; while (t < 2 * E * ((NWARPS + 3) / 4)) t *= 2; return t; }
    static constexpr int TCOLS = tcols();
    template <int J> static constexpr int NDEG = T::template NDEG<J>;

    // loads of block row J (see LmsTmem::pp_load): one float2 per edge
    template <int J, int PAR, bool WANT_EARLY, int Q = 0>
    static __device__ __forceinline__ void load(const float2* softl, const unsigned (&uoff)[NWARPS], float2 (&sv)[NDEG<J>])
    {
        if constexpr (J < B) {
            if constexpr (Q < NDEG<J>) {
                constexpr int e = K::RP[J < B ? J : 0] + Q;
                if constexpr (K::EARLY[e] == WANT_EARLY) {
                    constexpr int a = K::DELTA[e] / 32, b = K::DELTA[e] % 32;
                    constexpr int off = (T::template rbuf<PAR>(e) * C + K::COL[e]) * CS + b;
                    sv[Q] = softl[uoff[a] + off];
                }
                load<J, PAR, WANT_EARLY, Q + 1>(softl, uoff, sv);
            }
        }
    }
    template <int J, int PAR, int Q = 0>
    static __device__ __forceinline__ void put(float2* softn, unsigned* hbw, bool lane0, const float2 (&nv)[NDEG<J>])
    {
        if constexpr (Q < NDEG<J>) {
            constexpr int e = K::RP[J] + Q;
            constexpr int off = ((T::template rbuf<PAR>(e) ^ 1) * C + K::COL[e]) * CS;
            softn[off] = nv[Q];                                                                  // one float2 per lane
            if constexpr (K::LAST[e]) {                                                          // see LmsTmem::put_posterior
                const unsigned wx = __ballot_sync(0xffffffffu, nv[Q].x < 0.0f);
                const unsigned wy = __ballot_sync(0xffffffffu, nv[Q].y < 0.0f);
                if (lane0) { hbw[K::COL[e] * HW] = wx; hbw[HB_WORDS + K::COL[e] * HW] = wy; }
            }
            put<J, PAR, Q + 1>(softn, hbw, lane0, nv);
        }
    }
    // warp 0 repeats its words at positions Z .. Z+31 -- only for the columns whose next reader's DELTA is not a multiple
    // of 32 (the others never look there) -- behind a real branch (LmsTmem::pp_repeat explains the loop form)
    static __host__ __device__ constexpr bool needs_repeat(int e) { return T::next_delta(e) % 32 != 0; }
    static __host__ __device__ constexpr int rep_n(int J)
    {
        int n = 0;
        for (int e = K::RP[J]; e < K::RP[J + 1]; e++) n += needs_repeat(e) ? 1 : 0;
        return n;
    }
    static __host__ __device__ constexpr int rep_q(int J, int i)         // the i-th edge (index within the row) that needs the repeat
    {
        int n = 0;
        for (int e = K::RP[J]; e < K::RP[J + 1]; e++)
            if (needs_repeat(e)) { if (n == i) return e - K::RP[J]; n++; }
        return 0;
    }
    template <int J, int PAR, int I>
    static constexpr int rep_off = 8 * (((T::template rbuf<PAR>(K::RP[J] + rep_q(J, I)) ^ 1) * C + K::COL[K::RP[J] + rep_q(J, I)]) * CS + Z);
    template <int J, int PAR, int I = 0>
    static __device__ __forceinline__ void repeat(unsigned softn_s, unsigned notwarp0, const float2 (&nv)[NDEG<J>])
    {
#define REP2_HEAD "{\n\t.reg .pred p;\n\t.reg .u32 i;\n\tmov.u32 i, %0;\n\tREP2_LOOP:\n\tsetp.ne.u32 p, i, 0;\n\t@p bra.uni REP2_DONE;\n\t"
#define REP2_TAIL "add.u32 i, i, 1;\n\tbra.uni REP2_LOOP;\n\tREP2_DONE:\n\t}"
        constexpr int NREP = rep_n(J);
        if constexpr (I + 9 <= NREP) {
            constexpr int q0 = rep_q(J, I + 0);
            constexpr int q1 = rep_q(J, I + 1);
            constexpr int q2 = rep_q(J, I + 2);
            constexpr int q3 = rep_q(J, I + 3);
            constexpr int q4 = rep_q(J, I + 4);
            constexpr int q5 = rep_q(J, I + 5);
            constexpr int q6 = rep_q(J, I + 6);
            constexpr int q7 = rep_q(J, I + 7);
            constexpr int q8 = rep_q(J, I + 8);
            asm volatile(REP2_HEAD "st.shared.v2.f32 [%1+%2], {%11, %12};\n\tst.shared.v2.f32 [%1+%3], {%13, %14};\n\tst.shared.v2.f32 [%1+%4], {%15, %16};\n\tst.shared.v2.f32 [%1+%5], {%17, %18};\n\tst.shared.v2.f32 [%1+%6], {%19, %20};\n\tst.shared.v2.f32 [%1+%7], {%21, %22};\n\tst.shared.v2.f32 [%1+%8], {%23, %24};\n\tst.shared.v2.f32 [%1+%9], {%25, %26};\n\tst.shared.v2.f32 [%1+%10], {%27, %28};\n\t" REP2_TAIL
                         :: "r"(notwarp0), "r"(softn_s), "n"(rep_off<J, PAR, I + 0>), "n"(rep_off<J, PAR, I + 1>), "n"(rep_off<J, PAR, I + 2>), "n"(rep_off<J, PAR, I + 3>), "n"(rep_off<J, PAR, I + 4>), "n"(rep_off<J, PAR, I + 5>), "n"(rep_off<J, PAR, I + 6>), "n"(rep_off<J, PAR, I + 7>), "n"(rep_off<J, PAR, I + 8>),
                            "f"(nv[q0].x), "f"(nv[q0].y), "f"(nv[q1].x), "f"(nv[q1].y), "f"(nv[q2].x), "f"(nv[q2].y), "f"(nv[q3].x), "f"(nv[q3].y), "f"(nv[q4].x), "f"(nv[q4].y), "f"(nv[q5].x), "f"(nv[q5].y), "f"(nv[q6].x), "f"(nv[q6].y), "f"(nv[q7].x), "f"(nv[q7].y), "f"(nv[q8].x), "f"(nv[q8].y) : "memory");
            repeat<J, PAR, I + 9>(softn_s, notwarp0, nv);
        }
        else if constexpr (I + 8 <= NREP) {
            constexpr int q0 = rep_q(J, I + 0);
            constexpr int q1 = rep_q(J, I + 1);
            constexpr int q2 = rep_q(J, I + 2);
            constexpr int q3 = rep_q(J, I + 3);
            constexpr int q4 = rep_q(J, I + 4);
            constexpr int q5 = rep_q(J, I + 5);
            constexpr int q6 = rep_q(J, I + 6);
            constexpr int q7 = rep_q(J, I + 7);
            asm volatile(REP2_HEAD "st.shared.v2.f32 [%1+%2], {%10, %11};\n\tst.shared.v2.f32 [%1+%3], {%12, %13};\n\tst.shared.v2.f32 [%1+%4], {%14, %15};\n\tst.shared.v2.f32 [%1+%5], {%16, %17};\n\tst.shared.v2.f32 [%1+%6], {%18, %19};\n\tst.shared.v2.f32 [%1+%7], {%20, %21};\n\tst.shared.v2.f32 [%1+%8], {%22, %23};\n\tst.shared.v2.f32 [%1+%9], {%24, %25};\n\t" REP2_TAIL
                         :: "r"(notwarp0), "r"(softn_s), "n"(rep_off<J, PAR, I + 0>), "n"(rep_off<J, PAR, I + 1>), "n"(rep_off<J, PAR, I + 2>), "n"(rep_off<J, PAR, I + 3>), "n"(rep_off<J, PAR, I + 4>), "n"(rep_off<J, PAR, I + 5>), "n"(rep_off<J, PAR, I + 6>), "n"(rep_off<J, PAR, I + 7>),
                            "f"(nv[q0].x), "f"(nv[q0].y), "f"(nv[q1].x), "f"(nv[q1].y), "f"(nv[q2].x), "f"(nv[q2].y), "f"(nv[q3].x), "f"(nv[q3].y), "f"(nv[q4].x), "f"(nv[q4].y), "f"(nv[q5].x), "f"(nv[q5].y), "f"(nv[q6].x), "f"(nv[q6].y), "f"(nv[q7].x), "f"(nv[q7].y) : "memory");
            repeat<J, PAR, I + 8>(softn_s, notwarp0, nv);
        }
        else if constexpr (I + 7 <= NREP) {
            constexpr int q0 = rep_q(J, I + 0);
            constexpr int q1 = rep_q(J, I + 1);
            constexpr int q2 = rep_q(J, I + 2);
            constexpr int q3 = rep_q(J, I + 3);
            constexpr int q4 = rep_q(J, I + 4);
            constexpr int q5 = rep_q(J, I + 5);
            constexpr int q6 = rep_q(J, I + 6);
            asm volatile(REP2_HEAD "st.shared.v2.f32 [%1+%2], {%9, %10};\n\tst.shared.v2.f32 [%1+%3], {%11, %12};\n\tst.shared.v2.f32 [%1+%4], {%13, %14};\n\tst.shared.v2.f32 [%1+%5], {%15, %16};\n\tst.shared.v2.f32 [%1+%6], {%17, %18};\n\tst.shared.v2.f32 [%1+%7], {%19, %20};\n\tst.shared.v2.f32 [%1+%8], {%21, %22};\n\t" REP2_TAIL
                         :: "r"(notwarp0), "r"(softn_s), "n"(rep_off<J, PAR, I + 0>), "n"(rep_off<J, PAR, I + 1>), "n"(rep_off<J, PAR, I + 2>), "n"(rep_off<J, PAR, I + 3>), "n"(rep_off<J, PAR, I + 4>), "n"(rep_off<J, PAR, I + 5>), "n"(rep_off<J, PAR, I + 6>),
                            "f"(nv[q0].x), "f"(nv[q0].y), "f"(nv[q1].x), "f"(nv[q1].y), "f"(nv[q2].x), "f"(nv[q2].y), "f"(nv[q3].x), "f"(nv[q3].y), "f"(nv[q4].x), "f"(nv[q4].y), "f"(nv[q5].x), "f"(nv[q5].y), "f"(nv[q6].x), "f"(nv[q6].y) : "memory");
            repeat<J, PAR, I + 7>(softn_s, notwarp0, nv);
        }
        else if constexpr (I + 6 <= NREP) {
            constexpr int q0 = rep_q(J, I + 0);
            constexpr int q1 = rep_q(J, I + 1);
            constexpr int q2 = rep_q(J, I + 2);
            constexpr int q3 = rep_q(J, I + 3);
            constexpr int q4 = rep_q(J, I + 4);
            constexpr int q5 = rep_q(J, I + 5);
            asm volatile(REP2_HEAD "st.shared.v2.f32 [%1+%2], {%8, %9};\n\tst.shared.v2.f32 [%1+%3], {%10, %11};\n\tst.shared.v2.f32 [%1+%4], {%12, %13};\n\tst.shared.v2.f32 [%1+%5], {%14, %15};\n\tst.shared.v2.f32 [%1+%6], {%16, %17};\n\tst.shared.v2.f32 [%1+%7], {%18, %19};\n\t" REP2_TAIL
                         :: "r"(notwarp0), "r"(softn_s), "n"(rep_off<J, PAR, I + 0>), "n"(rep_off<J, PAR, I + 1>), "n"(rep_off<J, PAR, I + 2>), "n"(rep_off<J, PAR, I + 3>), "n"(rep_off<J, PAR, I + 4>), "n"(rep_off<J, PAR, I + 5>),
                            "f"(nv[q0].x), "f"(nv[q0].y), "f"(nv[q1].x), "f"(nv[q1].y), "f"(nv[q2].x), "f"(nv[q2].y), "f"(nv[q3].x), "f"(nv[q3].y), "f"(nv[q4].x), "f"(nv[q4].y), "f"(nv[q5].x), "f"(nv[q5].y) : "memory");
            repeat<J, PAR, I + 6>(softn_s, notwarp0, nv);
        }
        else if constexpr (I + 5 <= NREP) {
            constexpr int q0 = rep_q(J, I + 0);
            constexpr int q1 = rep_q(J, I + 1);
            constexpr int q2 = rep_q(J, I + 2);
            constexpr int q3 = rep_q(J, I + 3);
            constexpr int q4 = rep_q(J, I + 4);
            asm volatile(REP2_HEAD "st.shared.v2.f32 [%1+%2], {%7, %8};\n\tst.shared.v2.f32 [%1+%3], {%9, %10};\n\tst.shared.v2.f32 [%1+%4], {%11, %12};\n\tst.shared.v2.f32 [%1+%5], {%13, %14};\n\tst.shared.v2.f32 [%1+%6], {%15, %16};\n\t" REP2_TAIL
                         :: "r"(notwarp0), "r"(softn_s), "n"(rep_off<J, PAR, I + 0>), "n"(rep_off<J, PAR, I + 1>), "n"(rep_off<J, PAR, I + 2>), "n"(rep_off<J, PAR, I + 3>), "n"(rep_off<J, PAR, I + 4>),
                            "f"(nv[q0].x), "f"(nv[q0].y), "f"(nv[q1].x), "f"(nv[q1].y), "f"(nv[q2].x), "f"(nv[q2].y), "f"(nv[q3].x), "f"(nv[q3].y), "f"(nv[q4].x), "f"(nv[q4].y) : "memory");
            repeat<J, PAR, I + 5>(softn_s, notwarp0, nv);
        }
        else if constexpr (I + 4 <= NREP) {
            constexpr int q0 = rep_q(J, I + 0);
            constexpr int q1 = rep_q(J, I + 1);
            constexpr int q2 = rep_q(J, I + 2);
            constexpr int q3 = rep_q(J, I + 3);
            asm volatile(REP2_HEAD "st.shared.v2.f32 [%1+%2], {%6, %7};\n\tst.shared.v2.f32 [%1+%3], {%8, %9};\n\tst.shared.v2.f32 [%1+%4], {%10, %11};\n\tst.shared.v2.f32 [%1+%5], {%12, %13};\n\t" REP2_TAIL
                         :: "r"(notwarp0), "r"(softn_s), "n"(rep_off<J, PAR, I + 0>), "n"(rep_off<J, PAR, I + 1>), "n"(rep_off<J, PAR, I + 2>), "n"(rep_off<J, PAR, I + 3>),
                            "f"(nv[q0].x), "f"(nv[q0].y), "f"(nv[q1].x), "f"(nv[q1].y), "f"(nv[q2].x), "f"(nv[q2].y), "f"(nv[q3].x), "f"(nv[q3].y) : "memory");
            repeat<J, PAR, I + 4>(softn_s, notwarp0, nv);
        }
        else if constexpr (I + 3 <= NREP) {
            constexpr int q0 = rep_q(J, I + 0);
            constexpr int q1 = rep_q(J, I + 1);
            constexpr int q2 = rep_q(J, I + 2);
            asm volatile(REP2_HEAD "st.shared.v2.f32 [%1+%2], {%5, %6};\n\tst.shared.v2.f32 [%1+%3], {%7, %8};\n\tst.shared.v2.f32 [%1+%4], {%9, %10};\n\t" REP2_TAIL
                         :: "r"(notwarp0), "r"(softn_s), "n"(rep_off<J, PAR, I + 0>), "n"(rep_off<J, PAR, I + 1>), "n"(rep_off<J, PAR, I + 2>),
                            "f"(nv[q0].x), "f"(nv[q0].y), "f"(nv[q1].x), "f"(nv[q1].y), "f"(nv[q2].x), "f"(nv[q2].y) : "memory");
            repeat<J, PAR, I + 3>(softn_s, notwarp0, nv);
        }
        else if constexpr (I + 2 <= NREP) {
            constexpr int q0 = rep_q(J, I + 0);
            constexpr int q1 = rep_q(J, I + 1);
            asm volatile(REP2_HEAD "st.shared.v2.f32 [%1+%2], {%4, %5};\n\tst.shared.v2.f32 [%1+%3], {%6, %7};\n\t" REP2_TAIL
                         :: "r"(notwarp0), "r"(softn_s), "n"(rep_off<J, PAR, I + 0>), "n"(rep_off<J, PAR, I + 1>),
                            "f"(nv[q0].x), "f"(nv[q0].y), "f"(nv[q1].x), "f"(nv[q1].y) : "memory");
            repeat<J, PAR, I + 2>(softn_s, notwarp0, nv);
        }
        else if constexpr (I + 1 <= NREP) {
            constexpr int q0 = rep_q(J, I + 0);
            asm volatile(REP2_HEAD "st.shared.v2.f32 [%1+%2], {%3, %4};\n\t" REP2_TAIL
                         :: "r"(notwarp0), "r"(softn_s), "n"(rep_off<J, PAR, I + 0>),
                            "f"(nv[q0].x), "f"(nv[q0].y) : "memory");
            repeat<J, PAR, I + 1>(softn_s, notwarp0, nv);
        }
#undef REP2_HEAD
#undef REP2_TAIL
    }

    // msg / sv: the row's old messages (2 per edge: slot x, slot y; tcgen05.ld in flight) and the early posteriors, both issued
    // by the previous block row; msgn / svn: the same for the next one
    template <int J, int PAR>
    static __device__ __forceinline__ void layer(float2* softn, const float2* softl, const unsigned (&uoff)[NWARPS], unsigned* hbw, unsigned trow,
                                                 bool lane0, bool warp0, unsigned (&msg)[2 * NDEG<J>], float2 (&sv)[NDEG<J>],
                                                 unsigned (&msgn)[2 * NDEG<J + 1>], float2 (&svn)[NDEG<J + 1>])
    {
        constexpr int E0 = K::RP[J], DEG = NDEG<J>;
        float vx[DEG], vy[DEG], mx[DEG], my[DEG];
        float2 nv[DEG];
        load<J, PAR, false>(softl, uoff, sv);
        tmem_wait_ld<2 * DEG>(msg);
#pragma unroll
        for (int q = 0; q < DEG; q++) {                                                          // decoders.cpp:5152-5158, both slots per FADD2
            const float2 v = f2_sub(sv[q], make_float2(__uint_as_float(msg[2 * q]), __uint_as_float(msg[2 * q + 1])));
            vx[q] = v.x; vy[q] = v.y;
        }
        // messages as minima over the other edges: see LmsTmem::layer
        const unsigned sx = T::template sign_xor<DEG, 0, DEG>(vx) & 0x80000000u, sy = T::template sign_xor<DEG, 0, DEG>(vy) & 0x80000000u;
        const float ronex = __uint_as_float(sx | 0x3f800000u), roney = __uint_as_float(sy | 0x3f800000u);
        const float2 rhalf = make_float2(__fmul_rn(ronex, 0.5f), __fmul_rn(roney, 0.5f));
        const float2 nhalf = make_float2(__fmul_rn(rhalf.x, -0.4f), __fmul_rn(rhalf.y, -0.4f));
        T::template min_of_others<DEG>(vx, mx, 32767.400390625f);
        T::template min_of_others<DEG>(vy, my, 32767.400390625f);
#pragma unroll
        for (int q = 0; q < DEG; q++) {
            const float2 th = f2_fma(make_float2(mx[q], my[q]), rhalf, nhalf);
            msg[2 * q] = __float_as_uint(__fmaf_rn(fabsf(th.x), ronex, th.x)) ^ (__float_as_uint(vx[q]) & 0x80000000u);
            msg[2 * q + 1] = __float_as_uint(__fmaf_rn(fabsf(th.y), roney, th.y)) ^ (__float_as_uint(vy[q]) & 0x80000000u);
            nv[q] = f2_add(make_float2(vx[q], vy[q]), make_float2(__uint_as_float(msg[2 * q]), __uint_as_float(msg[2 * q + 1])));   // :5199-5204
        }
        put<J, PAR>(softn, hbw, lane0, nv);
        repeat<J, PAR>((unsigned)__cvta_generic_to_shared(softn), warp0 ? 0u : 1u, nv);
        tmem_st_n<2 * DEG>(trow + 2 * E0, msg);                                                  // :5179
        if constexpr (J + 1 < B) {
            constexpr int E1 = K::RP[J + 1 < B ? J + 1 : 0];
            tmem_ld_n<2 * NDEG<J + 1>>(trow + 2 * E1, msgn);
            load<J + 1, PAR, true>(softl, uoff, svn);
        }
    }
    template <int J, int PAR>
    static __device__ __forceinline__ void layers(float2* softn, const float2* softl, const unsigned (&uoff)[NWARPS], unsigned* hbw, unsigned trow,
                                                  bool lane0, bool warp0, unsigned (&msg)[2 * NDEG<J>], float2 (&sv)[NDEG<J>])
    {
        if constexpr (J < B) {
            unsigned msgn[2 * NDEG<J + 1>];
            float2 svn[NDEG<J + 1>];
            layer<J, PAR>(softn, softl, uoff, hbw, trow, lane0, warp0, msg, sv, msgn, svn);
            __syncthreads();
            layers<J + 1, PAR>(softn, softl, uoff, hbw, trow, lane0, warp0, msgn, svn);
        }
    }
    template <int PAR>
    static __device__ __forceinline__ void iteration(float2* softn, const float2* softl, const unsigned (&uoff)[NWARPS], unsigned* hbw, unsigned trow,
                                                     bool lane0, bool warp0)
    {
        unsigned msg[2 * NDEG<0>];
        float2 sv[NDEG<0>];
        tmem_wait_st();                                                                          // last iteration's messages are in place
        tmem_ld_n<2 * NDEG<0>>(trow, msg);
        layers<0, PAR>(softn, softl, uoff, hbw, trow, lane0, warp0, msg, sv);
    }

    // ---- one slot's frame: load, results
    // word (32 bit) of position p of column col in buffer b, slot s
    static __device__ __forceinline__ int widx(int b, int col, int p, int s) { return 2 * ((b * C + col) * CS + p) + s; }
    static __device__ __forceinline__ void store_pos(float* sw, int b, int col, int p, int s, float x)
    {
        sw[widx(b, col, p, s)] = x;
        if (p < 32) sw[widx(b, col, p + Z, s)] = x;
    }

    // the slot's messages := 0 (prev[] = 0, decoders.cpp:5088-5108): read-modify-write, the other slot's stay
    template <int OFF = 0>
    static __device__ __forceinline__ void zero_messages(unsigned trow, int s)
    {
        if constexpr (OFF < 2 * E) {
            constexpr int P = tmem_chunk(2 * E - OFF);
            unsigned r[P];
            TmemRow<P>::ld(trow + OFF, r);
            tmem_wait_ld<P>(r);
#pragma unroll
            for (int i = 0; i < P; i++) if (((OFF + i) & 1) == s) r[i] = 0u;
            TmemRow<P>::st(trow + OFF, r);
            zero_messages<OFF + P>(trow, s);
        }
    }

    static __device__ __forceinline__ void load_frame(const FrameIO& io, float* sw, unsigned* hb, const int* s_cw, unsigned trow, int s, int f, int par, int tid)
    {
        const int lane = tid & 31, warp = tid >> 5;
        unsigned* hbs = hb + s * HB_WORDS;
        bool packed = false;
        if (io.ch.enabled) {
            const unsigned long long frame = io.ch.first_frame + (unsigned long long)f;
            if (io.ch.m > 2) {
                const int half = io.ch.m >> 1, ncomp = 2 * (N / io.ch.m);
                for (int c = tid; c < ncomp; c += ZP) {
                    float o[4];
                    channel_llr_qam_component(io.ch, frame, c, o);
                    const int i0 = (c >> 1) * io.ch.m + (c & 1) * half;
                    for (int b = 0; b < half; b++) {
                        const int i = channel_dest(io.ch, i0 + b), col = i / Z, k = i - col * Z;
                        store_pos(sw, (par * s_cw[col]) & 1, col, T::pos_of(col, k), s, i >= io.ch.punct_start ? io.ch.punct_value : o[b]);
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
                        store_pos(sw, (par * s_cw[col]) & 1, col, T::pos_of(col, k), s, o[b]);
                    }
                }
            }
        } else if (io.llr_dtype == 1) {                      // LDPCB200_F32
            const float* y = (const float*)io.llr + (size_t)f * N;
#pragma unroll
            for (int c0 = 0; c0 < C; c0 += 16) {             // 16 loads in flight per thread
                float x[16];
#pragma unroll
                for (int u = 0; u < 16; u++) {               // position tid of column col holds bit (tid + ROT) mod Z
                    const int col = c0 + u;
                    if (col < C) {
                        int k = tid + K::rt_rot()[col];
                        if (k >= Z) k -= Z;
                        x[u] = __ldcs(y + col * Z + k);
                    }
                }
#pragma unroll
                for (int u = 0; u < 16; u++) {
                    const int col = c0 + u;
                    if (col < C) {
                        store_pos(sw, (par * s_cw[col]) & 1, col, tid, s, x[u]);
                        const unsigned w = __ballot_sync(0xffffffffu, x[u] < 0.0f);             // the packed decisions of the channel values
                        if (lane == 0) hbs[col * HW + warp] = w;
                    }
                }
            }
            packed = true;
        } else {
            const double* y = (const double*)io.llr + (size_t)f * N;
#pragma unroll 8
            for (int col = 0; col < C; col++) {
                int k = tid + K::rt_rot()[col];
                if (k >= Z) k -= Z;
                store_pos(sw, (par * s_cw[col]) & 1, col, tid, s, (float)__ldcs(y + col * Z + k));
            }
        }
        zero_messages(trow, s);
        tmem_wait_st();
        __syncthreads();
        if (!packed) {
#pragma unroll 8
            for (int col = 0; col < C; col++) {
                const unsigned w = __ballot_sync(0xffffffffu, sw[widx((par * s_cw[col]) & 1, col, tid, s)] < 0.0f);
                if (lane == 0) hbs[col * HW + warp] = w;
            }
            __syncthreads();
        }
    }

    // results of slot s (frame f): posteriors of buffer parity `par`, decisions from the slot's packed words
    static __device__ __forceinline__ void finish(const FrameIO& io, const float* sw, const unsigned* hb, const int* s_cw, int* s_misc, int s, int f, int par, int ret, int tid)
    {
        const unsigned* hbs = hb + s * HB_WORDS;
        if (tid == 0) { s_misc[1] = 0; s_misc[2] = 0; }
        __syncthreads();
        if (io.post) {
            if (io.post_dtype == 1) {
                float* p = (float*)io.post + (size_t)f * N;
                for (int col = 0; col < C; col++) p[col * Z + tid] = sw[widx((par * s_cw[col]) & 1, col, T::pos_of(col, tid), s)];
            } else {
                double* p = (double*)io.post + (size_t)f * N;
                for (int col = 0; col < C; col++) p[col * Z + tid] = (double)sw[widx((par * s_cw[col]) & 1, col, T::pos_of(col, tid), s)];
            }
        }
        // error counts are popcounts of the packed words (rotation-invariant); block columns >= B are information bits
        // (bp_simulation.cpp:738)
        const int lane = tid & 31;
        int nerr = 0, nerr_info = 0;
        for (int t = tid; t < ((C * HW + 31) & ~31); t += ZP) {
            const unsigned w = t < C * HW ? hbs[t] : 0u;
            const int pc = __popc(w);
            nerr += pc;
            if (t >= B * HW) nerr_info += pc;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            nerr += __shfl_xor_sync(0xffffffffu, nerr, o);
            nerr_info += __shfl_xor_sync(0xffffffffu, nerr_info, o);
        }
        if (lane == 0 && nerr) { atomicAdd(&s_misc[1], nerr); atomicAdd(&s_misc[2], nerr_info); }
        if (io.hard_words) {
            unsigned* out = io.hard_words + (size_t)f * NWORDS;
            for (int g = tid; g < NWORDS; g += ZP) {           // output word g = a 32-bit window of the column's packed words
                const int col = g / HW, k0 = 32 * (g - col * HW);
                int start = k0 + K::rt_ri()[col];
                if (start >= Z) start -= Z;
                const unsigned* hc = hbs + col * HW;
                const int i0 = start >> 5, i1 = i0 + 1 < HW ? i0 + 1 : 0;
                out[g] = __funnelshift_r(hc[i0], hc[i1], start & 31);
            }
        }
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

    static __device__ __forceinline__ void kernel(const FrameIO& io)
    {
        extern __shared__ __align__(16) float smem_f[];
        float2* soft2 = (float2*)smem_f;
        unsigned* hb = (unsigned*)(smem_f + HB_OFF);
        unsigned* plan = (unsigned*)(smem_f + PLAN_OFF);
        int* s_misc = (int*)(smem_f + MISC_OFF);
        int* s_cw = (int*)(smem_f + CW_OFF);
        const int tid = threadIdx.x;
        const bool lane0 = (tid & 31) == 0;
        const bool noexit = io.flags & 8u;                       // LDPCB200_NO_EARLY_EXIT
        float2* softn = soft2 + tid;
        unsigned* hbw = hb + (tid >> 5);
        const unsigned wu = __shfl_sync(0xffffffffu, (unsigned)tid >> 5, 0);   // warp-uniform pieces of the read addresses
        const bool warp0 = wu == 0;
        const float2* softl = soft2 + (tid & 31);
        unsigned uoff[NWARPS];
#pragma unroll
        for (int a = 0; a < NWARPS; a++) uoff[a] = ((wu + a) % NWARPS) * 32u;
        for (int col = tid; col < C; col += ZP) {
            int wgt = 0;
            for (int e = 0; e < E; e++) wgt += K::rt_col()[e] == col;
            s_cw[col] = wgt;
        }
        T::build_plan(plan, tid);
        if (tid == 0) { hb[HB_WORDS - 1] = 0u; hb[2 * HB_WORDS - 1] = 0u; }

        // tensor memory: one warp allocates TCOLS columns for the CTA and frees them at the end
        if (tid < 32) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                         :: "r"((unsigned)__cvta_generic_to_shared(s_misc + 4)), "r"((unsigned)TCOLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const unsigned tbase = *(volatile unsigned*)(s_misc + 4);
        const unsigned trow = __shfl_sync(0xffffffffu, tbase + ((unsigned)(((tid >> 5) & 3) * 32) << 16) + (unsigned)((tid >> 7) * 2 * E), 0);

        int fr[2] = { -1, -1 }, it[2] = { 0, 0 }, ret[2] = { 0, 0 }, locked[2] = { 0, 0 }, parity[2] = { 1, 1 };
        bool live[2] = { false, false }, more = true;
        int par = 0;                                             // buffer parity of the next iteration (see LmsTmem::rbuf)
        for (;;) {
#pragma unroll
            for (int s = 0; s < 2; s++) {
                while (!live[s] && more) {                       // give the slot a frame that needs iterations
                    __syncthreads();
                    if (tid == 0) s_misc[0] = (int)atomicAdd(io.next_frame, 1u);
                    __syncthreads();
                    const int f = s_misc[0];
                    if (f >= io.nf) { more = false; break; }
                    load_frame(io, smem_f, hb, s_cw, trow, s, f, par, tid);
                    const int bad = T::syndrome(hb + s * HB_WORDS, plan, tid);              // :5111-5115
                    fr[s] = f; it[s] = 0; ret[s] = bad ? 0 : 1; locked[s] = !bad; parity[s] = bad;
                    if ((!bad && !noexit) || io.maxiter <= 0) finish(io, smem_f, hb, s_cw, s_misc, s, f, par, ret[s], tid);   // :5119
                    else live[s] = true;
                }
            }
            if (!live[0] && !live[1]) break;
            if (par) iteration<1>(softn, softl, uoff, hbw, trow, lane0, warp0);
            else iteration<0>(softn, softl, uoff, hbw, trow, lane0, warp0);
            par ^= 1;
#pragma unroll
            for (int s = 0; s < 2; s++) {
                if (!live[s]) continue;
                it[s]++;
                const int bad = T::syndrome(hb + s * HB_WORDS, plan, tid);                  // :5281-5284
                if (!locked[s]) { parity[s] = bad; if (!bad) { ret[s] = it[s]; locked[s] = 1; } }
                if ((!bad && !noexit) || it[s] >= io.maxiter) {
                    if (!locked[s]) ret[s] = -it[s];                                        // :5424
                    finish(io, smem_f, hb, s_cw, s_misc, s, fr[s], par, ret[s], tid);
                    live[s] = false;
                }
            }
        }

        // every thread's TMEM traffic is complete (wait::ld / wait::st above); hand the columns back
        tmem_wait_st();
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid < 32) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase), "r"((unsigned)TCOLS) : "memory");
        }
    }
};

} // namespace ldpcb200
