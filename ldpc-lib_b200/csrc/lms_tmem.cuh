// Code-specialised LMS_DEC kernel (fp32) with the check-to-variable messages kept in TENSOR MEMORY.
//
// Same arithmetic as lms_spec.cuh / lms_fast.cu (bit-identical results to orc_lms_f32), one frame per CTA at a
// time, lane n of the CTA = check row n of every block row.  What changes is where the decoder state lives:
//
//  * c2v messages -- one fp32 word per edge and lane, E*Z words per frame (128 KB at E = 128, Z = 256) -- are
//    stored UNCOMPRESSED in the SM's tensor memory (TMEM, 512 columns x 128 lanes x 32 bit, sm_100a).  A thread
//    owns TMEM lane 32*(warp%4)+lane and the columns [group*E, group*E + E), group = warp/4; block row J is the
//    column range RP[J]..RP[J+1]-1.  One `tcgen05.ld.32x32b.xN` brings a whole row's old messages into
//    registers and one `tcgen05.st` puts the new ones back: 2 instructions per ROW instead of the
//    select / shift / xor / bit-insert bookkeeping per EDGE that the register-compressed {min1, min2, sign word,
//    minimum flags} form of lms_spec.cuh needs (reference: prev[] {min1,min2,pos,sign} + signs[],
//    decoders.cpp:5152-5158, 5179).  The tensor cores themselves are not used -- there is no contraction here;
//    TMEM is the one on-chip store big enough for the messages that costs neither registers nor shared-memory
//    bandwidth.
//  * posteriors in shared memory, every block column stored twice back to back (2Z words) and kept in the
//    ROTATION OF ITS LAST WRITER: after block row J updated column k through shift s, lane n holds bit
//    (n + s) mod Z and writes it to positions n and n + Z -- two plain stores with immediate offsets, no wrap
//    predicate; the next block row J' reads its bit (n + s') mod Z at position n + ((s' - s) mod Z), one plain
//    load.  The rotation of every column at every point of the schedule is known at compile time (K::DELTA per
//    edge); at iteration boundaries column k is rotated by the shift of its last block row (K::ROT), which
//    the frame load, the syndrome windows (K::SYNSH) and the outputs take into account (K::RI = (Z - ROT) mod Z).
//
// Per edge-update this leaves: 1 LDS + 2 STS, one FADD2 (two edges per instruction), ~2.7 FMNMX (two smallest of the row
// through a balanced tree + the offset / clamp per row), 1/2 LOP3 (sign parity), FSETP + SEL + LOP3 (new message),
// 0.6 LDTM / STTM: 12.8 SASS instructions, 6.9 of them on the ALU pipe (tools/sass_mix.py) -- half of lms_spec.cuh.
//
// Generated `Code` (tools/gen_lms_spec.py kind "lmst", spec_jit.cpp variant 2) adds to the lms_spec fields:
//   static constexpr int DELTA[E], ROT[C], RI[C], SYNSH[E], TCOLS; bool LAST[E], EARLY[E];  rt_rot() / rt_ri() / rt_synsh() __constant__ copies.
// This header must stay free of #include (NVRTC compiles it as one string after lms_spec.cuh).
#pragma once
#ifndef LMS_TMEM_OTHERS
#define LMS_TMEM_OTHERS 1
#endif
#ifndef LMS_TMEM_PP
#define LMS_TMEM_PP 1           // Z a multiple of 32: padded single-copy columns, two buffers used alternately (see PP below)
#endif
#ifndef LMS_TMEM_FFMA2
#define LMS_TMEM_FFMA2 0        // PP: the offset / scale FFMA of two edges as one fma.rn.f32x2
#endif
#ifndef LMS_TMEM_WHATIF
#define LMS_TMEM_WHATIF 0       // development, WRONG RESULTS: 1 no barrier between block rows, 2 no minima, 3 no posterior stores, 4 no repeat stores
#endif
#ifndef LMS_TMEM_PP_BAL
#define LMS_TMEM_PP_BAL 0       // PP: column k is stored rotated by k mod NWARPS warps, so the repeat stores spread over the warps
#endif
#ifndef LMS_TMEM_PP_BRANCH
#define LMS_TMEM_PP_BRANCH 1    // PP: warp 0's repeat stores behind a branch (0: predicated stores in every warp)
#endif

namespace ldpcb200 {

// ---- tensor-memory plumbing (PTX ISA: tcgen05.alloc / ld / st / wait / dealloc, sm_100a)
template <int NX> struct TmemRow;
template <> struct TmemRow<1> {
    static __device__ __forceinline__ void ld(unsigned t, unsigned* r)
    { asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r[0]) : "r"(t)); }
    static __device__ __forceinline__ void st(unsigned t, const unsigned* r)
    { asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" :: "r"(t), "r"(r[0]) : "memory"); }
};
template <> struct TmemRow<2> {
    static __device__ __forceinline__ void ld(unsigned t, unsigned* r)
    { asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(t)); }
    static __device__ __forceinline__ void st(unsigned t, const unsigned* r)
    { asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" :: "r"(t), "r"(r[0]), "r"(r[1]) : "memory"); }
};
template <> struct TmemRow<4> {
    static __device__ __forceinline__ void ld(unsigned t, unsigned* r)
    { asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(t)); }
    static __device__ __forceinline__ void st(unsigned t, const unsigned* r)
    { asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                   :: "r"(t), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory"); }
};
template <> struct TmemRow<8> {
    static __device__ __forceinline__ void ld(unsigned t, unsigned* r)
    { asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(t)); }
    static __device__ __forceinline__ void st(unsigned t, const unsigned* r)
    { asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                   :: "r"(t), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory"); }
};
template <> struct TmemRow<16> {
    static __device__ __forceinline__ void ld(unsigned t, unsigned* r)
    { asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                     "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(t)); }
    static __device__ __forceinline__ void st(unsigned t, const unsigned* r)
    { asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                   :: "r"(t), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory"); }
};

// the registers of a tcgen05.ld are valid only after this wait; passing them through the statement as
// read-write operands keeps the compiler from scheduling their first use above it
template <int N, int I = 0>
static __device__ __forceinline__ void tmem_touch(unsigned (&r)[N])
{
    if constexpr (I + 4 <= N) { asm volatile("" : "+r"(r[I]), "+r"(r[I + 1]), "+r"(r[I + 2]), "+r"(r[I + 3])); tmem_touch<N, I + 4>(r); }
    else if constexpr (I < N) { asm volatile("" : "+r"(r[I])); tmem_touch<N, I + 1>(r); }
}
template <int N>
static __device__ __forceinline__ void tmem_wait_ld(unsigned (&r)[N])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    tmem_touch<N>(r);
}
static __device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__host__ __device__ constexpr int tmem_chunk(int n) { return n >= 16 ? 16 : n >= 8 ? 8 : n >= 4 ? 4 : n >= 2 ? 2 : 1; }

// N consecutive columns starting at column address t, in power-of-two pieces (13 = 8 + 4 + 1)
template <int N, int OFF = 0>
static __device__ __forceinline__ void tmem_ld_n(unsigned t, unsigned (&r)[N])
{
    if constexpr (OFF < N) {
        constexpr int P = tmem_chunk(N - OFF);
        TmemRow<P>::ld(t + OFF, &r[OFF]);
        tmem_ld_n<N, OFF + P>(t, r);
    }
}
template <int N, int OFF = 0>
static __device__ __forceinline__ void tmem_st_n(unsigned t, const unsigned (&r)[N])
{
    if constexpr (OFF < N) {
        constexpr int P = tmem_chunk(N - OFF);
        TmemRow<P>::st(t + OFF, &r[OFF]);
        tmem_st_n<N, OFF + P>(t, r);
    }
}
template <int N, int OFF = 0>
static __device__ __forceinline__ void tmem_zero_n(unsigned t)
{
    if constexpr (OFF < N) {
        constexpr int P = tmem_chunk(N - OFF);
        unsigned z[P];
#pragma unroll
        for (int i = 0; i < P; i++) z[i] = 0u;
        TmemRow<P>::st(t + OFF, z);
        tmem_zero_n<N, OFF + P>(t);
    }
}

// two fp32 additions in one instruction (sm_100a FADD2: add.rn.f32x2 on register pairs); each half rounds exactly
// like a scalar add.rn.f32
static __device__ __forceinline__ void add_f32x2(float& d0, float& d1, float a0, float a1, float b0, float b1)
{
    unsigned long long a, b, d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}
static __device__ __forceinline__ void sub_f32x2(float& d0, float& d1, float a0, float a1, float b0, float b1)
{
    unsigned long long a, b, d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}

static __device__ __forceinline__ void fma_f32x2(float& d0, float& d1, float a0, float a1, float b0, float b1, float c0, float c1)
{
    unsigned long long a, b, c, d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(c0), "f"(c1));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}

template <class K>
struct LmsTmem {
    using S = LmsSpec<K>;
    static constexpr int B = K::B, C = K::C, Z = K::Z, N = K::C * K::Z, R = K::B * K::Z, ZP = K::ZP, E = K::E;
    static constexpr int HW = ZP / 32;
    static constexpr int NB = (Z + 31) / 32;
    static constexpr int NWORDS = (N + 31) / 32;
    static constexpr bool ALL_ACTIVE = (Z == ZP);
    // ---- PP layout (Z a multiple of 32).  A block column is Z + 32 words: position p < Z, and positions Z .. Z+31
    // repeat 0 .. 31.  Lane n = 32w + l reads position (n + DELTA) mod Z, DELTA = 32a + b, as word
    // 32 * ((w + a) mod NWARPS) + (l + b) <= Z + 30: the warp part is a warp-uniform base (one per value of a, held in
    // uniform registers), l + b needs no wrap because of the 32 repeated words -- one plain load, as with the doubled
    // column, but the writer stores ONE word per lane (warp 0 also stores the repeat): 2.125 shared-memory wavefronts
    // per edge instead of 3.  Every column has two such buffers; a block row reads the column's current buffer and
    // writes the other one, so no lane can overwrite a word that another warp of the same block row has yet to read
    // -- the write-after-read hazard that the doubled layout closes with the split mbarrier does not exist.  Which
    // buffer is current depends on how often the column has been written: (iteration * column weight + edges of the
    // column above this one) mod 2, a compile-time constant once two consecutive iterations are unrolled (PAR).
    static constexpr bool PP = LMS_TMEM_PP && (Z % 32 == 0);
    static constexpr int CS = PP ? Z + 32 : 2 * Z;
    static constexpr int BUF = C * CS;                        // PP: word offset of the second buffer
    static constexpr int SOFT_WORDS = PP ? 2 * C * CS : C * CS;
    // PP_BAL: the words of column k sit SEG(k) warps further (cyclically): lane n = 32w + l writes word 32 * ((w + SEG) mod NWARPS) + l,
    // so the 32 repeated words of column k are the ones of warp (NWARPS - SEG(k)) mod NWARPS -- a different warp for every column
    // of a block row instead of warp 0 for all of them (with warp 0 as the only one, every block row waited for it at its barrier)
    static __host__ __device__ constexpr int SEG(int col) { return (LMS_TMEM_PP_BAL && PP) ? col % (ZP / 32) : 0; }
    static __device__ __forceinline__ int seg_words(unsigned wu, int col) { return (int)((wu + (unsigned)SEG(col)) % (unsigned)(ZP / 32)) * 32; }
    static __device__ __forceinline__ int phys(int col, int p) { int q = p + 32 * SEG(col); return q >= Z ? q - Z : q; }
    static constexpr int TCOLS = K::TCOLS;                    // power of two >= 32, >= E * ceil(warps / 4)
    static constexpr int NWARPS = ZP / 32;
    // shared memory (words): posteriors | packed decisions hb (+ one zero word) | syndrome plan | mbarrier (8-byte aligned) | misc
    static constexpr int HB_WORDS = (C * HW > 3 ? C * HW : 3) + 1;
    static constexpr int PLAN_OFF = SOFT_WORDS + (PP ? 2 : 1) * HB_WORDS;      // PP: the decisions of two consecutive iterations
    static constexpr int PLAN_WORDS = ((B * NB + 8 * NWARPS - 1) / (8 * NWARPS)) * ((K::MAXDEG + 3) / 4) * ZP;
    static constexpr int PLAN1_OFF = PLAN_OFF + PLAN_WORDS;      // quick syndrome look (syndrome_quick): one word per thread
    static constexpr int MBAR_OFF = (PLAN1_OFF + ZP + 1) & ~1;
    static constexpr int MISC_OFF = MBAR_OFF + 2;
    static constexpr int CW_OFF = MISC_OFF + 4;                  // PP: column weights (which buffer holds the result)
    static constexpr int SMEM_WORDS = CW_OFF + (PP ? C : 0);

    // Edge Q of block row J reads a column that block row J - 1 does not write: its value is final once the barrier
    // before block row J - 1 has passed, so the load can be issued one block row early (and its latency hidden behind
    // that row's arithmetic and barrier) -- prefetch() below.  Block row 0 follows the syndrome check: nothing early.
    // K::EARLY[e] (generated) says so; NDEG<J> = weight of block row J (1 for J = B, the array that is never used).
    template <int J> static constexpr int NDEG = J < B ? K::RP[(J < B ? J : 0) + 1] - K::RP[J < B ? J : 0] : 1;

    template <int J, int Q>
    static __device__ __forceinline__ void load_soft(const float* softn, float (&sv)[K::RP[J + 1] - K::RP[J]], const float (&pre)[NDEG<J>])
    {
        constexpr int E0 = K::RP[J], DEG = K::RP[J + 1] - K::RP[J];
        if constexpr (Q < DEG) {
            constexpr int off = K::COL[E0 + Q] * CS + K::DELTA[E0 + Q];
            if constexpr (K::EARLY[E0 + Q]) sv[Q] = pre[Q];
            else sv[Q] = softn[off];
            load_soft<J, Q + 1>(softn, sv, pre);
        }
    }
    template <int JN, int Q>
    static __device__ __forceinline__ void prefetch(const float* softn, float (&nxt)[NDEG<JN>])
    {
        if constexpr (JN < B) {
            if constexpr (Q < K::RP[JN + 1] - K::RP[JN]) {
                constexpr int off = K::COL[K::RP[JN] + Q] * CS + K::DELTA[K::RP[JN] + Q];
                if constexpr (K::EARLY[K::RP[JN] + Q]) nxt[Q] = softn[off];
                prefetch<JN, Q + 1>(softn, nxt);
            }
        }
    }

    // hbw = hb + warp, lane0: this lane stores the warp's packed words
    template <int J, int Q>
    static __device__ __forceinline__ unsigned new_message(const float (&v)[K::RP[J + 1] - K::RP[J]], float c1, unsigned m1x, unsigned m2x)
    {
        const bool ismin = fabsf(v[Q]) == c1;
        return (ismin ? m2x : m1x) ^ (__float_as_uint(v[Q]) & 0x80000000u);                      // decoders.cpp:5193-5198
    }
    template <int J, int Q>
    static __device__ __forceinline__ void put_posterior(float* softn, unsigned* hbw, bool lane0, bool active, float nv)
    {
        constexpr int E0 = K::RP[J];
        constexpr int off = K::COL[E0 + Q] * CS;
        if (ALL_ACTIVE || active) { softn[off] = nv; softn[off + Z] = nv; }                      // lane-aligned, both copies
        if constexpr (K::LAST[E0 + Q]) {
            // this block row is the last one of the iteration to touch the column: its values are the
            // iteration's posteriors, so their signs are the hard decisions the syndrome is taken of (:5281)
            const unsigned w = __ballot_sync(0xffffffffu, (ALL_ACTIVE || active) && nv < 0.0f);
            if (lane0) hbw[K::COL[E0 + Q] * HW] = w;
        }
    }
    template <int J, int Q>
    static __device__ __forceinline__ void phase2(float* softn, unsigned* hbw, bool lane0, bool active,
                                                  const float (&v)[K::RP[J + 1] - K::RP[J]], float c1,
                                                  unsigned m1x, unsigned m2x, unsigned (&msg)[K::RP[J + 1] - K::RP[J]])
    {
        constexpr int DEG = K::RP[J + 1] - K::RP[J];
        if constexpr (Q + 1 < DEG) {
            msg[Q] = new_message<J, Q>(v, c1, m1x, m2x);
            msg[Q + 1] = new_message<J, Q + 1>(v, c1, m1x, m2x);
            float n0, n1;
            add_f32x2(n0, n1, v[Q], v[Q + 1], __uint_as_float(msg[Q]), __uint_as_float(msg[Q + 1]));    // :5199-5204
            put_posterior<J, Q>(softn, hbw, lane0, active, n0);
            put_posterior<J, Q + 1>(softn, hbw, lane0, active, n1);
            phase2<J, Q + 2>(softn, hbw, lane0, active, v, c1, m1x, m2x, msg);
        } else if constexpr (Q < DEG) {
            msg[Q] = new_message<J, Q>(v, c1, m1x, m2x);
            put_posterior<J, Q>(softn, hbw, lane0, active, v[Q] + __uint_as_float(msg[Q]));
        }
    }

    // ---- split barrier between the loads and the stores of a layer.  A lane reads position n + DELTA of a
    // column and writes position n, so the word one lane reads is written by ANOTHER lane of the same layer;
    // within a warp program order protects it, across warps every warp announces "my loads are done"
    // (mbarrier arrive, one lane per warp) and checks that all warps have (try_wait) before its first store.
    // The loads come first in every warp, so the wait is practically never taken.
    static __device__ __forceinline__ void loads_done(unsigned mbar, bool lane0)
    {
        __syncwarp();
        if (lane0) asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" :: "r"(mbar) : "memory");
    }
    static __device__ __forceinline__ void wait_loads(unsigned mbar, unsigned parity)
    {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "WAIT_LOADS:\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                     "@!p bra WAIT_LOADS;\n\t}" :: "r"(mbar), "r"(parity) : "memory");
    }

    // two smallest magnitudes of v[LO..HI): balanced tree (depth log2 instead of the linear chain of S::reduce) so that
    // the few warps of an SM find independent instructions; min / max only, hence exact and order independent
    template <int DEG, int LO, int HI>
    static __device__ __forceinline__ typename S::RowAcc two_smallest(const float (&v)[DEG])
    {
        typename S::RowAcc r;
        if constexpr (HI - LO == 1) {
            r.c1 = fabsf(v[LO]); r.c2 = __int_as_float(0x7f800000);
        } else if constexpr (HI - LO == 2) {
            r.c1 = fminf(fabsf(v[LO]), fabsf(v[LO + 1])); r.c2 = fmaxf(fabsf(v[LO]), fabsf(v[LO + 1]));
        } else {
            constexpr int MID = LO + (((HI - LO) / 2 + 1) & ~1);         // even-sized left half
            const typename S::RowAcc a = two_smallest<DEG, LO, MID>(v), b = two_smallest<DEG, MID, HI>(v);
            r.c1 = fminf(a.c1, b.c1);
            r.c2 = fminf(fminf(a.c2, b.c2), fmaxf(a.c1, b.c1));
        }
        return r;
    }
    template <int DEG, int LO, int HI>
    static __device__ __forceinline__ unsigned sign_xor(const float (&v)[DEG])
    {
        if constexpr (HI - LO == 1) return __float_as_uint(v[LO]);
        else if constexpr (HI - LO == 2) return __float_as_uint(v[LO]) ^ __float_as_uint(v[LO + 1]);
        else if constexpr (HI - LO == 3) return __float_as_uint(v[LO]) ^ __float_as_uint(v[LO + 1]) ^ __float_as_uint(v[LO + 2]);
        else {
            constexpr int T = (HI - LO) / 3;
            return sign_xor<DEG, LO, LO + T>(v) ^ sign_xor<DEG, LO + T, LO + 2 * T>(v) ^ sign_xor<DEG, LO + 2 * T, HI>(v);
        }
    }

    // m[q] = min(cap, min over p != q of |v[p]|): groups of three edges, one three-input minimum per edge
    template <int DEG>
    static __device__ __forceinline__ void min_of_others(const float (&v)[DEG], float (&m)[DEG], const float CAP)
    {
        constexpr int G = (DEG + 2) / 3;
        float g[G], rest[G];
#pragma unroll
        for (int k = 0; k < G; k++) {
            const int a = 3 * k, b = a + 1 < DEG ? a + 1 : a, c = a + 2 < DEG ? a + 2 : a;
            g[k] = fminf(fminf(fabsf(v[a]), fabsf(v[b])), fabsf(v[c]));
        }
        if constexpr (G == 1) rest[0] = CAP;
        else if constexpr (G == 2) { rest[0] = fminf(g[1], CAP); rest[1] = fminf(g[0], CAP); }
        else if constexpr (G == 3) {
            rest[0] = fminf(fminf(g[1], g[2]), CAP); rest[1] = fminf(fminf(g[0], g[2]), CAP); rest[2] = fminf(fminf(g[0], g[1]), CAP);
        } else {
            float pre[G], suf[G];                                // minima of the groups before / after group k
            pre[0] = CAP; suf[G - 2] = g[G - 1];
#pragma unroll
            for (int k = 1; k < G; k++) pre[k] = fminf(pre[k - 1], g[k - 1]);
#pragma unroll
            for (int k = G - 3; k >= 0; k--) suf[k] = fminf(suf[k + 1], g[k + 1]);
#pragma unroll
            for (int k = 0; k < G - 1; k++) rest[k] = fminf(pre[k], suf[k]);
            rest[G - 1] = pre[G - 1];
        }
#pragma unroll
        for (int k = 0; k < G; k++) {
            const int a = 3 * k;
            if (a + 2 < DEG) {
                m[a] = fminf(fminf(fabsf(v[a + 1]), fabsf(v[a + 2])), rest[k]);
                m[a + 1] = fminf(fminf(fabsf(v[a]), fabsf(v[a + 2])), rest[k]);
                m[a + 2] = fminf(fminf(fabsf(v[a]), fabsf(v[a + 1])), rest[k]);
            } else if (a + 1 < DEG) {
                m[a] = fminf(fabsf(v[a + 1]), rest[k]);
                m[a + 1] = fminf(fabsf(v[a]), rest[k]);
            } else m[a] = rest[k];
        }
    }
    template <int J, int Q>
    static __device__ __forceinline__ void phase2o(float* softn, unsigned* hbw, bool lane0, bool active,
                                                   const float (&v)[K::RP[J + 1] - K::RP[J]], const unsigned (&msg)[K::RP[J + 1] - K::RP[J]])
    {
        constexpr int DEG = K::RP[J + 1] - K::RP[J];
        if constexpr (Q + 1 < DEG) {
            float n0, n1;
            add_f32x2(n0, n1, v[Q], v[Q + 1], __uint_as_float(msg[Q]), __uint_as_float(msg[Q + 1]));    // :5199-5204
            put_posterior<J, Q>(softn, hbw, lane0, active, n0);
            put_posterior<J, Q + 1>(softn, hbw, lane0, active, n1);
            phase2o<J, Q + 2>(softn, hbw, lane0, active, v, msg);
        } else if constexpr (Q < DEG) {
            put_posterior<J, Q>(softn, hbw, lane0, active, v[Q] + __uint_as_float(msg[Q]));
        }
    }

    template <int J>
    static __device__ __forceinline__ void layer(float* softn, unsigned* hbw, unsigned trow, unsigned mbar, unsigned& ph, bool lane0, bool active,
                                                 const float (&pre)[NDEG<J>], float (&nxt)[NDEG<J + 1>])
    {
        constexpr int E0 = K::RP[J], DEG = K::RP[J + 1] - K::RP[J];
        unsigned msg[DEG];
        float sv[DEG], v[DEG];
        tmem_ld_n<DEG>(trow + E0, msg);                                                          // old c2v of this row
        load_soft<J, 0>(softn, sv, pre);
        prefetch<J + 1, 0>(softn, nxt);                                                          // the next block row's untouched columns
        tmem_wait_ld<DEG>(msg);
#pragma unroll
        for (int q = 0; q + 1 < DEG; q += 2)                                                     // :5152-5158, two edges per FADD2
            sub_f32x2(v[q], v[q + 1], sv[q], sv[q + 1], __uint_as_float(msg[q]), __uint_as_float(msg[q + 1]));
        if constexpr (DEG & 1) v[DEG - 1] = sv[DEG - 1] - __uint_as_float(msg[DEG - 1]);
        loads_done(mbar, lane0);
#if LMS_TMEM_OTHERS
        // the message of edge q is f(min over the OTHER edges of |v2c|), f(x) = min(max(x - 0.4, 0), 32767) (:5166-5168,
        // :5131-5137): three-input minima over groups of three edges (FMNMX3, |.| is an operand modifier), one per edge,
        // instead of tracking the two smallest and selecting per edge.  The ceiling is folded into the per-group
        // "rest" term as CAP = 32767 + 205/512, the smallest float whose x - 0.4f rounds to 32767 or more (it rounds to
        // exactly 32767.0, and every smaller float to less; tests/test_message_forms.py).  Offset, floor and the row's sign product s = +-1 are two
        // FFMAs: th = m * (s/2) - 0.4f * (s/2) = s * (m - 0.4f) / 2 with the same single rounding as m - 0.4f (scaling
        // by 1/2 is exact), r = |th| * s + th = s * max(m - 0.4f, 0) exactly; the edge's own sign is one LOP3.  Same
        // posteriors and decisions bit for bit as the two-smallest form (tests/test_gpu_tmem.py); a zero message may
        // carry the other sign, which no later operation can see (x - (+-0) = x, and a posterior is never -0).
        // 4.1 instead of 6.9 instructions per edge on the half-rate ALU pipe, 11.8 instead of 12.8 in total.
        float m[DEG];
        const unsigned sacc = sign_xor<DEG, 0, DEG>(v) & 0x80000000u;
        const float rone = __uint_as_float(sacc | 0x3f800000u), rhalf = __fmul_rn(rone, 0.5f);
        const float nhalf = __fmul_rn(rhalf, -0.4f);                                             // exact: -(s/2) * 0.4f
        min_of_others<DEG>(v, m, 32767.400390625f);
#pragma unroll
        for (int q = 0; q < DEG; q++) {
            const float th = __fmaf_rn(m[q], rhalf, nhalf);
            msg[q] = __float_as_uint(__fmaf_rn(fabsf(th), rone, th)) ^ (__float_as_uint(v[q]) & 0x80000000u);
        }
        if constexpr (B % 2 == 0) wait_loads(mbar, J & 1);                                       // before the first store
        else { wait_loads(mbar, ph); ph ^= 1u; }
        phase2o<J, 0>(softn, hbw, lane0, active, v, msg);
#else
        const typename S::RowAcc a = two_smallest<DEG, 0, DEG>(v);
        const unsigned sacc = sign_xor<DEG, 0, DEG>(v);
        const float n1 = fminf(fmaxf(a.c1 - 0.4f, 0.0f), 32767.0f);                              // :5166-5168, :5131-5137
        const float n2 = fminf(fmaxf(a.c2 - 0.4f, 0.0f), 32767.0f);
        const unsigned rs = sacc & 0x80000000u;
        const unsigned m1x = __float_as_uint(n1) ^ rs, m2x = __float_as_uint(n2) ^ rs;
        if constexpr (B % 2 == 0) wait_loads(mbar, J & 1);
        else { wait_loads(mbar, ph); ph ^= 1u; }
        phase2<J, 0>(softn, hbw, lane0, active, v, a.c1, m1x, m2x, msg);
#endif
        tmem_st_n<DEG>(trow + E0, msg);                                                          // :5179
    }

    template <int J>
    static __device__ __forceinline__ void layers(float* softn, unsigned* hbw, unsigned trow, unsigned mbar, unsigned& ph, bool lane0, bool active,
                                                  const float (&pre)[NDEG<J>])
    {
        if constexpr (J < B) {
            float nxt[NDEG<J + 1>];
            layer<J>(softn, hbw, trow, mbar, ph, lane0, active, pre, nxt);
            __syncthreads();
            layers<J + 1>(softn, hbw, trow, mbar, ph, lane0, active, nxt);
        }
    }

    // ---- PP block rows.  Compile-time bookkeeping of the two buffers of a column
    static __host__ __device__ constexpr int touch_index(int e) { int t = 0; for (int i = 0; i < e; i++) t += K::COL[i] == K::COL[e] ? 1 : 0; return t; }
    static __host__ __device__ constexpr int col_weight(int col) { int t = 0; for (int i = 0; i < E; i++) t += K::COL[i] == col ? 1 : 0; return t; }
    // DELTA of the edge that reads what edge e writes (the next edge of the same column, cyclically)
    static __host__ __device__ constexpr int next_delta(int e)
    {
        for (int i = 1; i <= E; i++) if (K::COL[(e + i) % E] == K::COL[e]) return K::DELTA[(e + i) % E];
        return 0;
    }
    // the buffer edge e reads in an iteration of parity PAR (it writes the other one)
    template <int PAR> static __host__ __device__ constexpr int rbuf(int e) { return (PAR * col_weight(K::COL[e]) + touch_index(e)) & 1; }

    // loads of block row J: the edges whose column the previous block row does not write (WANT_EARLY, issued before the
    // barrier that ends the previous block row) or the others (after it).  softl = buffer 0 + lane, uoff[a] = 32 * ((warp + a) mod NWARPS)
    template <int J, int PAR, bool WANT_EARLY, int Q = 0>
    static __device__ __forceinline__ void pp_load(const float* softl, const unsigned (&uoff)[NWARPS], float (&sv)[NDEG<J>])
    {
        if constexpr (J < B) {
            if constexpr (Q < NDEG<J>) {
                constexpr int e = K::RP[J < B ? J : 0] + Q;
                if constexpr (K::EARLY[e] == WANT_EARLY) {
                    constexpr int a = (K::DELTA[e] / 32 + SEG(K::COL[e])) % NWARPS, b = K::DELTA[e] % 32;
                    constexpr int off = (rbuf<PAR>(e) * C + K::COL[e]) * CS + b;
                    sv[Q] = softl[uoff[a] + off];
                }
                pp_load<J, PAR, WANT_EARLY, Q + 1>(softl, uoff, sv);
            }
        }
    }
    template <int J, int PAR, int Q = 0>
    static __device__ __forceinline__ void pp_put(float* softn, const float* softl, const unsigned (&uoff)[NWARPS], unsigned* hbw, bool lane0, const float (&nv)[NDEG<J>])
    {
        if constexpr (Q < NDEG<J>) {
            constexpr int e = K::RP[J] + Q;
            constexpr int off = ((rbuf<PAR>(e) ^ 1) * C + K::COL[e]) * CS;
            if constexpr (SEG(K::COL[e]) == 0) softn[off] = nv[Q];                               // one word per lane
            else const_cast<float*>(softl)[uoff[SEG(K::COL[e])] + off] = nv[Q];
#if LMS_TMEM_PP_BAL
            // the repeat: the warp whose words are 0 .. 31 of this column; not needed when the next reader's DELTA is a multiple of 32
            if constexpr (next_delta(e) % 32 != 0) {
                if (uoff[SEG(K::COL[e])] == 0) const_cast<float*>(softl)[off + Z] = nv[Q];
            }
#endif
            if constexpr (K::LAST[e]) {                                                          // see put_posterior
                const unsigned w = __ballot_sync(0xffffffffu, nv[Q] < 0.0f);
                if (lane0) hbw[K::COL[e] * HW] = w;
            }
            pp_put<J, PAR, Q + 1>(softn, softl, uoff, hbw, lane0, nv);
        }
    }
    template <int J, int PAR, int Q = 0>
    static __device__ __forceinline__ void pp_repeat_plain(float* softn, const float (&nv)[NDEG<J>])
    {
        if constexpr (Q < NDEG<J>) {
            constexpr int e = K::RP[J] + Q;
            constexpr int off = ((rbuf<PAR>(e) ^ 1) * C + K::COL[e]) * CS + Z;
            softn[off] = nv[Q];
            pp_repeat_plain<J, PAR, Q + 1>(softn, nv);
        }
    }
    // warp 0 repeats its words at positions Z .. Z+31.  The stores sit behind a real branch (inline PTX: the compiler turns
    // an `if` around a few stores into predicated stores, which would cost the other warps an issue slot each)
    template <int J, int PAR, int Q>
    static constexpr int pp_rep_off = 4 * (((rbuf<PAR>(K::RP[J] + (Q < NDEG<J> ? Q : 0)) ^ 1) * C + K::COL[K::RP[J] + (Q < NDEG<J> ? Q : 0)]) * CS + Z);
    template <int J, int PAR, int Q = 0>
    static __device__ __forceinline__ void pp_repeat(unsigned softn_s, unsigned notwarp0, const float (&nv)[NDEG<J>])
    {
        // a loop that runs once for warp 0 and not at all for the others: the compiler keeps it a branch
#define REP_HEAD "{\n\t.reg .pred p;\n\t.reg .u32 i;\n\tmov.u32 i, %0;\n\tREP_LOOP:\n\tsetp.ne.u32 p, i, 0;\n\t@p bra.uni REP_DONE;\n\t"
#define REP_TAIL "add.u32 i, i, 1;\n\tbra.uni REP_LOOP;\n\tREP_DONE:\n\t}"
        constexpr int DEG = NDEG<J>;
        if constexpr (Q + 14 <= DEG) {
            asm volatile(REP_HEAD "st.shared.f32 [%1+%2], %16;\n\tst.shared.f32 [%1+%3], %17;\n\tst.shared.f32 [%1+%4], %18;\n\tst.shared.f32 [%1+%5], %19;\n\tst.shared.f32 [%1+%6], %20;\n\tst.shared.f32 [%1+%7], %21;\n\tst.shared.f32 [%1+%8], %22;\n\tst.shared.f32 [%1+%9], %23;\n\tst.shared.f32 [%1+%10], %24;\n\tst.shared.f32 [%1+%11], %25;\n\tst.shared.f32 [%1+%12], %26;\n\tst.shared.f32 [%1+%13], %27;\n\tst.shared.f32 [%1+%14], %28;\n\tst.shared.f32 [%1+%15], %29;\n\t" REP_TAIL
                         :: "r"(notwarp0), "r"(softn_s), "n"(pp_rep_off<J, PAR, Q + 0>), "n"(pp_rep_off<J, PAR, Q + 1>), "n"(pp_rep_off<J, PAR, Q + 2>), "n"(pp_rep_off<J, PAR, Q + 3>), "n"(pp_rep_off<J, PAR, Q + 4>), "n"(pp_rep_off<J, PAR, Q + 5>), "n"(pp_rep_off<J, PAR, Q + 6>), "n"(pp_rep_off<J, PAR, Q + 7>), "n"(pp_rep_off<J, PAR, Q + 8>), "n"(pp_rep_off<J, PAR, Q + 9>), "n"(pp_rep_off<J, PAR, Q + 10>), "n"(pp_rep_off<J, PAR, Q + 11>), "n"(pp_rep_off<J, PAR, Q + 12>), "n"(pp_rep_off<J, PAR, Q + 13>),
                            "f"(nv[Q + 0]), "f"(nv[Q + 1]), "f"(nv[Q + 2]), "f"(nv[Q + 3]), "f"(nv[Q + 4]), "f"(nv[Q + 5]), "f"(nv[Q + 6]), "f"(nv[Q + 7]), "f"(nv[Q + 8]), "f"(nv[Q + 9]), "f"(nv[Q + 10]), "f"(nv[Q + 11]), "f"(nv[Q + 12]), "f"(nv[Q + 13]) : "memory");
            pp_repeat<J, PAR, Q + 14>(softn_s, notwarp0, nv);
        }
        else if constexpr (Q + 13 <= DEG) {
            asm volatile(REP_HEAD "st.shared.f32 [%1+%2], %15;\n\tst.shared.f32 [%1+%3], %16;\n\tst.shared.f32 [%1+%4], %17;\n\tst.shared.f32 [%1+%5], %18;\n\tst.shared.f32 [%1+%6], %19;\n\tst.shared.f32 [%1+%7], %20;\n\tst.shared.f32 [%1+%8], %21;\n\tst.shared.f32 [%1+%9], %22;\n\tst.shared.f32 [%1+%10], %23;\n\tst.shared.f32 [%1+%11], %24;\n\tst.shared.f32 [%1+%12], %25;\n\tst.shared.f32 [%1+%13], %26;\n\tst.shared.f32 [%1+%14], %27;\n\t" REP_TAIL
                         :: "r"(notwarp0), "r"(softn_s), "n"(pp_rep_off<J, PAR, Q + 0>), "n"(pp_rep_off<J, PAR, Q + 1>), "n"(pp_rep_off<J, PAR, Q + 2>), "n"(pp_rep_off<J, PAR, Q + 3>), "n"(pp_rep_off<J, PAR, Q + 4>), "n"(pp_rep_off<J, PAR, Q + 5>), "n"(pp_rep_off<J, PAR, Q + 6>), "n"(pp_rep_off<J, PAR, Q + 7>), "n"(pp_rep_off<J, PAR, Q + 8>), "n"(pp_rep_off<J, PAR, Q + 9>), "n"(pp_rep_off<J, PAR, Q + 10>), "n"(pp_rep_off<J, PAR, Q + 11>), "n"(pp_rep_off<J, PAR, Q + 12>),
                            "f"(nv[Q + 0]), "f"(nv[Q + 1]), "f"(nv[Q + 2]), "f"(nv[Q + 3]), "f"(nv[Q + 4]), "f"(nv[Q + 5]), "f"(nv[Q + 6]), "f"(nv[Q + 7]), "f"(nv[Q + 8]), "f"(nv[Q + 9]), "f"(nv[Q + 10]), "f"(nv[Q + 11]), "f"(nv[Q + 12]) : "memory");
            pp_repeat<J, PAR, Q + 13>(softn_s, notwarp0, nv);
        }
        else if constexpr (Q + 12 <= DEG) {
            asm volatile(REP_HEAD "st.shared.f32 [%1+%2], %14;\n\tst.shared.f32 [%1+%3], %15;\n\tst.shared.f32 [%1+%4], %16;\n\tst.shared.f32 [%1+%5], %17;\n\tst.shared.f32 [%1+%6], %18;\n\tst.shared.f32 [%1+%7], %19;\n\tst.shared.f32 [%1+%8], %20;\n\tst.shared.f32 [%1+%9], %21;\n\tst.shared.f32 [%1+%10], %22;\n\tst.shared.f32 [%1+%11], %23;\n\tst.shared.f32 [%1+%12], %24;\n\tst.shared.f32 [%1+%13], %25;\n\t" REP_TAIL
                         :: "r"(notwarp0), "r"(softn_s), "n"(pp_rep_off<J, PAR, Q + 0>), "n"(pp_rep_off<J, PAR, Q + 1>), "n"(pp_rep_off<J, PAR, Q + 2>), "n"(pp_rep_off<J, PAR, Q + 3>), "n"(pp_rep_off<J, PAR, Q + 4>), "n"(pp_rep_off<J, PAR, Q + 5>), "n"(pp_rep_off<J, PAR, Q + 6>), "n"(pp_rep_off<J, PAR, Q + 7>), "n"(pp_rep_off<J, PAR, Q + 8>), "n"(pp_rep_off<J, PAR, Q + 9>), "n"(pp_rep_off<J, PAR, Q + 10>), "n"(pp_rep_off<J, PAR, Q + 11>),
                            "f"(nv[Q + 0]), "f"(nv[Q + 1]), "f"(nv[Q + 2]), "f"(nv[Q + 3]), "f"(nv[Q + 4]), "f"(nv[Q + 5]), "f"(nv[Q + 6]), "f"(nv[Q + 7]), "f"(nv[Q + 8]), "f"(nv[Q + 9]), "f"(nv[Q + 10]), "f"(nv[Q + 11]) : "memory");
            pp_repeat<J, PAR, Q + 12>(softn_s, notwarp0, nv);
        }
        else if constexpr (Q + 11 <= DEG) {
            asm volatile(REP_HEAD "st.shared.f32 [%1+%2], %13;\n\tst.shared.f32 [%1+%3], %14;\n\tst.shared.f32 [%1+%4], %15;\n\tst.shared.f32 [%1+%5], %16;\n\tst.shared.f32 [%1+%6], %17;\n\tst.shared.f32 [%1+%7], %18;\n\tst.shared.f32 [%1+%8], %19;\n\tst.shared.f32 [%1+%9], %20;\n\tst.shared.f32 [%1+%10], %21;\n\tst.shared.f32 [%1+%11], %22;\n\tst.shared.f32 [%1+%12], %23;\n\t" REP_TAIL
                         :: "r"(notwarp0), "r"(softn_s), "n"(pp_rep_off<J, PAR, Q + 0>), "n"(pp_rep_off<J, PAR, Q + 1>), "n"(pp_rep_off<J, PAR, Q + 2>), "n"(pp_rep_off<J, PAR, Q + 3>), "n"(pp_rep_off<J, PAR, Q + 4>), "n"(pp_rep_off<J, PAR, Q + 5>), "n"(pp_rep_off<J, PAR, Q + 6>), "n"(pp_rep_off<J, PAR, Q + 7>), "n"(pp_rep_off<J, PAR, Q + 8>), "n"(pp_rep_off<J, PAR, Q + 9>), "n"(pp_rep_off<J, PAR, Q + 10>),
                            "f"(nv[Q + 0]), "f"(nv[Q + 1]), "f"(nv[Q + 2]), "f"(nv[Q + 3]), "f"(nv[Q + 4]), "f"(nv[Q + 5]), "f"(nv[Q + 6]), "f"(nv[Q + 7]), "f"(nv[Q + 8]), "f"(nv[Q + 9]), "f"(nv[Q + 10]) : "memory");
            pp_repeat<J, PAR, Q + 11>(softn_s, notwarp0, nv);
        }
        else if constexpr (Q + 10 <= DEG) {
            asm volatile(REP_HEAD "st.shared.f32 [%1+%2], %12;\n\tst.shared.f32 [%1+%3], %13;\n\tst.shared.f32 [%1+%4], %14;\n\tst.shared.f32 [%1+%5], %15;\n\tst.shared.f32 [%1+%6], %16;\n\tst.shared.f32 [%1+%7], %17;\n\tst.shared.f32 [%1+%8], %18;\n\tst.shared.f32 [%1+%9], %19;\n\tst.shared.f32 [%1+%10], %20;\n\tst.shared.f32 [%1+%11], %21;\n\t" REP_TAIL
                         :: "r"(notwarp0), "r"(softn_s), "n"(pp_rep_off<J, PAR, Q + 0>), "n"(pp_rep_off<J, PAR, Q + 1>), "n"(pp_rep_off<J, PAR, Q + 2>), "n"(pp_rep_off<J, PAR, Q + 3>), "n"(pp_rep_off<J, PAR, Q + 4>), "n"(pp_rep_off<J, PAR, Q + 5>), "n"(pp_rep_off<J, PAR, Q + 6>), "n"(pp_rep_off<J, PAR, Q + 7>), "n"(pp_rep_off<J, PAR, Q + 8>), "n"(pp_rep_off<J, PAR, Q + 9>),
                            "f"(nv[Q + 0]), "f"(nv[Q + 1]), "f"(nv[Q + 2]), "f"(nv[Q + 3]), "f"(nv[Q + 4]), "f"(nv[Q + 5]), "f"(nv[Q + 6]), "f"(nv[Q + 7]), "f"(nv[Q + 8]), "f"(nv[Q + 9]) : "memory");
            pp_repeat<J, PAR, Q + 10>(softn_s, notwarp0, nv);
        }
        else if constexpr (Q + 9 <= DEG) {
            asm volatile(REP_HEAD "st.shared.f32 [%1+%2], %11;\n\tst.shared.f32 [%1+%3], %12;\n\tst.shared.f32 [%1+%4], %13;\n\tst.shared.f32 [%1+%5], %14;\n\tst.shared.f32 [%1+%6], %15;\n\tst.shared.f32 [%1+%7], %16;\n\tst.shared.f32 [%1+%8], %17;\n\tst.shared.f32 [%1+%9], %18;\n\tst.shared.f32 [%1+%10], %19;\n\t" REP_TAIL
                         :: "r"(notwarp0), "r"(softn_s), "n"(pp_rep_off<J, PAR, Q + 0>), "n"(pp_rep_off<J, PAR, Q + 1>), "n"(pp_rep_off<J, PAR, Q + 2>), "n"(pp_rep_off<J, PAR, Q + 3>), "n"(pp_rep_off<J, PAR, Q + 4>), "n"(pp_rep_off<J, PAR, Q + 5>), "n"(pp_rep_off<J, PAR, Q + 6>), "n"(pp_rep_off<J, PAR, Q + 7>), "n"(pp_rep_off<J, PAR, Q + 8>),
                            "f"(nv[Q + 0]), "f"(nv[Q + 1]), "f"(nv[Q + 2]), "f"(nv[Q + 3]), "f"(nv[Q + 4]), "f"(nv[Q + 5]), "f"(nv[Q + 6]), "f"(nv[Q + 7]), "f"(nv[Q + 8]) : "memory");
            pp_repeat<J, PAR, Q + 9>(softn_s, notwarp0, nv);
        }
        else if constexpr (Q + 8 <= DEG) {
            asm volatile(REP_HEAD "st.shared.f32 [%1+%2], %10;\n\tst.shared.f32 [%1+%3], %11;\n\tst.shared.f32 [%1+%4], %12;\n\tst.shared.f32 [%1+%5], %13;\n\tst.shared.f32 [%1+%6], %14;\n\tst.shared.f32 [%1+%7], %15;\n\tst.shared.f32 [%1+%8], %16;\n\tst.shared.f32 [%1+%9], %17;\n\t" REP_TAIL
                         :: "r"(notwarp0), "r"(softn_s), "n"(pp_rep_off<J, PAR, Q + 0>), "n"(pp_rep_off<J, PAR, Q + 1>), "n"(pp_rep_off<J, PAR, Q + 2>), "n"(pp_rep_off<J, PAR, Q + 3>), "n"(pp_rep_off<J, PAR, Q + 4>), "n"(pp_rep_off<J, PAR, Q + 5>), "n"(pp_rep_off<J, PAR, Q + 6>), "n"(pp_rep_off<J, PAR, Q + 7>),
                            "f"(nv[Q + 0]), "f"(nv[Q + 1]), "f"(nv[Q + 2]), "f"(nv[Q + 3]), "f"(nv[Q + 4]), "f"(nv[Q + 5]), "f"(nv[Q + 6]), "f"(nv[Q + 7]) : "memory");
            pp_repeat<J, PAR, Q + 8>(softn_s, notwarp0, nv);
        }
        else if constexpr (Q + 7 <= DEG) {
            asm volatile(REP_HEAD "st.shared.f32 [%1+%2], %9;\n\tst.shared.f32 [%1+%3], %10;\n\tst.shared.f32 [%1+%4], %11;\n\tst.shared.f32 [%1+%5], %12;\n\tst.shared.f32 [%1+%6], %13;\n\tst.shared.f32 [%1+%7], %14;\n\tst.shared.f32 [%1+%8], %15;\n\t" REP_TAIL
                         :: "r"(notwarp0), "r"(softn_s), "n"(pp_rep_off<J, PAR, Q + 0>), "n"(pp_rep_off<J, PAR, Q + 1>), "n"(pp_rep_off<J, PAR, Q + 2>), "n"(pp_rep_off<J, PAR, Q + 3>), "n"(pp_rep_off<J, PAR, Q + 4>), "n"(pp_rep_off<J, PAR, Q + 5>), "n"(pp_rep_off<J, PAR, Q + 6>),
                            "f"(nv[Q + 0]), "f"(nv[Q + 1]), "f"(nv[Q + 2]), "f"(nv[Q + 3]), "f"(nv[Q + 4]), "f"(nv[Q + 5]), "f"(nv[Q + 6]) : "memory");
            pp_repeat<J, PAR, Q + 7>(softn_s, notwarp0, nv);
        }
        else if constexpr (Q + 6 <= DEG) {
            asm volatile(REP_HEAD "st.shared.f32 [%1+%2], %8;\n\tst.shared.f32 [%1+%3], %9;\n\tst.shared.f32 [%1+%4], %10;\n\tst.shared.f32 [%1+%5], %11;\n\tst.shared.f32 [%1+%6], %12;\n\tst.shared.f32 [%1+%7], %13;\n\t" REP_TAIL
                         :: "r"(notwarp0), "r"(softn_s), "n"(pp_rep_off<J, PAR, Q + 0>), "n"(pp_rep_off<J, PAR, Q + 1>), "n"(pp_rep_off<J, PAR, Q + 2>), "n"(pp_rep_off<J, PAR, Q + 3>), "n"(pp_rep_off<J, PAR, Q + 4>), "n"(pp_rep_off<J, PAR, Q + 5>),
                            "f"(nv[Q + 0]), "f"(nv[Q + 1]), "f"(nv[Q + 2]), "f"(nv[Q + 3]), "f"(nv[Q + 4]), "f"(nv[Q + 5]) : "memory");
            pp_repeat<J, PAR, Q + 6>(softn_s, notwarp0, nv);
        }
        else if constexpr (Q + 5 <= DEG) {
            asm volatile(REP_HEAD "st.shared.f32 [%1+%2], %7;\n\tst.shared.f32 [%1+%3], %8;\n\tst.shared.f32 [%1+%4], %9;\n\tst.shared.f32 [%1+%5], %10;\n\tst.shared.f32 [%1+%6], %11;\n\t" REP_TAIL
                         :: "r"(notwarp0), "r"(softn_s), "n"(pp_rep_off<J, PAR, Q + 0>), "n"(pp_rep_off<J, PAR, Q + 1>), "n"(pp_rep_off<J, PAR, Q + 2>), "n"(pp_rep_off<J, PAR, Q + 3>), "n"(pp_rep_off<J, PAR, Q + 4>),
                            "f"(nv[Q + 0]), "f"(nv[Q + 1]), "f"(nv[Q + 2]), "f"(nv[Q + 3]), "f"(nv[Q + 4]) : "memory");
            pp_repeat<J, PAR, Q + 5>(softn_s, notwarp0, nv);
        }
        else if constexpr (Q + 4 <= DEG) {
            asm volatile(REP_HEAD "st.shared.f32 [%1+%2], %6;\n\tst.shared.f32 [%1+%3], %7;\n\tst.shared.f32 [%1+%4], %8;\n\tst.shared.f32 [%1+%5], %9;\n\t" REP_TAIL
                         :: "r"(notwarp0), "r"(softn_s), "n"(pp_rep_off<J, PAR, Q + 0>), "n"(pp_rep_off<J, PAR, Q + 1>), "n"(pp_rep_off<J, PAR, Q + 2>), "n"(pp_rep_off<J, PAR, Q + 3>),
                            "f"(nv[Q + 0]), "f"(nv[Q + 1]), "f"(nv[Q + 2]), "f"(nv[Q + 3]) : "memory");
            pp_repeat<J, PAR, Q + 4>(softn_s, notwarp0, nv);
        }
        else if constexpr (Q + 3 <= DEG) {
            asm volatile(REP_HEAD "st.shared.f32 [%1+%2], %5;\n\tst.shared.f32 [%1+%3], %6;\n\tst.shared.f32 [%1+%4], %7;\n\t" REP_TAIL
                         :: "r"(notwarp0), "r"(softn_s), "n"(pp_rep_off<J, PAR, Q + 0>), "n"(pp_rep_off<J, PAR, Q + 1>), "n"(pp_rep_off<J, PAR, Q + 2>),
                            "f"(nv[Q + 0]), "f"(nv[Q + 1]), "f"(nv[Q + 2]) : "memory");
            pp_repeat<J, PAR, Q + 3>(softn_s, notwarp0, nv);
        }
        else if constexpr (Q + 2 <= DEG) {
            asm volatile(REP_HEAD "st.shared.f32 [%1+%2], %4;\n\tst.shared.f32 [%1+%3], %5;\n\t" REP_TAIL
                         :: "r"(notwarp0), "r"(softn_s), "n"(pp_rep_off<J, PAR, Q + 0>), "n"(pp_rep_off<J, PAR, Q + 1>),
                            "f"(nv[Q + 0]), "f"(nv[Q + 1]) : "memory");
            pp_repeat<J, PAR, Q + 2>(softn_s, notwarp0, nv);
        }
        else if constexpr (Q + 1 <= DEG) {
            asm volatile(REP_HEAD "st.shared.f32 [%1+%2], %3;\n\t" REP_TAIL
                         :: "r"(notwarp0), "r"(softn_s), "n"(pp_rep_off<J, PAR, Q + 0>),
                            "f"(nv[Q + 0]) : "memory");
            pp_repeat<J, PAR, Q + 1>(softn_s, notwarp0, nv);
        }
#undef REP_HEAD
#undef REP_TAIL
    }

    // msg / sv: the row's old messages (tcgen05.ld in flight) and the early posteriors, both issued by the previous block row;
    // msgn / svn: the same for the next one
    template <int J, int PAR, bool LOAD = true>
    static __device__ __forceinline__ void pp_layer(float* softn, const float* softl, const unsigned (&uoff)[NWARPS], unsigned* hbw, unsigned trow,
                                                    bool lane0, bool warp0, unsigned (&msg)[NDEG<J>], float (&sv)[NDEG<J>],
                                                    unsigned (&msgn)[NDEG<J + 1>], float (&svn)[NDEG<J + 1>])
    {
        constexpr int E0 = K::RP[J], DEG = NDEG<J>;
        float v[DEG], m[DEG], nv[DEG];
        if constexpr (LOAD) pp_load<J, PAR, false>(softl, uoff, sv);
        tmem_wait_ld<DEG>(msg);
#pragma unroll
        for (int q = 0; q + 1 < DEG; q += 2)                                                     // :5152-5158, two edges per FADD2
            sub_f32x2(v[q], v[q + 1], sv[q], sv[q + 1], __uint_as_float(msg[q]), __uint_as_float(msg[q + 1]));
        if constexpr (DEG & 1) v[DEG - 1] = sv[DEG - 1] - __uint_as_float(msg[DEG - 1]);
        const unsigned sacc = sign_xor<DEG, 0, DEG>(v) & 0x80000000u;                            // see layer()
        const float rone = __uint_as_float(sacc | 0x3f800000u), rhalf = __fmul_rn(rone, 0.5f);
        const float nhalf = __fmul_rn(rhalf, -0.4f);
#if LMS_TMEM_WHATIF == 2
#pragma unroll
        for (int q = 0; q < DEG; q++) m[q] = fabsf(v[(q + 1) % DEG]);
#else
        min_of_others<DEG>(v, m, 32767.400390625f);
#endif
#if LMS_TMEM_FFMA2
        {
            float th[DEG];
#pragma unroll
            for (int q = 0; q + 1 < DEG; q += 2) fma_f32x2(th[q], th[q + 1], m[q], m[q + 1], rhalf, rhalf, nhalf, nhalf);
            if constexpr (DEG & 1) th[DEG - 1] = __fmaf_rn(m[DEG - 1], rhalf, nhalf);
#pragma unroll
            for (int q = 0; q < DEG; q++)
                msg[q] = __float_as_uint(__fmaf_rn(fabsf(th[q]), rone, th[q])) ^ (__float_as_uint(v[q]) & 0x80000000u);
        }
#else
#pragma unroll
        for (int q = 0; q < DEG; q++) {
            const float th = __fmaf_rn(m[q], rhalf, nhalf);
            msg[q] = __float_as_uint(__fmaf_rn(fabsf(th), rone, th)) ^ (__float_as_uint(v[q]) & 0x80000000u);
        }
#endif
#pragma unroll
        for (int q = 0; q + 1 < DEG; q += 2)                                                     // :5199-5204
            add_f32x2(nv[q], nv[q + 1], v[q], v[q + 1], __uint_as_float(msg[q]), __uint_as_float(msg[q + 1]));
        if constexpr (DEG & 1) nv[DEG - 1] = v[DEG - 1] + __uint_as_float(msg[DEG - 1]);
#if LMS_TMEM_WHATIF == 3
        softn[0] = nv[0];
#else
        pp_put<J, PAR>(softn, softl, uoff, hbw, lane0, nv);
#endif
#if LMS_TMEM_WHATIF == 3 || LMS_TMEM_WHATIF == 4 || LMS_TMEM_PP_BAL
#elif LMS_TMEM_PP_BRANCH
        pp_repeat<J, PAR>((unsigned)__cvta_generic_to_shared(softn), warp0 ? 0u : 1u, nv);
#else
        if (warp0) pp_repeat_plain<J, PAR>(softn, nv);
#endif
        tmem_st_n<DEG>(trow + E0, msg);                                                          // :5179
        if constexpr (J + 1 < B) {
            constexpr int E1 = K::RP[J + 1 < B ? J + 1 : 0];
            tmem_ld_n<NDEG<J + 1>>(trow + E1, msgn);
            pp_load<J + 1, PAR, true>(softl, uoff, svn);
        }
    }
    template <int J, int PAR>
    static __device__ __forceinline__ void pp_layers(float* softn, const float* softl, const unsigned (&uoff)[NWARPS], unsigned* hbw, unsigned trow,
                                                     bool lane0, bool warp0, unsigned (&msg)[NDEG<J>], float (&sv)[NDEG<J>])
    {
        if constexpr (J < B) {
            unsigned msgn[NDEG<J + 1>];
            float svn[NDEG<J + 1>];
            pp_layer<J, PAR>(softn, softl, uoff, hbw, trow, lane0, warp0, msg, sv, msgn, svn);
#if LMS_TMEM_WHATIF != 1
            __syncthreads();
#endif
            pp_layers<J + 1, PAR>(softn, softl, uoff, hbw, trow, lane0, warp0, msgn, svn);
        }
    }
    // One PP iteration.  The syndrome check of the decisions `hb_prev` (the previous iteration's, or the channel's) has
    // not been made yet: its quick look (the full pass where there is no quick look) runs inside block row 0 and its
    // CTA-wide OR rides on the barrier that ends block row 0, so the check costs its instructions and nothing else.
    // Block row 0 is therefore speculative: it writes the OTHER buffer of its columns, the other set of packed decisions
    // (hbw_next) and its messages; when verdict() says "stop" the state before it is still complete (the caller
    // does not count the iteration).  verdict(unsatisfied) -> true: stop here.
    template <int PAR, class F>
    static __device__ __forceinline__ bool pp_iteration(float* softn, const float* softl, const unsigned (&uoff)[NWARPS], const unsigned* hb_prev,
                                                        unsigned* hbw_next, const unsigned* plan, unsigned trow, bool lane0, bool warp0, int tid, F&& verdict)
    {
        unsigned msg[NDEG<0>], msgn[NDEG<1>];
        float sv[NDEG<0>], svn[NDEG<1>];
        tmem_wait_st();                                                                          // last iteration's messages are in place
        tmem_ld_n<NDEG<0>>(trow, msg);
        pp_load<0, PAR, false>(softl, uoff, sv);
        const unsigned part = SYN_QUICK ? quick_partial(hb_prev, plan, tid) : full_partial(hb_prev, plan, tid);
        pp_layer<0, PAR, false>(softn, softl, uoff, hbw_next, trow, lane0, warp0, msg, sv, msgn, svn);
        int bad = __syncthreads_or(part != 0);
        if constexpr (SYN_QUICK) {
            if (!bad) bad = __syncthreads_or(full_partial(hb_prev, plan, tid) != 0);             // rare: a (nearly) converged frame
        }
        if (verdict(bad)) { tmem_wait_ld<NDEG<1>>(msgn); return true; }
        pp_layers<1, PAR>(softn, softl, uoff, hbw_next, trow, lane0, warp0, msgn, svn);
        return false;
    }

    // ---- syndrome of the hard decisions on packed bits.  hb[col * HW + w] = signs of positions 32w..32w+31 of
    // column col (the column's rotated order).  During the iterations the words are written by the layers
    // themselves (phase2, K::LAST edges); pack() builds them for the channel values before the first iteration.
    static __device__ __forceinline__ void pack(const float* soft2, unsigned* hb, int tid)
    {
        const int lane = tid & 31, warp = tid >> 5;
#pragma unroll 8
        for (int col = 0; col < C; col++) {
            const int bit = (ALL_ACTIVE || tid < Z) ? soft2[col * CS + phys(col, tid)] < 0.0f : 0;
            const unsigned w = __ballot_sync(0xffffffffu, bit);
            if (lane == 0) hb[col * HW + warp] = w;
        }
        __syncthreads();
    }

    // One task = 32 check rows (block row j, word w): XOR over the row's edges of the 32-bit window of the
    // edge's column that starts at bit 32w + SYNSH.  Four lanes share a task (edges q, q+4, ...), 8 tasks per warp,
    // SYN_ROUNDS rounds.  What a thread has to fetch for each of its (task, edge) pairs never changes, so it is
    // worked out once per CTA (build_plan) into one word per pair:
    //   bits 0-9 column word offset in hb | 10-14 first word | 15-19 second word | 20-24 bit shift | 25-30 valid bits
    // (a pair that does not exist points at the always-zero word hb[HB_WORDS - 1]).
    static constexpr int SYN_NT = B * NB, SYN_STEP = 8 * NWARPS, SYN_ROUNDS = (SYN_NT + SYN_STEP - 1) / SYN_STEP;
    static constexpr int SYN_QE = (K::MAXDEG + 3) / 4;

    static __device__ __forceinline__ void build_plan(unsigned* plan, int tid)
    {
        const int lane = tid & 31, warp = tid >> 5, q = lane & 3;
        for (int r = 0; r < SYN_ROUNDS; r++) {
            const int t = r * SYN_STEP + warp * 8 + (lane >> 2);
            for (int i = 0; i < SYN_QE; i++) {
                unsigned pk = (unsigned)(HB_WORDS - 1) | (32u << 25);
                if (t < SYN_NT) {
                    const int j = t / NB, w = t - j * NB;
                    const int e = K::rt_rp()[j] + q + 4 * i;
                    if (e < K::rt_rp()[j + 1]) {
                        int start = 32 * w + K::rt_synsh()[e];
                        if (start >= Z) start -= Z;
                        const int i0 = start >> 5, i1 = i0 + 1 < HW ? i0 + 1 : (Z % 32 == 0 ? 0 : HW - 1);
                        const int nvalid = Z - start < 32 ? Z - start : 32;
                        pk = (unsigned)(K::rt_col()[e] * HW) | ((unsigned)i0 << 10) | ((unsigned)i1 << 15) | ((unsigned)(start & 31) << 20)
                             | ((unsigned)nvalid << 25);
                    }
                }
                plan[(r * SYN_QE + i) * ZP + tid] = pk;
            }
        }
        {   // quick look: the first 2 * NWARPS tasks, 16 lanes each, one (task, edge) pair per lane
            const int t = warp * 2 + (lane >> 4), q16 = lane & 15;
            unsigned pk = (unsigned)(HB_WORDS - 1) | (32u << 25);
            if (t < SYN_NT) {
                const int j = t / NB, w = t - j * NB;
                const int e = K::rt_rp()[j] + q16;
                if (e < K::rt_rp()[j + 1]) {
                    int start = 32 * w + K::rt_synsh()[e];
                    if (start >= Z) start -= Z;
                    const int i0 = start >> 5, i1 = i0 + 1 < HW ? i0 + 1 : (Z % 32 == 0 ? 0 : HW - 1);
                    const int nvalid = Z - start < 32 ? Z - start : 32;
                    pk = (unsigned)(K::rt_col()[e] * HW) | ((unsigned)i0 << 10) | ((unsigned)i1 << 15) | ((unsigned)(start & 31) << 20)
                         | ((unsigned)nvalid << 25);
                }
            }
            plan[PLAN_WORDS + tid] = pk;
        }
    }

    static __device__ __forceinline__ unsigned window(const unsigned* hb, unsigned pk)
    {
        const unsigned* hc = hb + (pk & 1023u);
        unsigned win = __funnelshift_r(hc[(pk >> 10) & 31u], hc[(pk >> 15) & 31u], pk >> 20);     // shift uses the low 5 bits
        if constexpr (Z % 32 != 0) {
            const unsigned nvalid = (pk >> 25) & 63u;       // the window runs over the end of the column: the rest wraps to bit 0
            if (nvalid < 32u) win = (win & ((1u << nvalid) - 1u)) | (hc[0] << nvalid);
        }
        return win;
    }

    // A cheap look at the first 2 * NWARPS tasks (a sixteenth to an eighth of the check rows): before a frame has
    // converged nearly every look finds an unsatisfied check and the full pass is skipped.  The verdict "non-zero" is
    // exact; "zero" only means that syndrome() has to decide.
    static constexpr bool SYN_QUICK = K::MAXDEG <= 16 && SYN_NT > 2 * NWARPS;
    // this thread's share of the quick look / of the full pass (non-zero: an unsatisfied check); the CTA-wide OR is the caller's
    static __device__ __forceinline__ unsigned quick_partial(const unsigned* hb, const unsigned* plan, int tid)
    {
        unsigned acc = window(hb, plan[PLAN_WORDS + tid]);
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) acc ^= __shfl_xor_sync(0xffffffffu, acc, o);
        if constexpr (Z % 32 != 0) {
            const int t = (tid >> 5) * 2 + ((tid & 31) >> 4);
            const int lanes = Z - 32 * (t % NB);
            if (lanes < 32) acc &= (1u << lanes) - 1u;
        }
        return acc;
    }
    static __device__ __forceinline__ unsigned full_partial(const unsigned* hb, const unsigned* plan, int tid)
    {
        unsigned bad = 0;
#pragma unroll
        for (int r = 0; r < SYN_ROUNDS; r++) {
            unsigned acc = 0;
#pragma unroll
            for (int i = 0; i < SYN_QE; i++) {
                acc ^= window(hb, plan[(r * SYN_QE + i) * ZP + tid]);
            }
            acc ^= __shfl_xor_sync(0xffffffffu, acc, 1);
            acc ^= __shfl_xor_sync(0xffffffffu, acc, 2);
            if constexpr (Z % 32 != 0) {
                const int t = r * SYN_STEP + (tid >> 5) * 8 + ((tid & 31) >> 2);
                const int lanes = Z - 32 * (t % NB);
                if (lanes < 32) acc &= (1u << lanes) - 1u;
            }
            bad |= acc;
        }
        return bad;
    }
    static __device__ __forceinline__ int syndrome_quick(const unsigned* hb, const unsigned* plan, int tid)
    {
        return __syncthreads_or(quick_partial(hb, plan, tid) != 0);
    }

    static __device__ __forceinline__ int syndrome(const unsigned* hb, const unsigned* plan, int tid)
    {
        if constexpr (SYN_QUICK) {
            if (syndrome_quick(hb, plan, tid)) return 1;
        }
        return __syncthreads_or(full_partial(hb, plan, tid) != 0);
    }

    // position of bit k of block column col in the doubled column (first copy)
    static __device__ __forceinline__ int pos_of(int col, int k)
    {
        int p = k + K::rt_ri()[col];
        return p >= Z ? p - Z : p;
    }
    static __device__ __forceinline__ void put(float* soft2, int col, int k, float x)
    {
        const int q = phys(col, pos_of(col, k)), p = col * CS + q;
        soft2[p] = x;
        if (!PP || q < 32) soft2[p + Z] = x;
    }

    static __device__ __forceinline__ void kernel(const FrameIO& io)
    {
        extern __shared__ __align__(16) float soft2[];
        unsigned* hb = (unsigned*)(soft2 + SOFT_WORDS);
        unsigned* plan = (unsigned*)(soft2 + PLAN_OFF);
        int* s_misc = (int*)(soft2 + MISC_OFF);
        const int tid = threadIdx.x;
        const bool active = tid < Z;
        const bool lane0 = (tid & 31) == 0;
        const bool noexit = io.flags & 8u;                       // LDPCB200_NO_EARLY_EXIT
        float* softn = soft2 + tid;
        unsigned* hbw = hb + (tid >> 5);
        const unsigned mbar = (unsigned)__cvta_generic_to_shared(soft2 + MBAR_OFF);
        unsigned ph = 0;
        // PP: warp-uniform pieces of the read addresses (the shuffle tells the compiler so)
        const unsigned wu = __shfl_sync(0xffffffffu, (unsigned)tid >> 5, 0);
        const bool warp0 = wu == 0;
        const float* softl = soft2 + (tid & 31);
        unsigned uoff[NWARPS];
#pragma unroll
        for (int a = 0; a < NWARPS; a++) uoff[a] = ((wu + a) % NWARPS) * 32u;
        int* s_cw = (int*)(soft2 + CW_OFF);
        if constexpr (PP) {
            for (int col = tid; col < C; col += ZP) {
                int wgt = 0;
                for (int e = 0; e < E; e++) wgt += K::rt_col()[e] == col;
                s_cw[col] = wgt;
            }
        }
        build_plan(plan, tid);
        if (tid == 0) { hb[HB_WORDS - 1] = 0u; if (PP) hb[2 * HB_WORDS - 1] = 0u; }
        if (tid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar), "r"((unsigned)NWARPS) : "memory");

        // tensor memory: one warp allocates TCOLS columns for the CTA and frees them at the end
        if (tid < 32) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                         :: "r"((unsigned)__cvta_generic_to_shared(hb)), "r"((unsigned)TCOLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const unsigned tbase = *(volatile unsigned*)hb;
        // this thread's lane (bits 31:16) and first column (bits 15:0)
        // (the shuffle tells the compiler that the address is warp-uniform: LDTM / STTM take it from a uniform register)
        const unsigned trow = __shfl_sync(0xffffffffu, tbase + ((unsigned)(((tid >> 5) & 3) * 32) << 16) + (unsigned)((tid >> 7) * E), 0);

        // Frame tickets are drawn two frames ahead (thread 0 keeps them in registers): the round trip of the atomic is off the
        // path between two frames, and the NEXT frame's index is known early enough to pull its LLRs into L2 while this one
        // is decoded (buffer mode; 128 bytes per prefetch).
        int t_cur = 0, t_next = 0;
        if (tid == 0) { t_cur = (int)atomicAdd(io.next_frame, 1u); t_next = (int)atomicAdd(io.next_frame, 1u); }
        for (;;) {
            __syncthreads();
            if (tid == 0) {
                s_misc[0] = t_cur; s_misc[3] = t_next; s_misc[1] = 0; s_misc[2] = 0;
                t_cur = t_next;
                t_next = (int)atomicAdd(io.next_frame, 1u);
            }
            __syncthreads();
            const int f = s_misc[0];
            if (f >= io.nf) break;
            if (!io.ch.enabled) {
                const int fn = s_misc[3];
                if (fn < io.nf) {
                    const size_t fbytes = (size_t)N * (io.llr_dtype == 1 ? 4 : 8);
                    const char* nb = (const char*)io.llr + (size_t)fn * fbytes;
                    for (size_t o = (size_t)tid * 128; o < fbytes; o += (size_t)ZP * 128) asm volatile("prefetch.global.L2 [%0];" :: "l"(nb + o));
                }
            }

            bool packed = false;                                 // hb already holds the decisions of the channel values
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
                                put(soft2, col, k, i >= io.ch.punct_start ? io.ch.punct_value : o4[4 * q + b]);
                            }
                        }
                    }
                    for (int c = (ncomp & ~3) + tid; c < ncomp; c += ZP) {
                        float o[4];
                        channel_llr_qam_component(io.ch, frame, c, o);
                        const int i0 = (c >> 1) * io.ch.m + (c & 1) * half;
                        for (int b = 0; b < half; b++) {
                            const int i = channel_dest(io.ch, i0 + b), col = i / Z, k = i - col * Z;
                            put(soft2, col, k, i >= io.ch.punct_start ? io.ch.punct_value : o[b]);
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
                            put(soft2, col, k, o[b]);
                        }
                    }
                    for (int j = (N & ~3) + tid; j < N; j += ZP) {                    // tail when N is not a multiple of 4
                        const int i = channel_dest(io.ch, j), col = i / Z, k = i - col * Z;
                        put(soft2, col, k, channel_llr(io.ch, frame, i));
                    }
                }
            } else if (io.llr_dtype == 1) {                      // LDPCB200_F32
                const float* y = (const float*)io.llr + (size_t)f * N;
                const bool act = ALL_ACTIVE || active;
#pragma unroll
                for (int c0 = 0; c0 < C; c0 += 16) {             // 16 loads in flight per thread
                    float x[16];
#pragma unroll
                    for (int u = 0; u < 16; u++) {               // position tid of column col holds bit (tid + ROT) mod Z
                        const int col = c0 + u;
                        if (col < C) {
                            int k = tid + K::rt_rot()[col];
                            if (k >= Z) k -= Z;
                            x[u] = act ? __ldcs(y + col * Z + k) : 0.0f;
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 16; u++) {
                        const int col = c0 + u;
                        if (col < C) {
                            if constexpr (PP) {
                                const int sw = seg_words(wu, col);
                                soft2[col * CS + sw + (tid & 31)] = x[u];
                                if (sw == 0) soft2[col * CS + Z + (tid & 31)] = x[u];
                            } else if (act) { softn[col * CS] = x[u]; softn[col * CS + Z] = x[u]; }
                            const unsigned w = __ballot_sync(0xffffffffu, act && x[u] < 0.0f);   // the packed decisions of the
                            if (lane0) hbw[col * HW] = w;                                         // channel values, as pack() builds them
                        }
                    }
                }
                packed = true;
            } else {
                const double* y = (const double*)io.llr + (size_t)f * N;
                if (ALL_ACTIVE || active) {
#pragma unroll 8
                    for (int col = 0; col < C; col++) {
                        int k = tid + K::rt_rot()[col];
                        if (k >= Z) k -= Z;
                        const float x = (float)__ldcs(y + col * Z + k);
                        if constexpr (PP) {
                            const int sw = seg_words(wu, col);
                            soft2[col * CS + sw + (tid & 31)] = x;
                            if (sw == 0) soft2[col * CS + Z + (tid & 31)] = x;
                        } else { softn[col * CS] = x; softn[col * CS + Z] = x; }
                    }
                }
            }
            tmem_zero_n<E>(trow);                                                       // prev[] = 0, decoders.cpp:5088-5108
            tmem_wait_st();
            __syncthreads();

            if (!packed) pack(soft2, hb, tid);
            int ret = 0, locked = 0, iter = 0, done = 0, parity = 1, hsel = 0;
            if constexpr (PP) {
                // The check of the decisions after `iter` complete iterations (the channel values' for 0, :5111-5115; :5281-5284)
                // is made inside block row 0 of the next iteration (pp_iteration); only the last one stands alone.  With
                // LDPCB200_NO_EARLY_EXIT every iteration and every check runs (worst-case timing); `ret` keeps the count at
                // which the reference would have stopped.
                auto verdict = [&](int bad) -> bool {                                   // -> stop
                    if (!locked) { parity = bad; if (!bad) { ret = iter == 0 ? 1 : iter; locked = 1; } }
                    return !bad && !noexit;                                             // :5119
                };
                for (;;) {
                    const unsigned* hb_prev = hb + hsel * HB_WORDS;
                    unsigned* hbw_next = hbw + (hsel ^ 1) * HB_WORDS;
                    if (iter >= io.maxiter) { verdict(syndrome(hb_prev, plan, tid)); break; }
                    const bool stop = (iter & 1) ? pp_iteration<1>(softn, softl, uoff, hb_prev, hbw_next, plan, trow, lane0, warp0, tid, verdict)
                                                 : pp_iteration<0>(softn, softl, uoff, hb_prev, hbw_next, plan, trow, lane0, warp0, tid, verdict);
                    if (stop) break;
                    iter++; done++; hsel ^= 1;
                }
                if (!locked) ret = -iter;                                               // :5424
            } else {
                parity = syndrome(hb, plan, tid);                                       // :5111-5115
                if (!parity) { ret = 1; locked = 1; }
                for (iter = 0; iter < io.maxiter; iter++) {
                    if (!parity && !noexit) break;                                      // :5119
                    tmem_wait_st();                                                     // last iteration's messages are in place
                    { float none[NDEG<0>]; layers<0>(softn, hbw, trow, mbar, ph, lane0, active, none); }
                    done++;
                    const int par = syndrome(hb, plan, tid);                            // :5281-5284
                    if (!locked) { parity = par; if (!par) { ret = iter + 1; locked = 1; } }
                    if (!par && !noexit) break;
                }
                if (!locked) ret = parity ? -iter : iter + 1;                           // :5424
            }
            const unsigned* hbf = hb + hsel * HB_WORDS;                                 // the packed decisions of the result

            if (io.post) {
                // PP: the column's result is in buffer (iterations run * column weight) mod 2, single copy
                auto at = [&](int col) -> float {
                    if constexpr (PP) return soft2[((done * s_cw[col]) & 1) * BUF + col * CS + phys(col, pos_of(col, tid))];
                    else return soft2[col * CS + tid + K::rt_ri()[col]];
                };
                if (io.post_dtype == 1) {
                    float* p = (float*)io.post + (size_t)f * N;
                    for (int col = 0; col < C; col++)
                        if (ALL_ACTIVE || active) p[col * Z + tid] = at(col);
                } else {
                    double* p = (double*)io.post + (size_t)f * N;
                    for (int col = 0; col < C; col++)
                        if (ALL_ACTIVE || active) p[col * Z + tid] = (double)at(col);
                }
            }
            // hb holds the packed decisions of the final posteriors (:5421).  Error counts are popcounts -- the
            // rotation of a column does not matter for them; bits >= R are information bits (bp_simulation.cpp:738),
            // i.e. the block columns >= B
            {
                const int lane = tid & 31;
                int nerr = 0, nerr_info = 0;
                for (int t = tid; t < ((C * HW + 31) & ~31); t += ZP) {
                    const unsigned w = t < C * HW ? hbf[t] : 0u;
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
                    if constexpr (Z % 32 == 0) {
                        // output word g = bits k0..k0+31 of one column = positions (k0 + RI) mod Z ...: a 32-bit window of hb
                        for (int g = tid; g < NWORDS; g += ZP) {
                            const int col = g / HW, k0 = 32 * (g - col * HW);
                            int start = k0 + K::rt_ri()[col];
                            if (start >= Z) start -= Z;
                            const unsigned* hc = hbf + col * HW;
                            const int i0 = start >> 5, i1 = i0 + 1 < HW ? i0 + 1 : 0;
                            out[g] = __funnelshift_r(hc[i0], hc[i1], start & 31);
                        }
                    } else {
                        constexpr int NROUND = (N + 31) & ~31;
                        for (int i = tid; i < NROUND; i += ZP) {
                            int bit = 0;
                            if (i < N) { const int col = i / Z, k = i - col * Z; bit = soft2[col * CS + k + K::rt_ri()[col]] < 0.0f; }
                            const unsigned w = __ballot_sync(0xffffffffu, bit);
                            if (lane == 0) out[i >> 5] = w;
                        }
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
        }

        // every thread's TMEM traffic is complete (wait::ld / wait::st above); hand the columns back
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid < 32) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase), "r"((unsigned)TCOLS) : "memory");
        }
    }
};

} // namespace ldpcb200
