// Shared-memory throughput kernel for LMS_DEC -- layered offset min-sum, lmin_sum_decod_qc_lm,
// decoders.cpp:5064-5425 -- in fp32.  Results are bit-identical to the fp32 restatement of the
// reference algorithm (oracle orc_lms_f32), which gives the reference's (double) hard decisions and
// iteration counts on >= 99.99 % of frames (SURVEY.md §8a'); the bit-exact double path is the
// table-driven kernel in dec_minsum.cu.
//
// Mapping.  One frame per CTA at a time (persistent grid, atomic work counter, so a CTA whose frame
// exits early immediately pulls the next one); thread n of the CTA is lane n of every block row, i.e.
// check row (j, n) for j = 0..b-1.  Layers are sequential (__syncthreads between them), lanes of one
// layer touch disjoint bits and run in parallel -- the reference's own loop order over n is immaterial.
//
// Shared memory per frame:
//   soft2 : c block columns x 2Z floats.  Each column is stored TWICE back to back, so that lane n
//           reads bit (n + shift) mod Z at offset n + shift without the wrap: a read costs no address
//           arithmetic beyond one add of a per-edge constant.  Writes go to both copies.
//   min1, min2 : R floats, the two smallest |v2c| - beta of each check row (MS_DEC_STATE, decoders.h:115-121)
//   ps   : R words: the c2v sign bit of each edge of the row, left aligned (bit 31 - q for edge q),
//          and the local index of the minimum's edge in the low 8 bits
// That is 8N + 12R bytes (112 KB for the 16 x 32, Z = 256 code: two frames per SM).
//
// Arithmetic identities used (all exact, they commute with rounding because x -> x - beta and
// x -> max(x, 0) are monotone): the two minima are tracked on the raw |v2c| and the offset, the clamp
// at 0 and the 32767 ceiling (decoders.cpp:5131-5168, 4301) are applied once per row; the position of
// the minimum is only ever used to choose between min1 and min2, which are equal whenever the
// reference's first-minimum tie rule (:5012-5027) could pick a different edge.
#include <cstdlib>
#include <algorithm>
#include <string>
#include "kernels.h"
#include "channel.cuh"

namespace ldpcb200 {

namespace {

constexpr int MAXE = 704;
constexpr int MAXB = 96;
constexpr int MAXDEG_FAST = 24;

struct LmsTab {                     // lives in the kernel-parameter constant bank: uniform loads
    int b, c, Z, N, R, E, nwords;
    unsigned short rp[MAXB + 1];
    unsigned short thr4[MAXE];      // 4 * (Z - shift): lanes with 4n >= thr4 wrap
    unsigned short sh[MAXE];        // shift
    unsigned char col[MAXE];        // block column
    unsigned int rd[MAXE];          // byte offset of bit (col, shift) in soft2: 4 * (col * 2Z + shift)
};

enum { CNT_FRAMES = 0, CNT_FRAME_ERRORS, CNT_INFO_BIT_ERRORS, CNT_UNDETECTED, CNT_ITER_SUM, CNT_BIT_ERRORS };

__device__ __forceinline__ float lds_f(const unsigned char* base, unsigned off) { return *(const float*)(base + off); }
__device__ __forceinline__ void sts_f(unsigned char* base, unsigned off, float v) { *(float*)(base + off) = v; }

// One layer for one lane: DEG is the row weight (compile time, so that the v2c values stay in
// registers and every per-edge constant index is an immediate).
template <int DEG>
__device__ __forceinline__ void lms_layer(const LmsTab& T, unsigned char* sm, int e0, unsigned n4, unsigned Z4,
                                          float* min1, float* min2, unsigned* ps, int r)
{
    const float pm1 = min1[r], pm2 = min2[r];
    const unsigned pps = ps[r];
    const unsigned ppos = pps & 0xffu;
    float v[DEG];
    float c1 = __int_as_float(0x7f800000), c2 = c1;         // +inf; the 32767 ceiling is applied below
    unsigned sacc = 0;
#pragma unroll
    for (int q = 0; q < DEG; q++) {
        const float sv = lds_f(sm, T.rd[e0 + q] + n4);
        const float pabs = ppos == (unsigned)q ? pm2 : pm1;                              // :5152
        const float pval = __uint_as_float(__float_as_uint(pabs) ^ ((pps << q) & 0x80000000u));   // :5156
        const float vv = sv - pval;                                                      // :5158
        v[q] = vv;
        sacc ^= __float_as_uint(vv);
        const float a = fabsf(vv);
        c2 = fminf(c2, fmaxf(c1, a));                                                    // process_check_node :5012-5027
        c1 = fminf(c1, a);
    }
    const float m1 = fminf(fmaxf(c1 - 0.4f, 0.0f), 32767.0f);                            // :5166-5168, :5131-5137
    const float m2 = fminf(fmaxf(c2 - 0.4f, 0.0f), 32767.0f);
    const unsigned rs = sacc & 0x80000000u;                                              // sign of the row
    const unsigned m1x = __float_as_uint(m1) ^ rs, m2x = __float_as_uint(m2) ^ rs;
    unsigned S = 0, pos = 0;
#pragma unroll
    for (int q = DEG - 1; q >= 0; q--) {
        const bool ismin = fabsf(v[q]) == c1;
        pos = ismin ? (unsigned)q : pos;                                                 // first minimum wins (reverse scan)
        const unsigned cv = (ismin ? m2x : m1x) ^ (__float_as_uint(v[q]) & 0x80000000u); // :5193-5198
        S = (S >> 1) | (cv & 0x80000000u);
        const float nv = v[q] + __uint_as_float(cv);                                     // :5199-5204
        const unsigned a0 = T.rd[e0 + q] + n4;
        sts_f(sm, a0, nv);
        if (n4 >= T.thr4[e0 + q]) sts_f(sm, a0 - Z4, nv); else sts_f(sm, a0 + Z4, nv);
    }
    min1[r] = m1; min2[r] = m2; ps[r] = S | pos;                                          // :5179
}

__device__ __forceinline__ void lms_layer_dispatch(const LmsTab& T, unsigned char* sm, int e0, int deg, unsigned n4,
                                                   unsigned Z4, float* min1, float* min2, unsigned* ps, int r)
{
    switch (deg) {
#define L(D) case D: lms_layer<D>(T, sm, e0, n4, Z4, min1, min2, ps, r); break;
    L(1) L(2) L(3) L(4) L(5) L(6) L(7) L(8) L(9) L(10) L(11) L(12) L(13) L(14) L(15) L(16)
    L(17) L(18) L(19) L(20) L(21) L(22) L(23) L(24)
#undef L
    }
}

// Syndrome of the hard decisions (check_syndrome, decoders.cpp:793-814) on packed bits.
// Step 1: every warp packs the sign bits of 32 consecutive bits of each block column with one ballot
// (hb[col * HW + w], HW = Zp / 32 words; bits >= Z of the last word are 0).  Step 2: one thread per (block
// row, 32-lane block) XORs, edge by edge, the 32-bit window of the column's bit string that starts at
// (32 w + shift) mod Z -- a funnel shift of two words, plus the wrap when the window crosses bit Z.
// That is ~4 instructions per edge and 32 check rows instead of ~4 per edge and check row.
__device__ __forceinline__ int lms_syndrome(const LmsTab& T, const float* soft2, unsigned* hb, int tid, int nt)
{
    const int Z = T.Z, HW = nt >> 5, lane = tid & 31, warp = tid >> 5;
    for (int col = 0; col < T.c; col++) {
        const int bit = tid < Z ? soft2[col * 2 * Z + tid] < 0.0f : 0;
        const unsigned w = __ballot_sync(0xffffffffu, bit);
        if (lane == 0) hb[col * HW + warp] = w;
    }
    __syncthreads();
    const int NB = (Z + 31) >> 5;
    unsigned bad = 0;
    for (int t = tid; t < T.b * NB; t += nt) {
        const int j = t / NB, w = t - j * NB;
        unsigned acc = 0;
        for (int e = T.rp[j]; e < T.rp[j + 1]; e++) {
            const unsigned* hc = hb + T.col[e] * HW;
            int start = 32 * w + T.sh[e];
            if (start >= Z) start -= Z;
            unsigned win = __funnelshift_r(hc[start >> 5], hc[min((start >> 5) + 1, HW - 1)], start & 31);
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

__global__ void __launch_bounds__(512) lms_fast_kernel(const __grid_constant__ LmsTab T, const __grid_constant__ FrameIO io)
{
    extern __shared__ __align__(16) unsigned char sm[];
    const int Z = T.Z, N = T.N, R = T.R, nt = blockDim.x, tid = threadIdx.x;
    float* soft2 = (float*)sm;
    float* min1 = soft2 + 2 * N;
    float* min2 = min1 + R;
    unsigned* ps = (unsigned*)(min2 + R);
    int* s_misc = (int*)(ps + R);                   // [0] next frame, [1] bit errors, [2] info-bit errors
    unsigned* hb = (unsigned*)s_misc;               // packed hard decisions, c * nt / 32 words; aliases s_misc, which is
                                                    // only live between frames (keeps two frames per SM at Z = 256)
    const bool active = tid < Z;
    const unsigned n4 = 4u * tid, Z4 = 4u * Z;
    const bool noexit = io.flags & LDPCB200_NO_EARLY_EXIT;

    for (;;) {
        __syncthreads();
        if (tid == 0) { s_misc[0] = (int)atomicAdd(io.next_frame, 1u); s_misc[1] = 0; s_misc[2] = 0; }
        __syncthreads();
        const int f = s_misc[0];
        if (f >= io.nf) break;

        // ---- first load: channel LLRs -> both copies of every block column; check state := 0 (:5088-5108)
        if (io.ch.enabled) {
            const unsigned long long frame = io.ch.first_frame + (unsigned long long)f;
            if (io.ch.m > 2) {
                // QAM-16/64/256: one thread per PAM component, m/2 LLRs from one demodulation
                const int half = io.ch.m >> 1, ncomp = 2 * (N / io.ch.m);
                for (int c = tid; c < ncomp; c += nt) {
                    float o[4];
                    channel_llr_qam_component(io.ch, frame, c, o);
                    const int i0 = (c >> 1) * io.ch.m + (c & 1) * half;
                    for (int b = 0; b < half; b++) {
                        const int i = channel_dest(io.ch, i0 + b), col = i / Z, k = i - col * Z;
                        const float x = i >= io.ch.punct_start ? io.ch.punct_value : o[b];
                        soft2[col * 2 * Z + k] = x; soft2[col * 2 * Z + Z + k] = x;
                    }
                }
            } else {
                for (int i4 = tid; i4 < N / 4; i4 += nt) {                            // one Philox block -> four LLRs
                    float o[4];
                    int d[4];
                        channel_llr4_bpsk(io.ch, frame, i4, o, d);
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        const int i = d[b], col = i / Z, k = i - col * Z;
                        soft2[col * 2 * Z + k] = o[b]; soft2[col * 2 * Z + Z + k] = o[b];
                    }
                }
                for (int j = (N & ~3) + tid; j < N; j += nt) {
                    const int i = channel_dest(io.ch, j), col = i / Z, k = i - col * Z;
                    const float x = channel_llr(io.ch, frame, i);
                    soft2[col * 2 * Z + k] = x; soft2[col * 2 * Z + Z + k] = x;
                }
            }
        } else if (io.llr_dtype == LDPCB200_F32) {
            const float* y = (const float*)io.llr + (size_t)f * N;
            for (int col = 0; col < T.c; col++)
                for (int k = tid; k < Z; k += nt) {
                    const float x = __ldcs(y + col * Z + k);
                    soft2[col * 2 * Z + k] = x; soft2[col * 2 * Z + Z + k] = x;
                }
        } else {
            const double* y = (const double*)io.llr + (size_t)f * N;
            for (int col = 0; col < T.c; col++)
                for (int k = tid; k < Z; k += nt) {
                    const float x = (float)__ldcs(y + col * Z + k);
                    soft2[col * 2 * Z + k] = x; soft2[col * 2 * Z + Z + k] = x;
                }
        }
        for (int i = tid; i < R; i += nt) { min1[i] = 0.0f; min2[i] = 0.0f; ps[i] = 0u; }
        __syncthreads();

        int parity = lms_syndrome(T, soft2, hb, tid, nt);                               // :5111-5115
        int ret = 0, locked = 0, iter;
        if (!parity) { ret = 1; locked = 1; }                                           // already a codeword: 0 + 1
        for (iter = 0; iter < io.maxiter; iter++) {
            if (!parity && !noexit) break;                                              // :5119
            for (int j = 0; j < T.b; j++) {
                const int e0 = T.rp[j], deg = T.rp[j + 1] - e0;
                if (active) lms_layer_dispatch(T, sm, e0, deg, n4, Z4, min1, min2, ps, j * Z + tid);
                __syncthreads();
            }
            parity = lms_syndrome(T, soft2, hb, tid, nt);                               // :5281-5284
            if (!parity && !locked) { ret = iter + 1; locked = 1; }
            if (!parity && !noexit) break;
        }
        if (!locked) ret = parity ? -iter : iter + 1;                                   // :5424

        // ---- outputs: posterior, packed decisions (soft < 0, :5421), error counts vs the all-zero codeword
        if (io.post) {
            if (io.post_dtype == LDPCB200_F32) {
                float* p = (float*)io.post + (size_t)f * N;
                for (int col = 0; col < T.c; col++)
                    for (int k = tid; k < Z; k += nt) p[col * Z + k] = soft2[col * 2 * Z + k];
            } else {
                double* p = (double*)io.post + (size_t)f * N;
                for (int col = 0; col < T.c; col++)
                    for (int k = tid; k < Z; k += nt) p[col * Z + k] = (double)soft2[col * 2 * Z + k];
            }
        }
        __syncthreads();
        if (tid == 0) { s_misc[1] = 0; s_misc[2] = 0; }      // hb (aliased) is dead from here on
        __syncthreads();
        {
            const int lane = tid & 31;
            int nerr = 0, nerr_info = 0;
            const int nround = (N + 31) & ~31;
            for (int i = tid; i < nround; i += nt) {
                int bit = 0;
                if (i < N) { const int col = i / Z, k = i - col * Z; bit = soft2[col * 2 * Z + k] < 0.0f; }
                const unsigned w = __ballot_sync(0xffffffffu, bit);
                if (lane == 0) {
                    if (io.hard_words) io.hard_words[(size_t)f * T.nwords + (i >> 5)] = w;
                    nerr += __popc(w);
                    const int lo = R - i;                   // bits >= R are information bits (bp_simulation.cpp:738)
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
                    io.per_frame[f] = (e ? 0x80000000u : 0u) | (ret >= 0 ? 0x40000000u : 0u) | (uint32_t)min(ei, 0xFFFFFF);
                if (io.counters) {
                    atomicAdd(&io.counters[CNT_FRAMES], 1ull);
                    atomicAdd(&io.counters[CNT_ITER_SUM], (unsigned long long)(ret < 0 ? -ret : ret));
                    if (e) {
                        atomicAdd(&io.counters[CNT_FRAME_ERRORS], 1ull);
                        atomicAdd(&io.counters[CNT_INFO_BIT_ERRORS], (unsigned long long)ei);
                        atomicAdd(&io.counters[CNT_BIT_ERRORS], (unsigned long long)e);
                        if (ret >= 0) atomicAdd(&io.counters[CNT_UNDETECTED], 1ull);
                    }
                }
            }
        }
    }
}

size_t lms_fast_smem(const QcHost& g)
{
    const int zp = (g.Z + 31) & ~31;
    const size_t hb = (size_t)4 * g.c * (zp / 32);
    return (size_t)8 * g.N + (size_t)12 * g.R + (hb > 16 ? hb : 16);
}

} // namespace

int find_lms_spec_aot(const QcHost& g, int kind);
void lms_spec_aot_info(int idx, const char** name, int* threads, int* minb, size_t* smem);
cudaError_t launch_lms_spec_aot(int idx, const FrameIO& io, int grid, cudaStream_t s);

const void* lms_spec_jit(const QcHost& g, int zp, int minb, int variant, int kind, std::string& why);
cudaError_t launch_lms_spec_jit(const void* kernel, int zp, size_t smem, const FrameIO& io, int grid, cudaStream_t s);

// launch geometry of the code-specialised kernel for g; false if the code does not suit it.
// *variant = 0: every block column doubled in shared memory, all check state in registers (small codes);
//            1: single copy + sign/position words in shared memory, min1/min2 in registers (large codes)
//            2: doubled columns in the rotation of their last writer, c2v messages in tensor memory (lms_tmem.cuh);
//               chosen when a frame's E*Z messages fit the 512 TMEM columns often enough to keep >= 12 warps on an SM
size_t lms_tmem_smem_bytes(int b, int c, int Z, int maxdeg);
size_t lms_tmem_pad_smem(size_t smem, int minb);
size_t lms_tmem2_smem_bytes(int b, int c, int Z, int maxdeg);

bool lms_spec_geometry(const QcHost& g, int smem_per_sm, int smem_per_block, int* zp, int* minb, size_t* smem, int* variant, bool allow_tmem)
{
    if (g.E > 512 || g.Z > 1024 || g.maxdeg > MAXDEG_FAST) return false;
    *zp = (g.Z + 31) & ~31;
    const int hw = *zp / 32;
    const char* t2 = getenv("LDPCB200_TMEM2");
    if (allow_tmem && t2 && *t2 == '1' && g.Z % 32 == 0 && g.maxdeg <= 16 && g.c * hw < 1023 && 2 * g.E * ((hw + 3) / 4) <= 512 && hw >= 4) {
        // two frames per CTA (lms_tmem2.cuh): one CTA per SM owns all the tensor memory it needs
        const size_t need = lms_tmem2_smem_bytes(g.b, g.c, g.Z, g.maxdeg);
        if (need <= (size_t)smem_per_block) {
            *minb = 1;
            *smem = need;
            *variant = 3;
            return true;
        }
    }
    if (allow_tmem && g.maxdeg <= 32 && g.c * hw < 1023) {
        int tcols = 32;
        while (tcols < g.E * ((hw + 3) / 4)) tcols *= 2;
        const size_t need = lms_tmem_smem_bytes(g.b, g.c, g.Z, g.maxdeg);
        if (tcols <= 512 && need <= (size_t)smem_per_block) {
            int m = 512 / tcols;
            m = std::min(m, (int)((size_t)smem_per_sm / (need + 1024)));
            m = std::min(m, 2048 / *zp);
            m = std::min(m, 65536 / (*zp * 64));                    // at least 64 registers per thread
            if (m >= 1 && m * hw >= 12) {
                *minb = m;
                *smem = std::min(lms_tmem_pad_smem(need, m), (size_t)smem_per_block);
                *variant = 2;
                return true;
            }
        }
    }
    const size_t hb = (size_t)(g.c * hw > 4 ? g.c * hw : 4);
    for (int v = 0; v < 2; v++) {
        const size_t words = v == 0 ? 2 * (size_t)g.N + hb : (size_t)g.N + (size_t)g.R + hb;
        const int regs = (v == 0 ? 3 : 2) * g.b + 48;              // check state of every block row + working set
        if (v == 0 && (g.b > 32 || g.maxdeg > 16)) continue;
        if (sizeof(float) * words > (size_t)smem_per_block) continue;
        int m = (int)((size_t)smem_per_sm / (sizeof(float) * words + 1024));
        m = std::min(m, 2048 / *zp);
        m = std::min(m, 65536 / (*zp * regs));
        if (m < 1) continue;
        *smem = sizeof(float) * words;
        *minb = std::min(m, 16);
        *variant = v;
        return true;
    }
    return false;
}

FastPlan plan_lms_fast(const QcHost& g, int precision, int smem_per_sm, int smem_per_block, int allow_jit)
{
    FastPlan p;
    if (precision != 32) return p;                          // the double path stays on the bit-exact table-driven kernel
    const char* no_spec = getenv("LDPCB200_NO_SPEC");
    const char* no_tmem = getenv("LDPCB200_NO_TMEM");       // 1: keep the c2v messages register-compressed (lms_spec.cuh)
    const bool tmem = !(no_tmem && *no_tmem == '1');
    const char* no_aot = getenv("LDPCB200_NO_AOT");         // 1 (development): compile at run time even when an ahead-of-time instance exists
    int aot = -1;
    bool two = false;
    if (!(no_spec && *no_spec == '1') && !(no_aot && *no_aot == '1' && allow_jit)) {
        const char* t2 = getenv("LDPCB200_TMEM2");          // 1: two frames per CTA (lms_tmem2.cuh; measured slower) instead of one (lms_tmem.cuh)
        bool want_two = tmem && t2 && *t2 == '1';
        if (want_two) aot = find_lms_spec_aot(g, 6);
        if (aot >= 0) { p.tmem = 2; two = true; }
        else if (want_two && allow_jit) {}                  // no ahead-of-time two-frame instance: compile one below
        else if (tmem) {
            aot = find_lms_spec_aot(g, 3);                  // messages in tensor memory (lms_tmem.cuh)
            if (aot >= 0) p.tmem = 1;
        }
        if (aot < 0 && !(want_two && allow_jit)) aot = find_lms_spec_aot(g, 0);
    }
    if (aot >= 0) {                                          // a code-specialised instance exists for this matrix
        int minb = 1;
        lms_spec_aot_info(aot, nullptr, &p.threads, &minb, &p.smem_bytes);
        if (p.smem_bytes <= (size_t)smem_per_block) {
            p.ok = 1; p.variant = 1; p.frames_per_cta = two ? 2 : 1; p.ctas_per_sm = minb; p.spec_index = aot;
            return p;
        }
        p.tmem = 0;
    }
    if (allow_jit && !(no_spec && *no_spec == '1')) {       // compile one for this matrix
        int zp, minb, variant;
        size_t smem;
        if (lms_spec_geometry(g, smem_per_sm, smem_per_block, &zp, &minb, &smem, &variant, tmem)) {
            std::string why;
            const void* k = lms_spec_jit(g, zp, minb, variant, 0, why);
            if (k) {
                p.ok = 1; p.variant = 2; p.frames_per_cta = variant == 3 ? 2 : 1; p.ctas_per_sm = minb; p.threads = zp; p.smem_bytes = smem;
                p.jit_kernel = k;
                p.tmem = variant == 3 ? 2 : variant == 2;
                return p;
            }
            p.note = why;
        } else
            p.note = "code does not suit the code-specialised kernel (too many edges / block rows for registers and shared memory)";
    }
    if (g.E > MAXE || g.b > MAXB || g.maxdeg > MAXDEG_FAST || g.Z > 512) return p;
    const size_t smem = lms_fast_smem(g);
    if (smem > (size_t)smem_per_block) return p;
    p.ok = 1;
    p.variant = 0;
    p.threads = (g.Z + 31) & ~31;
    p.frames_per_cta = 1;
    p.smem_bytes = smem;
    int by_smem = (int)((size_t)smem_per_sm / (smem + 1024));
    int by_threads = 2048 / p.threads;
    p.ctas_per_sm = by_smem < by_threads ? by_smem : by_threads;
    if (p.ctas_per_sm < 1) p.ctas_per_sm = 1;
    if (p.ctas_per_sm > 16) p.ctas_per_sm = 16;
    p.tab.assign(sizeof(LmsTab), 0);
    LmsTab& T = *reinterpret_cast<LmsTab*>(p.tab.data());
    T.b = g.b; T.c = g.c; T.Z = g.Z; T.N = g.N; T.R = g.R; T.E = g.E; T.nwords = (g.N + 31) / 32;
    for (int j = 0; j <= g.b; j++) T.rp[j] = (unsigned short)g.rp[j];
    for (int e = 0; e < g.E; e++) {
        T.thr4[e] = (unsigned short)(4 * (g.Z - g.sh[e]));
        T.rd[e] = 4u * (unsigned)(g.col[e] * 2 * g.Z + g.sh[e]);
        T.sh[e] = (unsigned short)g.sh[e];
        T.col[e] = (unsigned char)g.col[e];
    }
    return p;
}

cudaError_t launch_lms_fast(const FastPlan& p, const FrameIO& io, int grid, cudaStream_t s)
{
    if (p.variant == 1) return launch_lms_spec_aot(p.spec_index, io, grid, s);
    if (p.variant == 2) return launch_lms_spec_jit(p.jit_kernel, p.threads, p.smem_bytes, io, grid, s);
    const LmsTab& T = *reinterpret_cast<const LmsTab*>(p.tab.data());
    cudaError_t e = cudaFuncSetAttribute(lms_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes);
    if (e != cudaSuccess) return e;
    lms_fast_kernel<<<grid, p.threads, p.smem_bytes, s>>>(T, io);
    return cudaGetLastError();
}

} // namespace ldpcb200
