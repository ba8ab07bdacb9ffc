// Host side of the code-specialised flooding min-sum kernels (ms_spec.cuh): MS_DEC in fp32 and IMS_DEC.
// Picks an ahead-of-time instance when the matrix is a built-in one (lms_spec_aot.cu), else compiles one at run time
// (spec_jit.cpp) when the handle allows it; otherwise the handle stays on the table-driven kernel of dec_minsum.cu.
#include <algorithm>
#include <cstdlib>
#include <string>

#include "kernels.h"

namespace ldpcb200 {

int find_lms_spec_aot(const QcHost& g, int kind);
void lms_spec_aot_info(int idx, const char** name, int* threads, int* minb, size_t* smem);
const void* lms_spec_aot_kernel(int idx);
const void* lms_spec_jit(const QcHost& g, int zp, int minb, int variant, int kind, std::string& why);
cudaError_t launch_ms_spec(const void* kernel, int zp, size_t smem, const FrameIO& io, const MsSpecParams& sp, int grid, cudaStream_t s);

// doubled posteriors + channel values + scratch + 512 doubles of staging for the IMS energy sum
size_t ms_spec_smem_bytes(int c, int Z)
{
    const size_t words = ((size_t)3 * c * Z + 4 + 1) & ~(size_t)1;
    return 4 * words + 8 * 512;
}

size_t ms_tmem_smem_bytes(int c, int Z, bool is_int);
size_t ims_h2_smem_bytes(int c, int Z, int groups);
size_t lms_tmem_pad_smem(size_t smem, int minb);

// groups of (Z rounded up to 32) threads per CTA of the fp16-pair IMS_DEC kernel: two when a group is at most four warps
// (LDPCB200_IMS_H2_GROUPS=1|2 overrides; development)
int ims_h2_groups(int Z)
{
    const int zp = (Z + 31) & ~31;
    int g = zp <= 128 ? 2 : 1;
    if (const char* e = getenv("LDPCB200_IMS_H2_GROUPS"))
        if (atoi(e) == 1 || (atoi(e) == 2 && 2 * zp <= 1024)) g = atoi(e);
    return g;
}

// *variant = 0: check state register-compressed (ms_spec.cuh); 2: messages in tensor memory (ms_tmem.cuh); 4 / 5: IMS_DEC with
// the frames as fp16 pairs, one / two groups of *zp threads per CTA, messages in tensor memory (ims_h2.cuh; the caller asks
// for it with *variant = 4).  *zp = threads per group.
bool ms_spec_geometry(const QcHost& g, int kind, int smem_per_sm, int smem_per_block, int* zp, int* minb, size_t* smem, int* variant, bool allow_tmem)
{
    const bool want_h2 = *variant == 4 && kind == 2;
    if (g.E > 512 || g.Z > 1024) return false;
    for (int i = 0; i < g.c; i++)
        if (g.cp[i + 1] == g.cp[i]) return false;          // pass A initialises a bit's accumulator through its first edge
    *zp = (g.Z + 31) & ~31;
    *variant = 0;
    if (allow_tmem && g.maxdeg <= 32) {                     // (plan_ms_fast clears allow_tmem when IMS_DEC's integers would not be exact in fp32)
        const int groups = want_h2 ? ims_h2_groups(g.Z) : 1, nt = groups * *zp;
        const int hw = nt / 32;
        int tcols = 32;
        while (tcols < g.E * ((hw + 3) / 4)) tcols *= 2;
        const size_t need = want_h2 ? ims_h2_smem_bytes(g.c, g.Z, groups) : ms_tmem_smem_bytes(g.c, g.Z, kind == 2);
        if (tcols <= 512 && need <= (size_t)smem_per_block && nt <= 1024) {
            int m = 512 / tcols;
            m = std::min(m, (int)((size_t)smem_per_sm / (need + 1024)));
            m = std::min(m, 2048 / nt);
            m = std::min(m, 65536 / (nt * (want_h2 ? 128 : 64)));
            if (m >= 1 && m * hw >= (want_h2 ? 8 : 12)) {
                *minb = m;
                *smem = std::min(lms_tmem_pad_smem(need, m), (size_t)smem_per_block);
                *variant = want_h2 ? 3 + groups : 2;
                return true;
            }
        }
    }
    if (want_h2) return false;
    if (g.b > 32 || g.maxdeg > 16) return false;
    *smem = ms_spec_smem_bytes(g.c, g.Z);
    if (*smem > (size_t)smem_per_block) return false;
    const int regs = 3 * g.b + 72;
    int m = (int)((size_t)smem_per_sm / (*smem + 1024));
    m = std::min(m, 2048 / *zp);
    m = std::min(m, 65536 / (*zp * regs));
    if (m < 1) return false;
    *minb = std::min(m, 16);
    return true;
}

// kind: 1 = MS_DEC (precision 32 only), 2 = IMS_DEC
FastPlan plan_ms_fast(const QcHost& g, int kind, int precision, int smem_per_sm, int smem_per_block, int allow_jit, const DecParams& dp)
{
    FastPlan p;
    if (kind == 1 && precision != 32) return p;             // the double MS_DEC has its own kernel (tasp_fast.cu ms64_fast_kernel)
    const char* no_spec = getenv("LDPCB200_NO_SPEC");
    if (no_spec && *no_spec == '1') return p;
    const char* no_tmem = getenv("LDPCB200_NO_TMEM");       // 1: keep the check state register-compressed (ms_spec.cuh)
    bool tmem = !(no_tmem && *no_tmem == '1');
    if (kind == 2) {
        // ms_tmem.cuh carries IMS_DEC's integers as floats: exact while max_data * ialpha < 2^24 (always, for sane alpha)
        const double ialpha = (double)(int)(dp.alpha * 16), max_data = (double)((1L << (dp.dbits - 1)) - 1);
        if (!(dp.alpha >= 0) || ialpha * max_data >= 16777216.0) tmem = false;
    }
    const char* no_aot = getenv("LDPCB200_NO_AOT");         // 1 (development): compile at run time even when an ahead-of-time instance exists
    const bool use_aot = !(no_aot && *no_aot == '1' && allow_jit);
    // IMS_DEC as fp16 pairs, two frames per CTA (ims_h2.cuh): exact while every quantity is an integer below 2048 and the
    // scaling constant of its header exists -- dbits <= 8, 0 <= ialpha <= 16; LDPCB200_IMS_H2=0 keeps one frame per CTA
    bool h2 = false;
    if (kind == 2 && tmem) {
        const int ialpha = (int)(dp.alpha * 16);
        const char* e = getenv("LDPCB200_IMS_H2");
        h2 = dp.dbits <= 8 && dp.qbits <= dp.dbits && ialpha >= 0 && ialpha <= 16 && !(e && *e == '0');
    }
    // order: ahead-of-time fp16-pair instance; run-time compiled fp16-pair instance; then the one-frame-per-CTA kernels
    // (tensor memory ahead of time, register-compressed ahead of time, run-time compiled)
    auto take_aot = [&](int idx, int tm, int fpc) {
        int minb = 1;
        lms_spec_aot_info(idx, nullptr, &p.threads, &minb, &p.smem_bytes);
        if (p.smem_bytes > (size_t)smem_per_block) return false;
        p.ok = 1; p.variant = 1; p.ctas_per_sm = minb; p.spec_index = idx; p.jit_kernel = lms_spec_aot_kernel(idx);
        p.tmem = tm; p.frames_per_cta = fpc;
        return true;
    };
    auto take_jit = [&](int want_variant, bool allow_tm) {
        int zp, minb, variant = want_variant;
        size_t smem;
        if (!ms_spec_geometry(g, kind, smem_per_sm, smem_per_block, &zp, &minb, &smem, &variant, allow_tm)) {
            p.note = "code does not suit the code-specialised kernel";
            return false;
        }
        if (want_variant == 4 && variant < 4) return false;
        std::string why;
        const void* k = lms_spec_jit(g, zp, minb, variant, kind, why);
        if (!k) { p.note = why; return false; }
        p.ok = 1; p.variant = 2; p.ctas_per_sm = minb; p.threads = variant == 5 ? 2 * zp : zp; p.smem_bytes = smem; p.jit_kernel = k;
        p.tmem = variant == 2 || variant >= 4;
        p.frames_per_cta = variant >= 4 ? 2 * (variant - 3) : 1;
        return true;
    };
    int aot;
    const int groups = ims_h2_groups(g.Z);
    if (h2 && use_aot && (aot = find_lms_spec_aot(g, 6 + groups)) >= 0 && take_aot(aot, 1, 2 * groups)) return p;
    if (h2 && allow_jit && take_jit(4, true)) return p;
    if (tmem && use_aot && (aot = find_lms_spec_aot(g, kind + 3)) >= 0 && take_aot(aot, 1, 1)) return p;   // 4 / 5: messages in tensor memory (ms_tmem.cuh)
    if (use_aot && (aot = find_lms_spec_aot(g, kind)) >= 0 && take_aot(aot, 0, 1)) return p;
    if (allow_jit && take_jit(0, tmem)) return p;
    return p;
}

cudaError_t launch_ms_fast(const FastPlan& p, const DecParams& dp, const FrameIO& io, int grid, cudaStream_t s)
{
    MsSpecParams sp;
    sp.alpha = (float)dp.alpha;
    sp.ialpha = (int)(dp.alpha * (1L << 4));                    // MS_ALPHA_FPP = 4, decoders.cpp:5458
    sp.max_data = (short)((1L << (dp.dbits - 1)) - 1);          // :5445
    sp.max_quant = (short)((1L << (dp.qbits - 1)) - 1);         // :5446
    sp.thr = dp.thr;
    return launch_ms_spec(p.jit_kernel, p.threads, p.smem_bytes, io, sp, grid, s);
}

} // namespace ldpcb200
