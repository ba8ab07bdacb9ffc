// Ahead-of-time instances of the code-specialised LMS_DEC kernel (lms_spec.cuh) for the frozen benchmark
// matrices, generated at build time by tools/gen_lms_spec.py into build/lms_spec_aot_gen.h.  A handle whose
// base matrix and lifting size match one of them runs it; every other code goes to the run-time compiled
// instance (spec_jit.cpp) or, failing that, to the table-driven kernel of lms_fast.cu.
#include <cstring>
#include <vector>

#include "kernels.h"
#include "channel.cuh"
#include "lms_spec.cuh"
#include "lms_tmem.cuh"
#include "lms_tmem2.cuh"
#include "ms_spec.cuh"
#include "ms_tmem.cuh"
#include "ims_h2.cuh"

namespace ldpcb200 {

size_t ms_spec_smem_bytes(int c, int Z);

// shared memory of LmsTmem<K> (lms_tmem.cuh, SMEM_WORDS): posteriors | packed decisions | syndrome plan | mbarrier | misc
size_t lms_tmem_smem_bytes(int b, int c, int Z, int maxdeg)
{
    const int zp = (Z + 31) / 32 * 32, hw = zp / 32, nb = (Z + 31) / 32, nwarps = zp / 32;
    // Z a multiple of 32: two buffers of padded single-copy columns (LmsTmem::PP), else one buffer of doubled columns
    const bool pp = Z % 32 == 0;
    const size_t soft = pp ? 2 * (size_t)c * (Z + 32) : 2 * (size_t)c * Z, hb = (size_t)(c * hw > 3 ? c * hw : 3) + 1;
    const size_t plan = (size_t)((b * nb + 8 * nwarps - 1) / (8 * nwarps)) * ((maxdeg + 3) / 4) * zp;
    const size_t mbar = (soft + (pp ? 2 : 1) * hb + plan + zp + 1) & ~(size_t)1;     // + the quick-look word per thread
    return sizeof(float) * (mbar + 2 + 4 + (pp ? c : 0));
}

// shared memory of LmsTmem2<K> (lms_tmem2.cuh, SMEM_WORDS): two buffers of float2 posteriors | packed decisions x 2 | syndrome plan | misc | column weights
size_t lms_tmem2_smem_bytes(int b, int c, int Z, int maxdeg)
{
    const int zp = (Z + 31) / 32 * 32, hw = zp / 32, nb = (Z + 31) / 32, nwarps = zp / 32;
    const size_t soft = 4 * (size_t)c * (Z + 32), hb = (size_t)(c * hw > 3 ? c * hw : 3) + 1;
    const size_t plan = (size_t)((b * nb + 8 * nwarps - 1) / (8 * nwarps)) * ((maxdeg + 3) / 4) * zp;
    const size_t misc = (soft + 2 * hb + plan + zp + 1) & ~(size_t)1;
    return sizeof(float) * (misc + 8 + c);
}

// shared memory of MsTmem<K, IS_INT> (ms_tmem.cuh, SMEM_WORDS)
size_t ms_tmem_smem_bytes(int c, int Z, bool is_int)
{
    const size_t y_off = 2 * (size_t)c * Z, n = (size_t)c * Z;
    const size_t mbar = (y_off + n + 1) & ~(size_t)1;
    const size_t sq = (mbar + 2 + 4 + 1) & ~(size_t)1;
    return sizeof(float) * (sq + (is_int ? 1024 : 0));
}

// shared memory of ImsH2<K> (ims_h2.cuh, SMEM_WORDS)
size_t ims_h2_smem_bytes(int c, int Z, int groups)
{
    const size_t y_off = 2 * (size_t)c * Z, n = (size_t)c * Z;
    const size_t group = (y_off + n + 1) & ~(size_t)1;
    const int zp = (Z + 31) / 32 * 32;
    return sizeof(float) * (groups * group + 2 + 16 + 2 * (groups * zp / 32 + 2));
}

size_t lms_tmem_pad_smem(size_t smem, int minb)
{
    const size_t per_sm = 233472, reserve = 1024;               // sm_100: 228 KB per SM, 1 KB reserved per CTA
    const size_t floor_ = per_sm / (size_t)(minb + 1) - reserve + 16;
    return smem > floor_ ? smem : (floor_ + 15) / 16 * 16;
}

struct SpecEntry {
    const char* name;
    const void* kernel;
    int kind;                       // 0 LMS_DEC, 1 MS_DEC fp32, 2 IMS_DEC; 3 / 4 / 5 = LMS_DEC / MS_DEC / IMS_DEC with the messages in tensor memory;
                                    // 6 = LMS_DEC, tensor memory, two frames per CTA (lms_tmem2.cuh); 7 / 8 = IMS_DEC, tensor memory, frames as fp16 pairs, one / two groups per CTA (ims_h2.cuh; zp = threads per CTA)
    int b, c, Z, E, zp, minb, maxdeg;
    const int *rp, *col, *sh;       // host copies for matching
};

static std::vector<SpecEntry>& registry()
{
    static std::vector<SpecEntry> r;
    return r;
}

struct SpecRegistrar {
    SpecRegistrar(const SpecEntry& e) { registry().push_back(e); }
};

} // namespace ldpcb200

#define LDPC_SPEC_REGISTER(NAME, B_, C_, Z_, E_, ZP_, MINB_)                                                   \
    static ldpcb200::SpecRegistrar reg_##NAME(ldpcb200::SpecEntry{#NAME, (const void*)lms_spec_##NAME, 0, B_, C_, Z_, E_, ZP_, MINB_, ldpcb200::gen_##NAME::Code::MAXDEG, \
        ldpcb200::gen_##NAME::Code::RP, ldpcb200::gen_##NAME::Code::COL, ldpcb200::gen_##NAME::Code::SH});
#define LDPC_MS_SPEC_KIND_ms 1
#define LDPC_MS_SPEC_KIND_ims 2
#define LDPC_MS_SPEC_KIND_lmst 3
#define LDPC_MS_SPEC_KIND_mst 4
#define LDPC_MS_SPEC_KIND_imst 5
#define LDPC_MS_SPEC_KIND_lmst2 6
#define LDPC_MS_SPEC_KIND_imsh 7
#define LDPC_MS_SPEC_KIND_imsh2 8
#define LDPC_MS_SPEC_REGISTER(KIND, NAME, B_, C_, Z_, E_, ZP_, MINB_)                                          \
    static ldpcb200::SpecRegistrar reg_##NAME(ldpcb200::SpecEntry{#NAME, (const void*)KIND##_spec_##NAME, LDPC_MS_SPEC_KIND_##KIND, B_, C_, Z_, E_, ZP_, MINB_, ldpcb200::gen_##NAME::Code::MAXDEG, \
        ldpcb200::gen_##NAME::Code::RP, ldpcb200::gen_##NAME::Code::COL, ldpcb200::gen_##NAME::Code::SH});

#include "lms_spec_aot_gen.h"

namespace ldpcb200 {

// -> index of the matching ahead-of-time instance, or -1
int find_lms_spec_aot(const QcHost& g, int kind)
{
    const std::vector<SpecEntry>& r = registry();
    for (size_t k = 0; k < r.size(); k++) {
        const SpecEntry& e = r[k];
        if (e.kind != kind || e.b != g.b || e.c != g.c || e.Z != g.Z || e.E != g.E) continue;
        bool same = true;
        for (int j = 0; j <= g.b && same; j++) same = e.rp[j] == g.rp[j];
        for (int i = 0; i < g.E && same; i++) same = e.col[i] == g.col[i] && e.sh[i] == g.sh[i];
        if (same) return (int)k;
    }
    return -1;
}

void lms_spec_aot_info(int idx, const char** name, int* threads, int* minb, size_t* smem)
{
    const SpecEntry& e = registry()[idx];
    const int hw = e.zp / 32;
    if (name) *name = e.name;
    if (threads) *threads = e.zp;
    if (minb) *minb = e.minb;
    if (smem) {
        if (e.kind == 0) *smem = sizeof(float) * (2 * (size_t)e.c * e.Z + (e.c * hw > 4 ? e.c * hw : 4));
        else if (e.kind == 3) *smem = lms_tmem_smem_bytes(e.b, e.c, e.Z, e.maxdeg);
        else if (e.kind == 6) *smem = lms_tmem2_smem_bytes(e.b, e.c, e.Z, e.maxdeg);
        else if (e.kind == 4 || e.kind == 5) *smem = ms_tmem_smem_bytes(e.c, e.Z, e.kind == 5);
        else if (e.kind == 7 || e.kind == 8) *smem = ims_h2_smem_bytes(e.c, e.Z, e.kind - 6);
        else *smem = ms_spec_smem_bytes(e.c, e.Z);
        // tensor-memory variant: exactly `minb` CTAs may share an SM (their TMEM columns add up to 512; one more
        // resident CTA would sit in tcgen05.alloc until another exits), so the request is padded until minb + 1
        // no longer fit into the 228 KB of an sm_100 SM
        if ((e.kind >= 3 && e.kind <= 5) || e.kind == 7 || e.kind == 8) *smem = lms_tmem_pad_smem(*smem, e.minb);
    }
}

const void* lms_spec_aot_kernel(int idx) { return registry()[idx].kernel; }

cudaError_t launch_lms_spec_aot(int idx, const FrameIO& io, int grid, cudaStream_t s)
{
    const SpecEntry& e = registry()[idx];
    size_t smem;
    lms_spec_aot_info(idx, nullptr, nullptr, nullptr, &smem);
    cudaError_t err = cudaFuncSetAttribute(e.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    void* args[] = { (void*)&io };
    return cudaLaunchKernel(e.kernel, dim3(grid), dim3(e.zp), args, smem, s);
}

} // namespace ldpcb200
