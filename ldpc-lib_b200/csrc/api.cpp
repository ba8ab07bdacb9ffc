// C ABI of the engine (include/ldpcb200.h): handle management, staging, the host<->device pipeline and
// dispatch to the CUDA kernels.  No decoding arithmetic lives here and nothing here falls back to a
// CPU path: without a usable CUDA device every compute entry point fails with LDPCB200_ENODEV.
#include <cmath>
#include <cstdlib>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>

#include "kernels.h"

namespace ldpcb200 {
size_t minsum_workspace_bytes(int decoder_id, int precision, const QcHost& g, int nt);
cudaError_t launch_minsum_generic(int decoder_id, int precision, const QcDev& g, const DecParams& dp,
                                  const FrameIO& io, char* ws, size_t ws_stride, size_t smem_ws, int grid, int nt, cudaStream_t s);
size_t sumprod_workspace_bytes(int decoder_id, const QcHost& g, int nt);
cudaError_t launch_sumprod_generic(int decoder_id, const QcDev& g, const DecParams& dp, const FrameIO& io,
                                   char* ws, size_t ws_stride, size_t smem_ws, int grid, int nt, cudaStream_t s);
}

namespace ldpcb200 {
bool lms_spec_geometry(const QcHost& g, int smem_per_sm, int smem_per_block, int* zp, int* minb, size_t* smem, int* variant, bool allow_tmem);
std::string lms_spec_generate(const QcHost& g, int zp, int minb, int variant, int kind);
bool lms_spec_compile(const std::string& gen, int major, int minor, std::vector<char>& cubin, std::string& why);
}

using namespace ldpcb200;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(e_ == cudaErrorMemoryAllocation ? LDPCB200_ENOMEM : LDPCB200_ECUDA,        \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

bool is_minsum(int id) { return id == LDPCB200_LMS_DEC || id == LDPCB200_MS_DEC || id == LDPCB200_IMS_DEC; }
bool valid_decoder(int id)
{
    switch (id) {
    case LDPCB200_BP_DEC: case LDPCB200_SP_DEC: case LDPCB200_ASP_DEC: case LDPCB200_MS_DEC: case LDPCB200_IMS_DEC:
    case LDPCB200_IASP_DEC: case LDPCB200_TASP_DEC: case LDPCB200_LMS_DEC: case LDPCB200_LCHE_DEC: return true;
    }
    return false;
}

// a device buffer that only ever grows
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, n);
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct Slot {                       // one stage of the host-buffer pipeline
    DevBuf llr, words, bytes, iters, post, aux, perframe;
    cudaEvent_t ev_in = nullptr, ev_k = nullptr, ev_out = nullptr;
};

} // namespace

struct ldpcb200_handle_s {
    QcHost g;
    QcDev gd{};
    int decoder_id = 0;
    ldpcb200_params p{};
    DecParams dp{};
    int device = 0, num_sms = 0, smem_per_sm = 0, smem_per_block = 0;
    cudaStream_t stream = nullptr, s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    DevBuf tables, ws, counters, next, bpsynd, coef, perm, cw, cw_words;
    bool has_cw = false;                    // cw: the transmitted codeword, one byte per bit; cw_words: packed (ldpcb200_set_codeword)
    bool has_perm = false;                  // perm holds direct[N] | inverse[N] (ldpcb200_set_interleaver)
    size_t ws_stride = 0, smem_ws = 0;      // smem_ws != 0: the table-driven kernel keeps its state in shared memory
    int grid = 0, nt = 0;
    Slot slot[2];
    FastPlan fast{};
    float last_ms = 0;
    int last_launches = 0;
};

namespace {

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int upload_tables(ldpcb200_handle_s* h)
{
    const QcHost& g = h->g;
    std::vector<int> blob;
    auto put = [&](const std::vector<int>& v) { size_t o = blob.size(); blob.insert(blob.end(), v.begin(), v.end()); while (blob.size() % 4) blob.push_back(0); return o; };
    size_t o_rp = put(g.rp), o_col = put(g.col), o_sh = put(g.sh), o_row = put(g.row), o_cp = put(g.cp), o_ce = put(g.cedge);
    std::vector<int> pk(g.pk.begin(), g.pk.end());
    size_t o_pk = put(pk);
    CU(h->tables.reserve(blob.size() * sizeof(int)));
    CU(cudaMemcpy(h->tables.p, blob.data(), blob.size() * sizeof(int), cudaMemcpyHostToDevice));
    const int* base = (const int*)h->tables.p;
    QcDev& d = h->gd;
    d.b = g.b; d.c = g.c; d.Z = g.Z; d.E = g.E; d.N = g.N; d.R = g.R; d.maxdeg = g.maxdeg; d.maxcdeg = g.maxcdeg;
    d.all_cw_2 = g.all_cw_2; d.nwords = (g.N + 31) / 32;
    d.rp = base + o_rp; d.col = base + o_col; d.sh = base + o_sh; d.row = base + o_row; d.cp = base + o_cp;
    d.cedge = base + o_ce; d.pk = (const uint32_t*)(base + o_pk);
    return 0;
}

size_t dtype_size(int dt) { return dt == LDPCB200_F64 ? 8 : dt == LDPCB200_F32 ? 4 : 2; }

int default_post_dtype(int decoder_id, int precision)
{
    if (decoder_id == LDPCB200_IMS_DEC) return LDPCB200_I16;
    if (decoder_id == LDPCB200_IASP_DEC) return LDPCB200_U16;
    return precision == 32 ? LDPCB200_F32 : LDPCB200_F64;
}

// launch the decoder for io.nf frames on the handle's stream
int launch_decoder(ldpcb200_handle_s* h, FrameIO& io)
{
    CU(cudaMemsetAsync(h->next.p, 0, sizeof(unsigned int), h->stream));
    io.next_frame = (unsigned int*)h->next.p;
    io.bp_syndrome = nullptr;
    int grid = std::min(h->grid, std::max(io.nf, 1));
    if (h->decoder_id == LDPCB200_BP_DEC && (io.flags & LDPCB200_BP_CHAIN_SYNDROME)) {
        io.bp_syndrome = (uint8_t*)h->bpsynd.p;
        grid = 1;                                   // frames must follow each other, as in the reference
    }
    const bool ims_own_energy = h->decoder_id == LDPCB200_IMS_DEC && h->fast.ok && h->fast.frames_per_cta >= 2 &&
                                !(io.post && io.post_dtype != default_post_dtype(h->decoder_id, h->p.precision));   // ims_h2.cuh sums the energy itself
    if (h->decoder_id == LDPCB200_IMS_DEC && !ims_own_energy) {                // energy pre-pass: the per-frame quantiser scale
        CU(h->coef.reserve(sizeof(double) * (size_t)std::max(io.nf, 1)));
        CU(launch_ims_energy(io, h->g.N, (double*)h->coef.p, h->stream));
        io.coef = (const double*)h->coef.p;
        h->last_launches++;
    }
    bool use_fast = h->fast.ok && !(io.post && io.post_dtype != default_post_dtype(h->decoder_id, h->p.precision));
    if (use_fast && (h->decoder_id == LDPCB200_LMS_DEC || h->decoder_id == LDPCB200_MS_DEC) && h->p.precision == 64) {
        int fgrid = std::min(h->num_sms * h->fast.ctas_per_sm, io.nf);
        CU(launch_tasp_fast(h->fast, h->decoder_id, h->gd, io, std::max(fgrid, 1), h->stream, h->dp.alpha));
    } else if (use_fast && h->decoder_id == LDPCB200_LMS_DEC) {
        int fgrid = std::min(h->num_sms * h->fast.ctas_per_sm, (io.nf + h->fast.frames_per_cta - 1) / h->fast.frames_per_cta);
        if (const char* g = getenv("LDPCB200_GRID_PER_SM"))        // development: fewer resident CTAs per SM (latency experiments)
            if (atoi(g) > 0) fgrid = std::min(fgrid, h->num_sms * atoi(g));
        CU(launch_lms_fast(h->fast, io, std::max(fgrid, 1), h->stream));
    } else if (use_fast && (h->decoder_id == LDPCB200_IMS_DEC || h->decoder_id == LDPCB200_MS_DEC)) {
        int fgrid = std::min(h->num_sms * h->fast.ctas_per_sm, (io.nf + h->fast.frames_per_cta - 1) / h->fast.frames_per_cta);   // ims_h2: frames in flight per CTA
        CU(launch_ms_fast(h->fast, h->dp, io, std::max(fgrid, 1), h->stream));
    } else if (use_fast && (h->decoder_id == LDPCB200_TASP_DEC || h->decoder_id == LDPCB200_ASP_DEC || h->decoder_id == LDPCB200_LCHE_DEC || h->decoder_id == LDPCB200_IASP_DEC ||
                            ((h->decoder_id == LDPCB200_BP_DEC || h->decoder_id == LDPCB200_SP_DEC) && !io.bp_syndrome))) {
        int fgrid = std::min(h->num_sms * h->fast.ctas_per_sm, io.nf);
        if (h->fast.bpsp4) CU(launch_bpsp4(h->fast, h->decoder_id, h->gd, io, std::max(fgrid, 1), h->stream));
        else CU(launch_tasp_fast(h->fast, h->decoder_id, h->gd, io, std::max(fgrid, 1), h->stream));
    } else if (is_minsum(h->decoder_id)) {
        CU(h->ws.reserve(h->smem_ws ? 256 : h->ws_stride * h->grid));
        CU(launch_minsum_generic(h->decoder_id, h->p.precision, h->gd, h->dp, io, (char*)h->ws.p, h->ws_stride, h->smem_ws, grid, h->nt, h->stream));
    } else {
        CU(h->ws.reserve(h->smem_ws ? 256 : h->ws_stride * h->grid));
        CU(launch_sumprod_generic(h->decoder_id, h->gd, h->dp, io, (char*)h->ws.p, h->ws_stride, h->smem_ws, grid, h->nt, h->stream));
    }
    h->last_launches++;
    return 0;
}

void fill_channel(const ldpcb200_handle_s* h, const ldpcb200_sim_params* sp, ChannelParams& ch)
{
    const QcHost& g = h->g;
    memset(&ch, 0, sizeof ch);
    ch.enabled = 1;
    ch.modulation = sp->modulation;
    ch.m = sp->modulation == LDPCB200_MOD_BPSK ? 1 : 2 * sp->modulation;        // QAM4 -> 2, 16 -> 4, 64 -> 6, 256 -> 8
    double sigma = ldpcb200_sigma(g.b, g.c, sp->punctured_blocks, sp->snr_db, sp->modulation);
    ch.sigma_d = sigma;
    ch.sigma = (float)sigma;
    ch.llr_scale = (float)(2.0 / (sigma * sigma));
    ch.T = sp->qam_T > 0 ? sp->qam_T : 26.0;                                   // bp_simulation.cpp:339
    ch.punct_start = g.N - sp->punctured_blocks * g.Z;                         // bp_simulation.cpp:702-703
    // init_val = DemodOutType == 1 ? 0 : 0.5, with the out_type table of bp_simulation.cpp:451-466
    bool out1 = h->decoder_id == LDPCB200_SP_DEC || h->decoder_id == LDPCB200_ASP_DEC || h->decoder_id == LDPCB200_IASP_DEC ||
                h->decoder_id == LDPCB200_TASP_DEC || h->decoder_id == LDPCB200_LCHE_DEC;
    ch.punct_value = out1 ? 0.0f : 0.5f;
    ch.seed = sp->seed;
    ch.stream = sp->stream;
    ch.first_frame = sp->first_frame;
    if (h->has_perm) { ch.perm_dir = (const int*)h->perm.p; ch.perm_inv = (const int*)h->perm.p + g.N; }
    if (h->has_cw) ch.cw = (const unsigned char*)h->cw.p;
    if (ch.m >= 4) {                                                           // channel.cuh pam_demod_factored
        const double N0 = 2.0 * sigma * sigma, sq1 = (double)((1 << (ch.m / 2)) - 1);
        const char* ex = getenv("LDPCB200_QAM_EXACT");
        ch.qam_w = 2.0 / N0;
        ch.qam_n0inv = 1.0 / N0;
        for (int j = 0; j < 8; j++) ch.qam_c[j] = exp(-(2.0 * j + 1) * (2.0 * j + 1) / N0);
        ch.qam_fast = !(ex && *ex == '1') && sq1 * (sq1 + 6.0 * sigma) * ch.qam_w <= 600.0;
    }
}

int check_sim(const ldpcb200_handle_s* h, const ldpcb200_sim_params* sp)
{
    if (!h || !sp) return fail(LDPCB200_EINVAL, "null argument");
    if (sp->modulation < LDPCB200_MOD_BPSK || sp->modulation > LDPCB200_MOD_QAM256) return fail(LDPCB200_EINVAL, "unknown modulation %d", sp->modulation);
    if (sp->punctured_blocks < 0 || sp->punctured_blocks >= h->g.c) return fail(LDPCB200_EINVAL, "punctured_blocks out of range");
    if (sp->modulation >= LDPCB200_MOD_QAM16 && h->g.N % (2 * sp->modulation) != 0)
        return fail(LDPCB200_EINVAL, "N = %d is not a multiple of the %d bits of a symbol", h->g.N, 2 * sp->modulation);
    if (sp->max_iterations < 0) return fail(LDPCB200_EINVAL, "max_iterations < 0");
    return 0;
}

} // namespace

extern "C" {

void ldpcb200_default_params(ldpcb200_params* p)
{
    if (!p) return;
    memset(p, 0, sizeof *p);
    p->alpha = 0.8;  p->beta = 0.4;  p->thr = 1.4;  p->qbits = 6;  p->dbits = 8;    // decoders.h:43-48
    p->precision = 64;
    p->device = -1;
    p->use_fast = 1;
}

const char* ldpcb200_last_error(void) { return g_err.c_str(); }
int ldpcb200_version(void) { return LDPCB200_VERSION; }

double ldpcb200_sigma(int b, int c, int punctured_blocks, double snr_db, int modulation)
{
    double bitrate = (double)(c - b) / (c - punctured_blocks);                  // bp_simulation.cpp:444
    if (modulation == LDPCB200_MOD_BPSK) return sqrt(pow(10, -snr_db / 10) / 2 / bitrate);   // :445
    int Q = 1 << (2 * modulation);
    int halfmlog = modulation;                                                  // :403-411
    double norm_factor = 2.0 * (Q - 1.0) / 3.0;                                 // :447
    return sqrt(pow(10., -snr_db / 10.) / (2 * bitrate * halfmlog * 2) * norm_factor);       // :449
}

int ldpcb200_create(const int16_t* hd, int b, int c, int Z, int decoder_id, const ldpcb200_params* params,
                    ldpcb200_handle* out)
{
    if (!out) return fail(LDPCB200_EINVAL, "out is null");
    *out = nullptr;
    if (!valid_decoder(decoder_id)) return fail(LDPCB200_EINVAL, "unknown decoder id %d", decoder_id);
    ldpcb200_params p;
    if (params) p = *params; else ldpcb200_default_params(&p);
    if (p.precision == 0) p.precision = 64;
    if (p.precision != 64 && p.precision != 32) return fail(LDPCB200_EINVAL, "precision must be 64 or 32");
    if (p.precision == 32 && decoder_id != LDPCB200_LMS_DEC && decoder_id != LDPCB200_MS_DEC && decoder_id != LDPCB200_IMS_DEC)
        return fail(LDPCB200_EUNSUPPORTED, "precision 32 exists for LMS_DEC and MS_DEC only; the sum-product decoders run in double");
    if (decoder_id == LDPCB200_IMS_DEC && (p.qbits < 2 || p.qbits > 15 || p.dbits < 2 || p.dbits > 15 || !(p.thr > 0)))
        return fail(LDPCB200_EINVAL, "IMS_DEC needs 2 <= qbits, dbits <= 15 and thr > 0");

    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(LDPCB200_ENODEV, "no CUDA device: %s (there is no CPU fallback)", ce == cudaSuccess ? "device count is 0" : cudaGetErrorString(ce));
    }
    int dev = p.device;
    if (dev < 0) CU(cudaGetDevice(&dev));
    if (dev >= ndev) return fail(LDPCB200_EINVAL, "device %d out of range (%d devices)", dev, ndev);

    ldpcb200_handle_s* h = new ldpcb200_handle_s();
    if (!h->g.build(hd, b, c, Z)) { delete h; return fail(LDPCB200_EINVAL, "bad base matrix (need 0 < b, 0 < c <= 255, 0 < Z <= 4095, at least one circulant)"); }
    if (h->g.maxdeg > LDPCB200_MAX_ROW_WEIGHT) { int d = h->g.maxdeg; delete h; return fail(LDPCB200_EUNSUPPORTED, "row weight %d exceeds LDPCB200_MAX_ROW_WEIGHT", d); }
    bool needs_rw2 = decoder_id == LDPCB200_TASP_DEC || decoder_id == LDPCB200_ASP_DEC || decoder_id == LDPCB200_IASP_DEC;
    if (needs_rw2 && h->g.mindeg < 2) { delete h; return fail(LDPCB200_EUNSUPPORTED, "map_bin needs every row weight >= 2 (decoders.cpp:2219 reads SB[1] uninitialised)"); }
    h->decoder_id = decoder_id; h->p = p; h->device = dev;
    h->dp.alpha = p.alpha; h->dp.thr = p.thr; h->dp.qbits = p.qbits; h->dp.dbits = p.dbits;

    DeviceGuard guard(dev);
    int rc = [&]() -> int {
        cudaDeviceProp prop;
        CU(cudaGetDeviceProperties(&prop, dev));
        h->num_sms = prop.multiProcessorCount;
        h->smem_per_sm = (int)prop.sharedMemPerMultiprocessor;
        h->smem_per_block = (int)prop.sharedMemPerBlockOptin;
        CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
        CU(cudaEventCreate(&h->ev0));
        CU(cudaEventCreate(&h->ev1));
        for (Slot& s : h->slot) {
            CU(cudaEventCreateWithFlags(&s.ev_in, cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&s.ev_k, cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&s.ev_out, cudaEventDisableTiming));
        }
        int urc = upload_tables(h);
        if (urc) return urc;
        h->nt = h->g.Z <= 256 ? 256 : 512;
        h->grid = h->num_sms * (h->nt == 256 ? 2 : 1);
        size_t wsb = is_minsum(decoder_id) ? minsum_workspace_bytes(decoder_id, p.precision, h->g, h->nt)
                                           : sumprod_workspace_bytes(decoder_id, h->g, h->nt);
        h->ws_stride = (wsb + 255) & ~(size_t)255;
        // Measured on B200 (profiles/r01_all_decoders_v1 vs _v2): staging the table-driven kernels' state in shared
        // memory does not pay -- they are bound by fp64 / exp / log throughput and occupancy, and the L2-resident
        // workspace allows more CTAs per SM -- so it is opt-in (LDPCB200_WS_SMEM=1).
        const char* sm = getenv("LDPCB200_WS_SMEM");
        if (h->ws_stride + 8192 <= (size_t)h->smem_per_block && sm && *sm == '1') {   // 8 KB: the kernels' static shared variables
            h->smem_ws = h->ws_stride;
            int per_sm = (int)((size_t)h->smem_per_sm / (h->smem_ws + 8192 + 1024));
            per_sm = std::min(per_sm, h->nt == 256 ? 2 : 1);
            h->grid = h->num_sms * std::max(per_sm, 1);
        }
        // (the table-driven kernels' workspace is allocated by the first launch that needs it: launch_decoder)
        CU(h->counters.reserve(8 * sizeof(unsigned long long)));
        CU(h->next.reserve(256));
        CU(h->bpsynd.reserve((size_t)h->g.R + 16));
        CU(cudaMemset(h->bpsynd.p, 0, (size_t)h->g.R + 16));
        if (p.use_fast) {
            if (decoder_id == LDPCB200_LMS_DEC && p.precision == 64) h->fast = plan_tasp_fast(h->g, decoder_id, h->smem_per_sm, h->smem_per_block);   // double: tasp_fast.cu
            else if (decoder_id == LDPCB200_LMS_DEC) h->fast = plan_lms_fast(h->g, p.precision, h->smem_per_sm, h->smem_per_block, p.use_fast >= 2);
            else if (decoder_id == LDPCB200_IMS_DEC) h->fast = plan_ms_fast(h->g, 2, p.precision, h->smem_per_sm, h->smem_per_block, p.use_fast >= 2, h->dp);
            else if (decoder_id == LDPCB200_MS_DEC && p.precision == 64) h->fast = plan_tasp_fast(h->g, decoder_id, h->smem_per_sm, h->smem_per_block);    // double: tasp_fast.cu
            else if (decoder_id == LDPCB200_MS_DEC) h->fast = plan_ms_fast(h->g, 1, p.precision, h->smem_per_sm, h->smem_per_block, p.use_fast >= 2, h->dp);
            else if (decoder_id == LDPCB200_BP_DEC || decoder_id == LDPCB200_SP_DEC) {
                h->fast = plan_bpsp4(h->g, decoder_id, h->smem_per_sm, h->smem_per_block);                     // four threads per check row
                if (!h->fast.ok) h->fast = plan_tasp_fast(h->g, decoder_id, h->smem_per_sm, h->smem_per_block);
            }
            else if (decoder_id == LDPCB200_TASP_DEC || decoder_id == LDPCB200_ASP_DEC || decoder_id == LDPCB200_LCHE_DEC || decoder_id == LDPCB200_IASP_DEC)
                h->fast = plan_tasp_fast(h->g, decoder_id, h->smem_per_sm, h->smem_per_block);
        }
        return 0;
    }();
    if (rc) { ldpcb200_destroy(h); return rc; }
    *out = h;
    return 0;
}

int ldpcb200_destroy(ldpcb200_handle h)
{
    if (!h) return 0;
    DeviceGuard guard(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->s_in) cudaStreamSynchronize(h->s_in);
    if (h->s_out) cudaStreamSynchronize(h->s_out);
    for (Slot& s : h->slot) {
        s.llr.release(); s.words.release(); s.bytes.release(); s.iters.release(); s.post.release(); s.aux.release(); s.perframe.release();
        if (s.ev_in) cudaEventDestroy(s.ev_in);
        if (s.ev_k) cudaEventDestroy(s.ev_k);
        if (s.ev_out) cudaEventDestroy(s.ev_out);
    }
    h->tables.release(); h->ws.release(); h->counters.release(); h->next.release(); h->bpsynd.release(); h->coef.release(); h->perm.release(); h->cw.release(); h->cw_words.release();
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->s_in) cudaStreamDestroy(h->s_in);
    if (h->s_out) cudaStreamDestroy(h->s_out);
    delete h;
    return 0;
}

int ldpcb200_info(ldpcb200_handle h, int* N, int* R, int* E, int* device)
{
    if (!h) return fail(LDPCB200_EINVAL, "null handle");
    if (N) *N = h->g.N;
    if (R) *R = h->g.R;
    if (E) *E = h->g.E;
    if (device) *device = h->device;
    return 0;
}

int ldpcb200_kernel_info(ldpcb200_handle h, int* fast, int* threads, int* frames_per_cta, int* ctas_per_sm, int* smem_bytes)
{
    if (!h) return fail(LDPCB200_EINVAL, "null handle");
    if (fast) *fast = h->fast.ok ? (1 + h->fast.variant) | (h->fast.tmem ? 16 : 0) | (h->fast.tmem == 2 ? 32 : 0) : 0;
    if (threads) *threads = h->fast.ok ? h->fast.threads : h->nt;
    if (frames_per_cta) *frames_per_cta = h->fast.ok ? h->fast.frames_per_cta : 1;
    if (ctas_per_sm) *ctas_per_sm = h->fast.ok ? h->fast.ctas_per_sm : h->grid / std::max(h->num_sms, 1);
    if (smem_bytes) *smem_bytes = h->fast.ok ? (int)h->fast.smem_bytes : (int)h->smem_ws;
    return 0;
}

int ldpcb200_decode_batch(ldpcb200_handle h, const void* llr, int llr_dtype, int n_frames, int maxiter,
                          uint32_t flags, void* hard, int32_t* iters, void* posterior, int post_dtype, void* aux)
{
    if (!h) return fail(LDPCB200_EINVAL, "null handle");
    if (n_frames < 0 || maxiter < 0) return fail(LDPCB200_EINVAL, "negative n_frames or maxiter");
    if (n_frames == 0) return 0;
    if (!llr) return fail(LDPCB200_EINVAL, "llr is null");
    if (llr_dtype != LDPCB200_F64 && llr_dtype != LDPCB200_F32) return fail(LDPCB200_EINVAL, "llr_dtype must be F64 or F32");
    if (posterior) {
        bool intdec = h->decoder_id == LDPCB200_IMS_DEC || h->decoder_id == LDPCB200_IASP_DEC;
        bool ok = intdec ? post_dtype == default_post_dtype(h->decoder_id, 64) : (post_dtype == LDPCB200_F64 || post_dtype == LDPCB200_F32);
        if (!ok) return fail(LDPCB200_EINVAL, "post_dtype %d does not fit decoder %d", post_dtype, h->decoder_id);
    }
    if (aux && h->decoder_id != LDPCB200_IMS_DEC) return fail(LDPCB200_EINVAL, "aux is an IMS_DEC output");
    DeviceGuard guard(h->device);
    const int N = h->g.N, nwords = h->gd.nwords;
    const bool in_dev = flags & LDPCB200_LLR_ON_DEVICE, out_dev = flags & LDPCB200_OUT_ON_DEVICE;
    const bool packed = flags & LDPCB200_HARD_PACKED;
    const size_t esz = dtype_size(llr_dtype), psz = dtype_size(post_dtype);
    h->last_launches = 0;

    FrameIO io;
    memset(&io, 0, sizeof io);
    io.llr_dtype = llr_dtype; io.maxiter = maxiter; io.flags = flags; io.post_dtype = post_dtype;

    if (in_dev && out_dev) {
        // everything is resident: one launch over the whole batch
        io.llr = llr; io.nf = n_frames; io.iters = iters; io.post = posterior; io.aux = (int16_t*)aux;
        Slot& s = h->slot[0];
        if (hard) {
            if (packed) io.hard_words = (uint32_t*)hard;
            else { CU(s.words.reserve((size_t)n_frames * nwords * 4)); io.hard_words = (uint32_t*)s.words.p; }
        }
        CU(cudaEventRecord(h->ev0, h->stream));
        int rc = launch_decoder(h, io);
        if (rc) return rc;
        CU(cudaEventRecord(h->ev1, h->stream));
        if (hard && !packed) CU(launch_unpack_hard(io.hard_words, (uint8_t*)hard, n_frames, N, nwords, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        CU(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
        return 0;
    }

    // host buffers on at least one side: three-stage pipeline (H2D | decode | D2H) over chunks of
    // frames, two slots, copies on their own streams.  Whatever way this function is left, no copy into a caller's
    // buffer or a slot may still be in flight: the guard drains the three streams on the error paths too.
    struct Drain {
        ldpcb200_handle_s* h;
        bool armed;
        ~Drain()
        {
            if (!armed) return;
            cudaStreamSynchronize(h->s_in); cudaStreamSynchronize(h->stream); cudaStreamSynchronize(h->s_out);
        }
    } drain{h, true};
    size_t per_frame_bytes = (size_t)N * esz + (size_t)nwords * 4 + (hard && !packed ? N : 0) + (posterior ? N * psz : 0) + (aux ? N * 2 : 0) + 8;
    int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)n_frames, ((size_t)192 << 20) / per_frame_bytes));
    if (chunk > 64) chunk &= ~63;
    CU(cudaEventRecord(h->ev0, h->stream));
    int nchunks = (n_frames + chunk - 1) / chunk;
    for (int ci = 0; ci < nchunks; ci++) {
        Slot& s = h->slot[ci & 1];
        const int f0 = ci * chunk, nf = std::min(chunk, n_frames - f0);
        io.nf = nf;
        // input
        if (in_dev) io.llr = (const char*)llr + (size_t)f0 * N * esz;
        else {
            CU(s.llr.reserve((size_t)chunk * N * esz));
            CU(cudaStreamWaitEvent(h->s_in, s.ev_k, 0));          // the decode two chunks ago has consumed this slot
            CU(cudaMemcpyAsync(s.llr.p, (const char*)llr + (size_t)f0 * N * esz, (size_t)nf * N * esz, cudaMemcpyHostToDevice, h->s_in));
            CU(cudaEventRecord(s.ev_in, h->s_in));
            CU(cudaStreamWaitEvent(h->stream, s.ev_in, 0));
            io.llr = s.llr.p;
        }
        // outputs
        CU(cudaStreamWaitEvent(h->stream, s.ev_out, 0));          // the D2H two chunks ago has drained this slot
        io.hard_words = nullptr; io.iters = nullptr; io.post = nullptr; io.aux = nullptr;
        if (hard) {
            if (out_dev && packed) io.hard_words = (uint32_t*)hard + (size_t)f0 * nwords;
            else { CU(s.words.reserve((size_t)chunk * nwords * 4)); io.hard_words = (uint32_t*)s.words.p; }
        }
        if (iters) { if (out_dev) io.iters = iters + f0; else { CU(s.iters.reserve((size_t)chunk * 4)); io.iters = (int32_t*)s.iters.p; } }
        if (posterior) { if (out_dev) io.post = (char*)posterior + (size_t)f0 * N * psz; else { CU(s.post.reserve((size_t)chunk * N * psz)); io.post = s.post.p; } }
        if (aux) { if (out_dev) io.aux = (int16_t*)aux + (size_t)f0 * N; else { CU(s.aux.reserve((size_t)chunk * N * 2)); io.aux = (int16_t*)s.aux.p; } }
        int rc = launch_decoder(h, io);
        if (rc) return rc;
        uint8_t* dbytes = nullptr;
        if (hard && !packed) {
            if (out_dev) dbytes = (uint8_t*)hard + (size_t)f0 * N;
            else { CU(s.bytes.reserve((size_t)chunk * N)); dbytes = (uint8_t*)s.bytes.p; }
            CU(launch_unpack_hard(io.hard_words, dbytes, nf, N, nwords, h->stream));
        }
        CU(cudaEventRecord(s.ev_k, h->stream));
        if (!out_dev) {
            CU(cudaStreamWaitEvent(h->s_out, s.ev_k, 0));
            if (hard) {
                if (packed) CU(cudaMemcpyAsync((uint32_t*)hard + (size_t)f0 * nwords, io.hard_words, (size_t)nf * nwords * 4, cudaMemcpyDeviceToHost, h->s_out));
                else CU(cudaMemcpyAsync((uint8_t*)hard + (size_t)f0 * N, dbytes, (size_t)nf * N, cudaMemcpyDeviceToHost, h->s_out));
            }
            if (iters) CU(cudaMemcpyAsync(iters + f0, io.iters, (size_t)nf * 4, cudaMemcpyDeviceToHost, h->s_out));
            if (posterior) CU(cudaMemcpyAsync((char*)posterior + (size_t)f0 * N * psz, io.post, (size_t)nf * N * psz, cudaMemcpyDeviceToHost, h->s_out));
            if (aux) CU(cudaMemcpyAsync((int16_t*)aux + (size_t)f0 * N, io.aux, (size_t)nf * N * 2, cudaMemcpyDeviceToHost, h->s_out));
            CU(cudaEventRecord(s.ev_out, h->s_out));
        }
    }
    CU(cudaEventRecord(h->ev1, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaStreamSynchronize(h->s_out));
    CU(cudaStreamSynchronize(h->s_in));
    drain.armed = false;
    CU(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
    return 0;
}

int ldpcb200_simulate(ldpcb200_handle h, const ldpcb200_sim_params* sp, ldpcb200_counters* out, uint32_t* per_frame)
{
    int rc = check_sim(h, sp);
    if (rc) return rc;
    if (!out) return fail(LDPCB200_EINVAL, "out is null");
    memset(out, 0, sizeof *out);
    if (sp->n_frames == 0) return 0;
    if (sp->n_frames > 0x7fffffffu) return fail(LDPCB200_EINVAL, "n_frames too large for one round");
    DeviceGuard guard(h->device);
    h->last_launches = 0;
    FrameIO io;
    memset(&io, 0, sizeof io);
    io.nf = (int)sp->n_frames; io.maxiter = sp->max_iterations; io.flags = sp->flags & LDPCB200_NO_EARLY_EXIT;
    io.llr_dtype = LDPCB200_F32; io.post_dtype = default_post_dtype(h->decoder_id, h->p.precision);
    fill_channel(h, sp, io.ch);
    io.counters = (unsigned long long*)h->counters.p;
    const bool pf_dev = sp->flags & LDPCB200_OUT_ON_DEVICE;
    Slot& s = h->slot[0];
    if (per_frame) {
        if (pf_dev) io.per_frame = per_frame;
        else { CU(s.perframe.reserve((size_t)io.nf * 4)); io.per_frame = (uint32_t*)s.perframe.p; }
    }
    CU(cudaMemsetAsync(h->counters.p, 0, 8 * sizeof(unsigned long long), h->stream));
    CU(cudaEventRecord(h->ev0, h->stream));
    if (h->has_cw) {
        // A transmitted codeword other than all-zero: the fused first loads and error counters of the decode kernels are
        // for the all-zero one, so the round is three launches per chunk -- LLRs of the codeword through the channel (per-bit
        // path, interleaver included), decode from that buffer into packed decisions, decisions against the codeword.
        const int N = h->g.N, nwords = h->gd.nwords;
        const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)io.nf, ((size_t)512 << 20) / ((size_t)N * 4)));
        CU(s.llr.reserve((size_t)chunk * N * 4));
        CU(s.words.reserve((size_t)chunk * nwords * 4));
        CU(s.iters.reserve((size_t)chunk * 4));
        for (int f0 = 0; f0 < io.nf; f0 += chunk) {
            const int nf = std::min(chunk, io.nf - f0);
            ChannelParams ch = io.ch;
            ch.first_frame = io.ch.first_frame + (unsigned long long)f0;
            CU(launch_generate_llr(ch, N, nf, s.llr.p, LDPCB200_F32, h->stream));
            FrameIO d;
            memset(&d, 0, sizeof d);
            d.llr = s.llr.p; d.llr_dtype = LDPCB200_F32; d.nf = nf; d.maxiter = io.maxiter; d.flags = io.flags;
            d.post_dtype = io.post_dtype; d.hard_words = (uint32_t*)s.words.p; d.iters = (int32_t*)s.iters.p;
            rc = launch_decoder(h, d);
            if (rc) return rc;
            CU(launch_count_errors((const uint32_t*)s.words.p, (const uint32_t*)h->cw_words.p, (const int*)s.iters.p, nf, N, h->g.R, nwords,
                                   (unsigned long long*)h->counters.p, io.per_frame ? io.per_frame + f0 : nullptr, h->stream));
            h->last_launches += 2;
        }
    } else {
        rc = launch_decoder(h, io);
        if (rc) return rc;
    }
    CU(cudaEventRecord(h->ev1, h->stream));
    unsigned long long c[6];
    CU(cudaMemcpyAsync(c, h->counters.p, sizeof c, cudaMemcpyDeviceToHost, h->stream));
    if (per_frame && !pf_dev) CU(cudaMemcpyAsync(per_frame, io.per_frame, (size_t)io.nf * 4, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
    out->frames = c[0]; out->frame_errors = c[1]; out->info_bit_errors = c[2]; out->undetected = c[3];
    out->iter_sum = c[4]; out->bit_errors = c[5];
    return 0;
}

int ldpcb200_simulate_codes(ldpcb200_handle h, int n_codes, const int16_t* hds, const ldpcb200_sim_params* sp,
                            ldpcb200_counters* out, uint32_t* per_frame)
{
    int rc = check_sim(h, sp);
    if (rc) return rc;
    if (n_codes < 0 || (n_codes > 0 && (!hds || !out))) return fail(LDPCB200_EINVAL, "bad argument");
    if (n_codes == 0) return 0;
    if (n_codes > 65535) return fail(LDPCB200_EINVAL, "at most 65535 codes per call (the grid's second dimension)");
    memset(out, 0, sizeof(*out) * (size_t)n_codes);
    if (sp->n_frames == 0) return 0;
    if (sp->n_frames > 0x7fffffffu) return fail(LDPCB200_EINVAL, "n_frames too large for one round");
    if (sp->flags & LDPCB200_OUT_ON_DEVICE) return fail(LDPCB200_EINVAL, "per_frame must be a host buffer here");
    if (h->has_cw) return fail(LDPCB200_EUNSUPPORTED, "simulate_codes sends the all-zero codeword (a codeword belongs to one matrix)");
    if (h->decoder_id != LDPCB200_TASP_DEC || !h->fast.ok)
        return fail(LDPCB200_EUNSUPPORTED, "simulate_codes needs a TASP_DEC handle on the tensor-memory kernel (use_fast >= 1, messages fitting tensor memory)");
    DeviceGuard guard(h->device);
    const QcHost& g0 = h->g;
    const int b = g0.b, c = g0.c, Z = g0.Z;
    // tables of all codes in one blob: per code rp (b + 1) | col (E) | sh (E), each padded to 4 ints
    auto pad4 = [](size_t n) { return (n + 3) & ~(size_t)3; };
    const size_t per_code = pad4(b + 1) + 2 * pad4(g0.E);
    std::vector<int> blob(per_code * (size_t)n_codes, 0);
    for (int k = 0; k < n_codes; k++) {
        QcHost g;
        if (!g.build(hds + (size_t)k * b * c, b, c, Z)) return fail(LDPCB200_EINVAL, "code %d: bad base matrix", k);
        if (g.E != g0.E) return fail(LDPCB200_EUNSUPPORTED, "code %d has %d circulants, the handle's shape has %d", k, g.E, g0.E);
        if (g.maxdeg > LDPCB200_MAX_ROW_WEIGHT || g.maxdeg > 20 || g.mindeg < 2) return fail(LDPCB200_EUNSUPPORTED, "code %d: row weights must be 2 .. 20", k);
        int* base = blob.data() + per_code * (size_t)k;
        std::copy(g.rp.begin(), g.rp.end(), base);
        std::copy(g.col.begin(), g.col.end(), base + pad4(b + 1));
        std::copy(g.sh.begin(), g.sh.end(), base + pad4(b + 1) + pad4(g0.E));
    }
    const size_t nf = sp->n_frames;
    DevBuf d_blob, d_gs, d_ios, d_cnt, d_pf;
    struct Free { DevBuf *a, *b, *c, *d, *e; ~Free() { a->release(); b->release(); c->release(); d->release(); e->release(); } } fr{&d_blob, &d_gs, &d_ios, &d_cnt, &d_pf};
    CU(d_blob.reserve(blob.size() * sizeof(int)));
    CU(d_gs.reserve(sizeof(QcDev) * (size_t)n_codes));
    CU(d_ios.reserve(sizeof(FrameIO) * (size_t)n_codes));
    CU(d_cnt.reserve(sizeof(unsigned long long) * 8 * (size_t)n_codes));              // per code: 6 counters | frame ticket | pad
    if (per_frame) CU(d_pf.reserve(sizeof(uint32_t) * nf * (size_t)n_codes));
    CU(cudaMemcpyAsync(d_blob.p, blob.data(), blob.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemsetAsync(d_cnt.p, 0, sizeof(unsigned long long) * 8 * (size_t)n_codes, h->stream));
    std::vector<QcDev> gs((size_t)n_codes);
    std::vector<FrameIO> ios((size_t)n_codes);
    FrameIO io;
    memset(&io, 0, sizeof io);
    io.nf = (int)nf; io.maxiter = sp->max_iterations; io.flags = sp->flags & LDPCB200_NO_EARLY_EXIT;
    io.llr_dtype = LDPCB200_F32; io.post_dtype = LDPCB200_F64;
    fill_channel(h, sp, io.ch);
    for (int k = 0; k < n_codes; k++) {
        const int* base = (const int*)d_blob.p + per_code * (size_t)k;
        QcDev d = h->gd;                                                               // geometry of the shape; the edge lists of code k
        d.rp = base; d.col = base + pad4(b + 1); d.sh = base + pad4(b + 1) + pad4(g0.E);
        d.row = nullptr; d.cp = nullptr; d.cedge = nullptr; d.pk = nullptr;            // (the TASP kernel reads rp, col, sh only)
        gs[(size_t)k] = d;
        FrameIO x = io;
        unsigned long long* cnt = (unsigned long long*)d_cnt.p + 8 * (size_t)k;
        x.counters = cnt;
        x.next_frame = (unsigned int*)(cnt + 6);
        x.per_frame = per_frame ? (uint32_t*)d_pf.p + nf * (size_t)k : nullptr;
        ios[(size_t)k] = x;
    }
    CU(cudaMemcpyAsync(d_gs.p, gs.data(), sizeof(QcDev) * (size_t)n_codes, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(d_ios.p, ios.data(), sizeof(FrameIO) * (size_t)n_codes, cudaMemcpyHostToDevice, h->stream));
    // CTAs per code: enough to fill the SMs over all codes, never more than frames
    int grid_x = (h->num_sms * h->fast.ctas_per_sm + n_codes - 1) / n_codes;
    grid_x = (int)std::max<size_t>(1, std::min<size_t>((size_t)grid_x, nf));
    h->last_launches = 0;
    CU(cudaEventRecord(h->ev0, h->stream));
    CU(launch_tasp_multi(h->fast, (const QcDev*)d_gs.p, (const FrameIO*)d_ios.p, n_codes, grid_x, h->stream));
    h->last_launches++;
    CU(cudaEventRecord(h->ev1, h->stream));
    std::vector<unsigned long long> cnt(8 * (size_t)n_codes);
    CU(cudaMemcpyAsync(cnt.data(), d_cnt.p, sizeof(unsigned long long) * cnt.size(), cudaMemcpyDeviceToHost, h->stream));
    if (per_frame) CU(cudaMemcpyAsync(per_frame, d_pf.p, sizeof(uint32_t) * nf * (size_t)n_codes, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
    for (int k = 0; k < n_codes; k++) {
        const unsigned long long* x = cnt.data() + 8 * (size_t)k;
        out[k].frames = x[0]; out[k].frame_errors = x[1]; out[k].info_bit_errors = x[2]; out[k].undetected = x[3];
        out[k].iter_sum = x[4]; out[k].bit_errors = x[5];
    }
    return 0;
}

int ldpcb200_generate_llr(ldpcb200_handle h, const ldpcb200_sim_params* sp, void* llr, int llr_dtype)
{
    int rc = check_sim(h, sp);
    if (rc) return rc;
    if (!llr) return fail(LDPCB200_EINVAL, "llr is null");
    if (llr_dtype != LDPCB200_F64 && llr_dtype != LDPCB200_F32) return fail(LDPCB200_EINVAL, "llr_dtype must be F64 or F32");
    if (sp->n_frames == 0) return 0;
    DeviceGuard guard(h->device);
    ChannelParams ch;
    fill_channel(h, sp, ch);
    const int N = h->g.N;
    const size_t esz = dtype_size(llr_dtype);
    if (sp->flags & LDPCB200_OUT_ON_DEVICE) {
        CU(launch_generate_llr(ch, N, (int)sp->n_frames, llr, llr_dtype, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        return 0;
    }
    Slot& s = h->slot[0];
    const uint32_t chunk = (uint32_t)std::max<size_t>(1, ((size_t)256 << 20) / ((size_t)N * esz));
    CU(s.llr.reserve((size_t)std::min(chunk, sp->n_frames) * N * esz));
    for (uint32_t f0 = 0; f0 < sp->n_frames; f0 += chunk) {
        uint32_t nf = std::min(chunk, sp->n_frames - f0);
        ChannelParams c2 = ch;
        c2.first_frame = ch.first_frame + f0;
        CU(launch_generate_llr(c2, N, (int)nf, s.llr.p, llr_dtype, h->stream));
        CU(cudaMemcpyAsync((char*)llr + (size_t)f0 * N * esz, s.llr.p, (size_t)nf * N * esz, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    }
    return 0;
}

int ldpcb200_set_codeword(ldpcb200_handle h, const uint8_t* bits)
{
    if (!h) return fail(LDPCB200_EINVAL, "null handle");
    if (!bits) { h->has_cw = false; return 0; }
    const int N = h->g.N, nwords = h->gd.nwords;
    std::vector<uint32_t> words((size_t)nwords, 0u);
    bool any = false;
    for (int i = 0; i < N; i++) {
        if (bits[i] > 1) return fail(LDPCB200_EINVAL, "codeword bit %d is %d (must be 0 or 1)", i, bits[i]);
        if (bits[i]) { words[(size_t)(i >> 5)] |= 1u << (i & 31); any = true; }
    }
    // (that the word satisfies the parity checks is the caller's business: host/encoder.cpp qc_encode / random_codeword make them)
    if (!any) { h->has_cw = false; return 0; }                     // all-zero: the fused path
    DeviceGuard guard(h->device);
    CU(cudaStreamSynchronize(h->stream));
    CU(h->cw.reserve((size_t)N));
    CU(h->cw_words.reserve((size_t)nwords * 4));
    CU(cudaMemcpy(h->cw.p, bits, (size_t)N, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(h->cw_words.p, words.data(), (size_t)nwords * 4, cudaMemcpyHostToDevice));
    h->has_cw = true;
    return 0;
}

int ldpcb200_generate_noise(ldpcb200_handle h, const ldpcb200_sim_params* sp, int n_samples, float* out)
{
    int rc = check_sim(h, sp);
    if (rc) return rc;
    if (!out || n_samples <= 0) return fail(LDPCB200_EINVAL, "bad argument");
    if (sp->n_frames == 0) return 0;
    DeviceGuard guard(h->device);
    ChannelParams ch;
    fill_channel(h, sp, ch);
    const size_t total = (size_t)sp->n_frames * (size_t)n_samples;
    Slot& s = h->slot[0];
    CU(s.llr.reserve(total * 4));
    CU(launch_generate_noise(ch, n_samples, (int)sp->n_frames, (float*)s.llr.p, h->stream));
    CU(cudaMemcpyAsync(out, s.llr.p, total * 4, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

int ldpcb200_set_interleaver(ldpcb200_handle h, const int32_t* direct, const int32_t* inverse)
{
    if (!h) return fail(LDPCB200_EINVAL, "null handle");
    if (!direct && !inverse) { h->has_perm = false; return 0; }
    if (!direct || !inverse) return fail(LDPCB200_EINVAL, "direct and inverse must both be given (or both NULL)");
    const int N = h->g.N;
    for (int i = 0; i < N; i++)
        if (inverse[i] < 0 || inverse[i] >= N || direct[inverse[i]] != i) return fail(LDPCB200_EINVAL, "direct / inverse are not mutually inverse permutations (entry %d)", i);
    for (int j = 0; j < N; j++)
        if (direct[j] < 0 || direct[j] >= N || inverse[direct[j]] != j) return fail(LDPCB200_EINVAL, "direct / inverse are not mutually inverse permutations (entry %d)", j);
    DeviceGuard guard(h->device);
    CU(cudaStreamSynchronize(h->stream));
    CU(h->perm.reserve(sizeof(int32_t) * 2 * (size_t)N));
    CU(cudaMemcpy(h->perm.p, direct, sizeof(int32_t) * N, cudaMemcpyHostToDevice));
    CU(cudaMemcpy((int32_t*)h->perm.p + N, inverse, sizeof(int32_t) * N, cudaMemcpyHostToDevice));
    h->has_perm = true;
    return 0;
}

static int pick_device(int device, int* dev)
{
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(LDPCB200_ENODEV, "no CUDA device (there is no CPU fallback)"); }
    *dev = device;
    if (device < 0) { if (cudaGetDevice(dev) != cudaSuccess) return fail(LDPCB200_ECUDA, "cudaGetDevice failed"); }
    if (*dev >= ndev) return fail(LDPCB200_EINVAL, "device %d out of range", *dev);
    return 0;
}

int ldpcb200_demodulate(int Q, int ns, double sigma, double T, int out_type, const double* x, double* res, int device)
{
    if (Q != 4 && Q != 16 && Q != 64 && Q != 256) return fail(LDPCB200_EINVAL, "Q must be 4, 16, 64 or 256");
    if (ns < 0 || !x || !res || !(sigma > 0)) return fail(LDPCB200_EINVAL, "bad argument");
    if (Q == 4 && out_type != 0) return fail(LDPCB200_EUNSUPPORTED, "QAM-4 probability output (a sum over the whole frame, QAM_demodulator.cpp:124-139) is not provided");
    if (ns == 0) return 0;
    int dev;
    int rc = pick_device(device, &dev);
    if (rc) return rc;
    DeviceGuard guard(dev);
    int m = Q == 4 ? 2 : Q == 16 ? 4 : Q == 64 ? 6 : 8;
    double *dx = nullptr, *dr = nullptr;
    CU(cudaMalloc(&dx, sizeof(double) * 2 * ns));
    cudaError_t e = cudaMalloc(&dr, sizeof(double) * (size_t)ns * m);
    if (e != cudaSuccess) { cudaFree(dx); return fail(LDPCB200_ENOMEM, "cudaMalloc failed"); }
    e = cudaMemcpy(dx, x, sizeof(double) * 2 * ns, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = launch_demodulate(m, ns, sigma, T, out_type, dx, dr, 0);
    if (e == cudaSuccess) e = cudaMemcpy(res, dr, sizeof(double) * (size_t)ns * m, cudaMemcpyDeviceToHost);
    cudaFree(dx); cudaFree(dr);
    if (e != cudaSuccess) return fail(LDPCB200_ECUDA, "demodulate: %s", cudaGetErrorString(e));
    return 0;
}

int ldpcb200_modulate(int Q, int ns, const uint8_t* bits, double* out, int device)
{
    if (Q != 4 && Q != 16 && Q != 64 && Q != 256) return fail(LDPCB200_EINVAL, "Q must be 4, 16, 64 or 256");
    if (ns < 0 || !bits || !out) return fail(LDPCB200_EINVAL, "bad argument");
    if (ns == 0) return 0;
    int dev;
    int rc = pick_device(device, &dev);
    if (rc) return rc;
    DeviceGuard guard(dev);
    int m = Q == 4 ? 2 : Q == 16 ? 4 : Q == 64 ? 6 : 8;
    uint8_t* db = nullptr; double* dout = nullptr;
    CU(cudaMalloc(&db, (size_t)ns * m));
    cudaError_t e = cudaMalloc(&dout, sizeof(double) * 2 * ns);
    if (e != cudaSuccess) { cudaFree(db); return fail(LDPCB200_ENOMEM, "cudaMalloc failed"); }
    e = cudaMemcpy(db, bits, (size_t)ns * m, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = launch_modulate(m, ns, db, dout, 0);
    if (e == cudaSuccess) e = cudaMemcpy(out, dout, sizeof(double) * 2 * ns, cudaMemcpyDeviceToHost);
    cudaFree(db); cudaFree(dout);
    if (e != cudaSuccess) return fail(LDPCB200_ECUDA, "modulate: %s", cudaGetErrorString(e));
    return 0;
}

int ldpcb200_jit_check(const int16_t* hd, int b, int c, int Z, int sm_major, int sm_minor, int* cubin_bytes)
{
    QcHost g;
    if (!g.build(hd, b, c, Z)) return fail(LDPCB200_EINVAL, "bad base matrix");
    int zp, minb, variant;
    size_t smem;
    const char* no_tmem = getenv("LDPCB200_NO_TMEM");
    if (!lms_spec_geometry(g, 233472, 232448, &zp, &minb, &smem, &variant, !(no_tmem && *no_tmem == '1'))) return fail(LDPCB200_EUNSUPPORTED, "code does not suit the code-specialised kernel");
    std::vector<char> cubin;
    std::string why;
    int kind = 0;
    if (const char* e = getenv("LDPCB200_JIT_CHECK"))           // development: "kind,variant,minb" of another kernel family (spec_jit.cpp)
        sscanf(e, "%d,%d,%d", &kind, &variant, &minb);
    if (!lms_spec_compile(lms_spec_generate(g, zp, minb, variant, kind), sm_major, sm_minor, cubin, why)) return fail(LDPCB200_EUNSUPPORTED, "%s", why.c_str());
    if (cubin_bytes) *cubin_bytes = (int)cubin.size();
    return 0;
}

int ldpcb200_last_kernel_ms(ldpcb200_handle h, float* ms, int* launches)
{
    if (!h) return fail(LDPCB200_EINVAL, "null handle");
    if (ms) *ms = h->last_ms;
    if (launches) *launches = h->last_launches;
    return 0;
}

void* ldpcb200_stream(ldpcb200_handle h) { return h ? (void*)h->stream : nullptr; }

} // extern "C"
