// Internal interface between the C-ABI layer (api.cpp) and the CUDA translation units.
#pragma once
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#include "../../include/ldpcb200.h"
#include "qc_layout.h"

namespace ldpcb200 {

// Channel description for on-device LLR generation (bp_simulation.cpp:444-449, 600-630).
struct ChannelParams {
    int enabled;                // 0: LLRs come from FrameIO::llr
    int modulation;             // enum ldpcb200_modulation
    int m;                      // bits per QAM symbol (1 for BPSK)
    float sigma;                // sigma (BPSK) or sigmaQAM
    float llr_scale;            // 2 / sigma^2
    double sigma_d;             // the same in double for the Demodulate arithmetic
    double T;                   // Demodulate clip
    int punct_start;            // first punctured bit (N if none)
    float punct_value;          // 0.5 for LLR-domain decoders, 0 otherwise (bp_simulation.cpp:700)
    unsigned long long seed;
    unsigned int stream;
    unsigned long long first_frame;
};

// Everything one decode launch reads and writes (all pointers are device pointers).
struct FrameIO {
    const void* llr;            // nf*N values of llr_dtype (F64 | F32); unused when ch.enabled
    int llr_dtype;
    int nf;
    int maxiter;
    uint32_t flags;
    uint32_t* hard_words;       // nf * nwords packed decisions (may be null)
    int32_t* iters;             // nf (may be null)
    void* post;                 // nf*N of post_dtype (may be null)
    int post_dtype;
    int16_t* aux;               // IMS: ims_y (may be null)
    uint32_t* per_frame;        // nf error records (may be null)
    unsigned long long* counters;   // 6 x u64 (may be null): frames, frame_errors, info_bit_errors,
                                    // undetected, iter_sum, bit_errors
    uint8_t* bp_syndrome;       // R bytes: BP_DEC chained syndrome (decoders.cpp:1742-1759), or null
    unsigned int* next_frame;   // work counter of the persistent grid (zeroed before the launch)
    ChannelParams ch;
};

struct DecParams {
    double alpha, thr;
    int qbits, dbits;
};

// ---- table-driven parity kernels (dec_minsum.cu, dec_sumprod.cu): one frame per CTA at a time,
// state in a per-CTA global workspace slice (L2 resident)
size_t generic_workspace_bytes(int decoder_id, int precision, const QcHost& g, int nthreads);
cudaError_t launch_generic(int decoder_id, int precision, const QcDev& g, const DecParams& dp,
                           const FrameIO& io, char* ws, size_t ws_stride, int grid, int nthreads,
                           cudaStream_t s);

// ---- shared-memory throughput kernels (lms_fast.cu, ims_fast.cu)
struct FastPlan {
    int ok = 0;                 // 0: this (code, decoder, precision) has no fast kernel
    int threads = 0, frames_per_cta = 1, ctas_per_sm = 1;
    size_t smem_bytes = 0;
    int variant = 0;
    std::vector<unsigned char> tab;     // the kernel's parameter-space copy of the edge lists
};
FastPlan plan_lms_fast(const QcHost& g, int precision, int smem_per_sm, int smem_per_block);
cudaError_t launch_lms_fast(const FastPlan& p, const FrameIO& io, int grid, cudaStream_t s);
FastPlan plan_ims_fast(const QcHost& g, const DecParams& dp, int smem_per_sm, int smem_per_block);
cudaError_t launch_ims_fast(const FastPlan& p, const QcDev& g, const DecParams& dp, const FrameIO& io,
                            double* coef, int grid, cudaStream_t s);

// ---- utilities (channel.cu)
// packed words -> one byte per bit
cudaError_t launch_unpack_hard(const uint32_t* words, uint8_t* bytes, int nf, int N, int nwords, cudaStream_t s);
// channel LLRs for frames [first_frame, first_frame + nf) written as F32 or F64
cudaError_t launch_generate_llr(const ChannelParams& ch, int N, int nf, void* llr, int llr_dtype, cudaStream_t s);
// Demodulate / QAM_modulator at the function boundary
cudaError_t launch_demodulate(int m, int ns, double sigma, double T, int out_type, const double* x, double* res, cudaStream_t s);
cudaError_t launch_modulate(int m, int ns, const uint8_t* bits, double* out, cudaStream_t s);

} // namespace ldpcb200
