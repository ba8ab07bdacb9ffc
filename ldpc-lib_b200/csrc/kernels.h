// Internal interface between the C-ABI layer (api.cpp) and the CUDA translation units.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include <cuda_runtime.h>

#include "../../include/ldpcb200.h"
#include "qc_layout.h"
#include "frame_io.h"

namespace ldpcb200 {

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
    int variant = 0;                    // 0 table-driven (lms_fast.cu), 1 code-specialised ahead of time, 2 run-time compiled
    int spec_index = -1;
    int tmem = 0;                       // 1: LMS_DEC with the c2v messages in tensor memory (lms_tmem.cuh)
    int msg32 = 0;                      // BP_DEC / SP_DEC (tasp_fast.cu): messages rounded to fp32, one tensor-memory column each
    int bpsp4 = 0;                      // BP_DEC / SP_DEC: four threads per check row (bpsp4.cu)
    const void* jit_kernel = nullptr;
    std::string note;                   // why a faster variant was not used
    std::vector<unsigned char> tab;     // the kernel's parameter-space copy of the edge lists
};
FastPlan plan_lms_fast(const QcHost& g, int precision, int smem_per_sm, int smem_per_block, int allow_jit);
cudaError_t launch_lms_fast(const FastPlan& p, const FrameIO& io, int grid, cudaStream_t s);
// flooding min-sum pair (ims_fast.cu -> ms_spec.cuh): kind 1 = MS_DEC fp32, 2 = IMS_DEC
FastPlan plan_ms_fast(const QcHost& g, int kind, int precision, int smem_per_sm, int smem_per_block, int allow_jit, const DecParams& dp);
cudaError_t launch_ms_fast(const FastPlan& p, const DecParams& dp, const FrameIO& io, int grid, cudaStream_t s);

// TASP_DEC / ASP_DEC in double with the messages in tensor memory (tasp_fast.cu); table-driven, any code that fits
FastPlan plan_tasp_fast(const QcHost& g, int decoder_id, int smem_per_sm, int smem_per_block);
cudaError_t launch_tasp_fast(const FastPlan& p, int decoder_id, const QcDev& g, const FrameIO& io, int grid, cudaStream_t s, double alpha = 0.0);
// BP_DEC / SP_DEC with four threads per check row (bpsp4.cu)
FastPlan plan_bpsp4(const QcHost& g, int decoder_id, int smem_per_sm, int smem_per_block);
cudaError_t launch_bpsp4(const FastPlan& p, int decoder_id, const QcDev& g, const FrameIO& io, int grid, cudaStream_t s);
// TASP_DEC on n_codes matrices of the handle's shape in one launch: grid (grid_x, n_codes), d_gs / d_ios device arrays
cudaError_t launch_tasp_multi(const FastPlan& p, const QcDev* d_gs, const FrameIO* d_ios, int n_codes, int grid_x, cudaStream_t s);

// ---- utilities (channel.cu)
// packed words -> one byte per bit
cudaError_t launch_unpack_hard(const uint32_t* words, uint8_t* bytes, int nf, int N, int nwords, cudaStream_t s);
// channel LLRs for frames [first_frame, first_frame + nf) written as F32 or F64
cudaError_t launch_generate_llr(const ChannelParams& ch, int N, int nf, void* llr, int llr_dtype, cudaStream_t s);
// error accounting against a transmitted codeword (packed decisions XOR packed codeword), and the raw noise samples (tests)
cudaError_t launch_count_errors(const uint32_t* hard_words, const uint32_t* cw_words, const int* iters, int nf, int N, int R, int nwords,
                                unsigned long long* counters, uint32_t* per_frame, cudaStream_t s);
cudaError_t launch_generate_noise(const ChannelParams& ch, int ns, int nf, float* out, cudaStream_t s);
// IMS_DEC quantiser pre-pass: coef[f] = sqrt(N / sum y^2), the sum in the reference's sequential order, one lane per frame
cudaError_t launch_ims_energy(const FrameIO& io, int N, double* coef, cudaStream_t s);
// Demodulate / QAM_modulator at the function boundary
cudaError_t launch_demodulate(int m, int ns, double sigma, double T, int out_type, const double* x, double* res, cudaStream_t s);
cudaError_t launch_modulate(int m, int ns, const uint8_t* bits, double* out, cudaStream_t s);

} // namespace ldpcb200
