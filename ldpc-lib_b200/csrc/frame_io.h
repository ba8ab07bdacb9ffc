// Plain-old-data structures shared by the host code, the ahead-of-time kernels and the run-time compiled
// (NVRTC) code-specialised kernels.  No includes and only built-in types, so that NVRTC can compile it as is.
#pragma once

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
    // bit interleaver (direct_inverse_perm.cpp; bp_simulation.cpp:573, :684), device tables of N entries or null for the
    // identity: perm_dir[j] = decoder input fed by transmitted position j, perm_inv[i] = transmitted position of input i
    const int* perm_dir;
    const int* perm_inv;
    // transmitted codeword (bp_simulation.cpp:567-577 sends the all-zero one; SURVEY.md 8f row 2 asks for real ones so that the
    // Gray map is exercised): N bytes 0 / 1 in CODEWORD order on the device, or null for all-zero.  Only the per-bit path
    // (channel_llr) knows it: the fused first loads are for the all-zero codeword and the API routes around them.
    const unsigned char* cw;
    // the factored Demodulate of the in-kernel channel (channel.cuh pam_demod_factored): 2 / N0, 1 / N0, exp(-L^2 / N0) for
    // L = 1, 3, ..., 15; qam_fast = 0: the reference's evaluation order (pam_demod)
    int qam_fast;
    double qam_w, qam_n0inv, qam_c[8];
};

// Everything one decode launch reads and writes (all pointers are device pointers).
struct FrameIO {
    const void* llr;            // nf*N values of llr_dtype (F64 | F32); unused when ch.enabled
    int llr_dtype;
    int nf;
    int maxiter;
    unsigned int flags;
    unsigned int* hard_words;       // nf * nwords packed decisions (may be null)
    int* iters;             // nf (may be null)
    void* post;                 // nf*N of post_dtype (may be null)
    int post_dtype;
    short* aux;               // IMS: ims_y (may be null)
    unsigned int* per_frame;        // nf error records (may be null)
    unsigned long long* counters;   // 6 x u64 (may be null): frames, frame_errors, info_bit_errors,
                                    // undetected, iter_sum, bit_errors
    unsigned char* bp_syndrome;       // R bytes: BP_DEC chained syndrome (decoders.cpp:1742-1759), or null
    const double* coef;         // IMS_DEC: per-frame quantiser scale sqrt(N / sum y^2) from the energy pre-pass (or null)
    unsigned int* next_frame;   // work counter of the persistent grid (zeroed before the launch)
    ChannelParams ch;
};

// run-time parameters of the code-specialised flooding min-sum kernels (ms_spec.cuh)
struct MsSpecParams {
    float alpha;            // MS_DEC: normalisation factor
    int ialpha;             // IMS_DEC: (int)(alpha * 16), decoders.cpp:5458
    int max_data;           // IMS_DEC: 2^(dbits-1) - 1
    int max_quant;          // IMS_DEC: 2^(qbits-1) - 1
    double thr;             // IMS_DEC: quantiser threshold
};

} // namespace ldpcb200
