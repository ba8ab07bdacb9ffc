// On-device channel: counter-based N(0,1) generator, BPSK / QAM-4 LLR and the exact Gray-PAM x PAM
// QAM-16/64/256 bit-LLR.  Device functions shared by the stand-alone kernels in channel.cu and by the
// decoders' first load (the LLRs of a simulated frame never exist in HBM).
//
// Reference behaviour restated here:
//   BPSK / QAM-4 LLR          bp_simulation.cpp:600-612     llr = -2 (sigma n + 2 c - 1) / sigma^2, c = 0
//   received QAM symbol       bp_simulation.cpp:621-625     intended channel r = s + sigmaQAM n (the
//                                                           reference accumulates unscaled noise: SURVEY fact 6)
//   Demodulate                QAM_demodulator.cpp:99-566    with m = log2(Q); caller negates (:627-628)
//   puncturing                bp_simulation.cpp:697-710
// The reference draws its noise from a global std::mt19937 (commons_portable.cpp:174-178); that stream
// cannot be reproduced in parallel, so noise sample i of frame f is instead
//   Philox4x32-10(key = seed, counter = (i / 4, f_lo, f_hi, stream))[i % 4]  ->  Box-Muller (fp32),
// which makes results independent of batch size, launch geometry and GPU count.
#pragma once
#include "frame_io.h"
#include "fastmath64.cuh"

namespace ldpcb200 {

struct u32x4 { unsigned int x, y, z, w; };

__host__ __device__ __forceinline__ u32x4 philox4x32_10(unsigned int c0, unsigned int c1, unsigned int c2, unsigned int c3,
                                                        unsigned int k0, unsigned int k1)
{
    const unsigned int M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int r = 0; r < 10; r++) {
        unsigned long long p0 = (unsigned long long)M0 * c0, p1 = (unsigned long long)M1 * c2;
        unsigned int hi0 = (unsigned int)(p0 >> 32), lo0 = (unsigned int)p0, hi1 = (unsigned int)(p1 >> 32), lo1 = (unsigned int)p1;
        c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    u32x4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

#ifdef __CUDACC__
// two N(0,1) samples from two 32-bit words
__device__ __forceinline__ void box_muller(unsigned int a, unsigned int b, float& z0, float& z1)
{
    // The transcendental steps use the special-function unit directly (lg2 / sqrt / sin / cos .approx: about 1e-6 absolute
    // on a unit-variance sample, far below anything a Monte-Carlo estimate resolves) -- with libm's logf / sincospif the
    // generator was a tenth of a whole decode.  Every path (fused first loads, ldpcb200_generate_llr, the IMS energy
    // pre-pass) shares this function, so buffers handed to the reference are still the values the decoder sees.
    float u = __fmaf_rn(__uint2float_rn(a), 2.3283064365386963e-10f, 1.1641532182693481e-10f);   // (a + 0.5) / 2^32
    float l2, r, s, c;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(__fmul_rn(l2, -1.3862943611198906f)));     // sqrt(-2 ln u), ln u = lg2 u * ln 2
    const float ang = __fmul_rn(__int2float_rn((int)b), 1.4629180792671596e-09f);                 // 2 pi b / 2^32 with b read as signed: [-pi, pi)
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(ang));
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(ang));
    z0 = __fmul_rn(r, c);
    z1 = __fmul_rn(r, s);
}

// noise sample `idx` of frame `frame`
__device__ __forceinline__ float channel_noise(const ChannelParams& ch, unsigned long long frame, unsigned int idx)
{
    u32x4 w = philox4x32_10(idx >> 2, (unsigned int)frame, (unsigned int)(frame >> 32), ch.stream,
                            (unsigned int)ch.seed, (unsigned int)(ch.seed >> 32));
    float z0, z1;
    if (idx & 2) box_muller(w.z, w.w, z0, z1); else box_muller(w.x, w.y, z0, z1);
    return (idx & 1) ? z1 : z0;
}

// four consecutive noise samples idx4*4 .. idx4*4+3
__device__ __forceinline__ void channel_noise4(const ChannelParams& ch, unsigned long long frame, unsigned int idx4, float z[4])
{
    u32x4 w = philox4x32_10(idx4, (unsigned int)frame, (unsigned int)(frame >> 32), ch.stream,
                            (unsigned int)ch.seed, (unsigned int)(ch.seed >> 32));
    box_muller(w.x, w.y, z[0], z[1]);
    box_muller(w.z, w.w, z[2], z[3]);
}

__device__ __forceinline__ float bpsk_llr(const ChannelParams& ch, float n)
{
    return __fmul_rn(__fmaf_rn(-ch.sigma, n, 1.0f), ch.llr_scale);
}

// one output of Demodulate's if-ladder (e.g. QAM_demodulator.cpp:215-239)
__device__ __forceinline__ double demod_out(double p0, double p1, double T, int out_type)
{
    if (p0 == 0.0) return out_type == 0 ? T : 1.0;
    if (p1 == 0.0) return out_type == 0 ? -T : 0.0;
    return out_type == 0 ? log(div_normal(p1, p0)) : p1;
}

// Exact bit metrics of one PAM component (I or Q) of a Gray-mapped QAM symbol, m = 4, 6 or 8 bits per
// symbol: o[0 .. m/2) in the order Demodulate writes them.  Evaluation order follows
// QAM_demodulator.cpp:181-199 (t = x - L; t *= t; t /= N0; P normalised before the bit sums) and the
// summation trees of :203-561, so only exp/log can differ from the reference, in the last ulp.
template <int SQ>
__device__ __forceinline__ void pam_likelihoods(double x, double N0, double T, double (&P)[16])
{
    double sum = 0;
#pragma unroll
    for (int i = 0; i < SQ; i++) {
        double t = x - (double)(2 * i - (SQ - 1));
        t *= t;
        t = div_normal(t, N0);                                   // t /= N0, :183
        P[i] = t < T ? exp(-t) : 0.0;
        sum += P[i];
    }
#pragma unroll
    for (int i = 0; i < SQ; i++) P[i] = div_normal(P[i], sum);   // :196-199
}

__device__ __forceinline__ void pam_demod(double x, double N0, double T, int m, int out_type, double* o)
{
    double P[16];
    if (m == 4) {
        pam_likelihoods<4>(x, N0, T, P);
        o[0] = demod_out(P[0] + P[1], P[2] + P[3], T, out_type);
        o[1] = demod_out(P[0] + P[3], P[1] + P[2], T, out_type);
    } else if (m == 6) {
        pam_likelihoods<8>(x, N0, T, P);
        double p12 = P[0] + P[1], p34 = P[2] + P[3], p56 = P[4] + P[5], p78 = P[6] + P[7];
        o[0] = demod_out(p12 + p34, p56 + p78, T, out_type);
        o[1] = demod_out(p12 + p78, p34 + p56, T, out_type);
        o[2] = demod_out(P[0] + P[3] + P[4] + P[7], P[1] + P[2] + P[5] + P[6], T, out_type);
    } else {
        pam_likelihoods<16>(x, N0, T, P);
        double p12 = P[0] + P[1], p34 = P[2] + P[3], p56 = P[4] + P[5], p78 = P[6] + P[7];
        double p9A = P[8] + P[9], pBC = P[10] + P[11], pDE = P[12] + P[13], pFG = P[14] + P[15];
        double p1234 = p12 + p34, p5678 = p56 + p78, p9ABC = p9A + pBC, pDEFG = pDE + pFG;
        o[0] = demod_out(p1234 + p5678, p9ABC + pDEFG, T, out_type);
        o[1] = demod_out(p1234 + pDEFG, p5678 + p9ABC, T, out_type);
        o[2] = demod_out(p12 + p78 + p9A + pFG, p34 + p56 + pBC + pDE, T, out_type);
        o[3] = demod_out(P[0] + P[3] + P[4] + P[7] + P[8] + P[11] + P[12] + P[15],
                         P[1] + P[2] + P[5] + P[6] + P[9] + P[10] + P[13] + P[14], T, out_type);
    }
}

// Demodulate for the IN-KERNEL channel, whose LLRs leave as fp32 (out[b] = (float)(-o[b]), LLR output type).  Same likelihoods,
// same clip, same sums -- but the eight (four, sixteen) exponentials and the normalising divisions of pam_likelihoods become
// ONE exponential and ONE division:
//     exp(-(x - L)^2 / N0) = exp(-x^2 / N0) * W^L * exp(-L^2 / N0),     W = exp(2 x / N0),
// the first factor cancels in every p1 / p0 (so does the reference's normalisation by the sum), the last is a per-SNR constant
// (ChannelParams::qam_c), and W^L for the odd L are a multiplication chain from W and 1 / W.  The clip test t < T runs on
// (x - L)^2 * (1 / N0).  log(p1 / p0) is fx_log_ratio.  The double results agree with pam_demod to a few 1e-15 relative
// (rounding, not algorithm), i.e. the fp32 LLR differs in its last bit about once in 10^7 values; ldpcb200_generate_llr and
// ldpcb200_demodulate, whose outputs are doubles, keep pam_demod.  The host enables it (qam_fast) when no power can leave
// the double range: (sqrt(Q) - 1) * (sqrt(Q) - 1 + 6 sigma) * 2 / N0 <= 600, i.e. every operating point of an LDPC code;
// LDPCB200_QAM_EXACT=1 turns it off.
template <int SQ>
__device__ __forceinline__ void pam_weights_factored(double x, const ChannelParams& ch, double (&P)[16])
{
    const double W = fx_exp(x * ch.qam_w);
    const double V = div_normal(1.0, W);
    const double W2 = W * W, V2 = V * V;
    double Wp[SQ / 2], Vp[SQ / 2];                               // W^(2 j + 1), W^-(2 j + 1)
    Wp[0] = W; Vp[0] = V;
#pragma unroll
    for (int j = 1; j < SQ / 2; j++) { Wp[j] = Wp[j - 1] * W2; Vp[j] = Vp[j - 1] * V2; }
#pragma unroll
    for (int i = 0; i < SQ; i++) {
        const int L = 2 * i - (SQ - 1), j = ((L < 0 ? -L : L) - 1) / 2;
        const double d = x - (double)L;
        const double t = d * d * ch.qam_n0inv;
        const double q = (L > 0 ? Wp[j] : Vp[j]) * ch.qam_c[j];
        P[i] = t < ch.T ? q : 0.0;
    }
}

__device__ __forceinline__ float demod_out_factored(double p0, double p1, double T)
{
    const double lg = fx_log_ratio(p1, p0);
    return (float)(p0 == 0.0 ? -T : p1 == 0.0 ? T : -lg);
}

template <int SQ>
__device__ __forceinline__ void pam_demod_factored_t(double x, const ChannelParams& ch, float* o)
{
    double P[16];
    const double T = ch.T;
    pam_weights_factored<SQ>(x, ch, P);
    if constexpr (SQ == 4) {
        o[0] = demod_out_factored(P[0] + P[1], P[2] + P[3], T);
        o[1] = demod_out_factored(P[0] + P[3], P[1] + P[2], T);
    } else if constexpr (SQ == 8) {
        double p12 = P[0] + P[1], p34 = P[2] + P[3], p56 = P[4] + P[5], p78 = P[6] + P[7];
        o[0] = demod_out_factored(p12 + p34, p56 + p78, T);
        o[1] = demod_out_factored(p12 + p78, p34 + p56, T);
        o[2] = demod_out_factored(P[0] + P[3] + P[4] + P[7], P[1] + P[2] + P[5] + P[6], T);
    } else {
        double p12 = P[0] + P[1], p34 = P[2] + P[3], p56 = P[4] + P[5], p78 = P[6] + P[7];
        double p9A = P[8] + P[9], pBC = P[10] + P[11], pDE = P[12] + P[13], pFG = P[14] + P[15];
        double p1234 = p12 + p34, p5678 = p56 + p78, p9ABC = p9A + pBC, pDEFG = pDE + pFG;
        o[0] = demod_out_factored(p1234 + p5678, p9ABC + pDEFG, T);
        o[1] = demod_out_factored(p1234 + pDEFG, p5678 + p9ABC, T);
        o[2] = demod_out_factored(p12 + p78 + p9A + pFG, p34 + p56 + pBC + pDE, T);
        o[3] = demod_out_factored(P[0] + P[3] + P[4] + P[7] + P[8] + P[11] + P[12] + P[15],
                                  P[1] + P[2] + P[5] + P[6] + P[9] + P[10] + P[13] + P[14], T);
    }
}

__device__ __forceinline__ void pam_demod_factored(double x, const ChannelParams& ch, float* o)
{
    if (ch.m == 4) pam_demod_factored_t<4>(x, ch, o);
    else if (ch.m == 6) pam_demod_factored_t<8>(x, ch, o);
    else pam_demod_factored_t<16>(x, ch, o);
}

// The decoder input that transmitted position j feeds (inverse permutation of bp_simulation.cpp:684, as a scatter)
__device__ __forceinline__ int channel_dest(const ChannelParams& ch, int j)
{
    return ch.perm_dir ? __ldg(ch.perm_dir + j) : j;
}

// The BPSK / QAM-4 LLRs of the four consecutive transmitted positions 4*i4 .. 4*i4+3 from ONE Philox block (bit-identical
// to channel_llr() of each bit; the per-bit form computes the block four times over).  dst[b] is the decoder input the
// value belongs to (the position itself without an interleaver); punctured inputs get the puncturing value
// (bp_simulation.cpp:697-710 overwrites y[] AFTER the inverse permutation).
__device__ __forceinline__ void channel_llr4_bpsk(const ChannelParams& ch, unsigned long long frame, int i4, float out[4], int dst[4])
{
    float z[4];
    channel_noise4(ch, frame, (unsigned int)i4, z);
#pragma unroll
    for (int b = 0; b < 4; b++) {
        dst[b] = channel_dest(ch, 4 * i4 + b);
        out[b] = (dst[b] >= ch.punct_start) ? ch.punct_value : bpsk_llr(ch, z[b]);
    }
}

// QAM-16/64/256 LLR of the bit at TRANSMITTED position t (kept out of line: it is heavy in registers and only used by C3-like runs)
static __device__ __noinline__ float channel_llr_qam(const ChannelParams& ch, unsigned long long frame, int t)
{
    // position t lives in symbol t / m; the first m/2 bits ride on I, the rest on Q
    const int half = ch.m >> 1;
    int sym = t / ch.m, r = t - sym * ch.m;
    int comp = r >= half;
    int bit = r - comp * half;
    float nz = channel_noise(ch, frame, (unsigned int)(2 * sym + comp));
    // the component's bits, MSB first, as a natural index -> gray[] -> coordinate 2 pos - (sqrt(Q) - 1)  (QAM_modulator.cpp:127-194);
    // all-zero bits -> index 0 -> gray[0] = 0 -> -(sqrt(Q) - 1)
    int pos = 0;
    if (ch.cw) {
        const int gray[16] = { 0, 1, 3, 2, 7, 6, 4, 5, 15, 14, 12, 13, 8, 9, 11, 10 };
        int z = 0;
        const int t0 = sym * ch.m + comp * half;
        for (int q = 0; q < half; q++) {
            const int src = ch.perm_dir ? __ldg(ch.perm_dir + t0 + q) : t0 + q;          // codeword bit sent at position t0 + q (direct permutation, :573)
            z = (z << 1) | (int)(ch.cw[src] & 1);
        }
        pos = gray[z];
    }
    double x = (double)nz * ch.sigma_d + (double)(2 * pos - ((1 << half) - 1));
    if (ch.qam_fast) {
        float of[4];
        pam_demod_factored(x, ch, of);
        return of[bit];
    }
    double o[4];
    pam_demod(x, 2.0 * ch.sigma_d * ch.sigma_d, ch.T, ch.m, 0, o);
    return (float)(-o[bit]);
}

// All m/2 LLRs carried by one PAM component (cidx = 2 * symbol + component) of a QAM-16/64/256 frame: one noise
// sample, one pass through pam_demod.  out[b] is the LLR of transmitted position (cidx / 2) * m + (cidx & 1) * m/2 + b;
// the caller maps it to its decoder input (channel_dest) and applies the puncturing.  Bit-identical to
// channel_llr_qam() bit by bit.
static __device__ __noinline__ void channel_llr_qam_component(const ChannelParams& ch, unsigned long long frame, int cidx, float out[4])
{
    const int half = ch.m >> 1;
    float nz = channel_noise(ch, frame, (unsigned int)cidx);
    double x = (double)nz * ch.sigma_d - (double)((1 << half) - 1);
    if (ch.qam_fast) {
        out[2] = out[3] = 0.0f;
        pam_demod_factored(x, ch, out);
        return;
    }
    double o[4];
    pam_demod(x, 2.0 * ch.sigma_d * ch.sigma_d, ch.T, ch.m, 0, o);
    for (int b = 0; b < 4; b++) out[b] = b < half ? (float)(-o[b]) : 0.0f;
}

// The four PAM components 4*c4 .. 4*c4+3 of an all-zero frame: their noise samples are the four outputs of ONE Philox block
// (channel_noise(idx) takes output idx & 3 of block idx >> 2), and the four factored demodulations are straight-line code next
// to each other, so their exp -> 1/W -> powers -> sums -> log chains overlap (12 warps per SM at C3 cannot hide one chain's
// latency).  out[4 q + b]: bit b of component 4*c4 + q, bit-identical to channel_llr_qam_component() of that component.
template <int SQ>
__device__ __forceinline__ void qam_components4_t(const ChannelParams& ch, const float z[4], float* out)
{
#pragma unroll
    for (int q = 0; q < 4; q++) {
        out[4 * q + 2] = out[4 * q + 3] = 0.0f;
        pam_demod_factored_t<SQ>((double)z[q] * ch.sigma_d - (double)(SQ - 1), ch, out + 4 * q);
    }
}
static __device__ __noinline__ void channel_llr_qam_component4(const ChannelParams& ch, unsigned long long frame, int c4, float* out)
{
    float z[4];
    channel_noise4(ch, frame, (unsigned int)c4, z);
    if (ch.qam_fast) {
        if (ch.m == 4) qam_components4_t<4>(ch, z, out);
        else if (ch.m == 6) qam_components4_t<8>(ch, z, out);
        else qam_components4_t<16>(ch, z, out);
        return;
    }
    const int half = ch.m >> 1;
    for (int q = 0; q < 4; q++) {
        double o[4];
        pam_demod((double)z[q] * ch.sigma_d - (double)((1 << half) - 1), 2.0 * ch.sigma_d * ch.sigma_d, ch.T, ch.m, 0, o);
        for (int b = 0; b < 4; b++) out[4 * q + b] = b < half ? (float)(-o[b]) : 0.0f;
    }
}

// Channel LLR (log P0/P1, the decoder-side sign) of decoder input i of frame f: the value received at transmitted
// position perm_inv[i] (bp_simulation.cpp:684), for the all-zero codeword or ch.cw.
__device__ __forceinline__ float channel_llr(const ChannelParams& ch, unsigned long long frame, int i)
{
    if (i >= ch.punct_start) return ch.punct_value;
    const int t = ch.perm_inv ? __ldg(ch.perm_inv + i) : i;
    if (ch.m <= 2) {
        // -2 (sigma n + 2 c - 1) / sigma^2 (:600-612): c = 0 -> (1 - sigma n) 2 / sigma^2, c = 1 -> (-1 - sigma n) 2 / sigma^2
        const float n = channel_noise(ch, frame, (unsigned int)t);
        if (ch.cw && (ch.cw[i] & 1)) return __fmul_rn(__fmaf_rn(-ch.sigma, n, -1.0f), ch.llr_scale);
        return bpsk_llr(ch, n);
    }
    return channel_llr_qam(ch, frame, t);
}
#endif

} // namespace ldpcb200
