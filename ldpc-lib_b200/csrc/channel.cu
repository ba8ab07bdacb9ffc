// Stand-alone channel / utility kernels (see channel.cuh for the device functions they share with the
// decoders' fused first load).
#include "kernels.h"
#include "channel.cuh"

namespace ldpcb200 {

__global__ void unpack_hard_kernel(const uint32_t* __restrict__ words, uint8_t* __restrict__ bytes, int nf, int N, int nwords)
{
    size_t total = (size_t)nf * N;
    for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (size_t)gridDim.x * blockDim.x) {
        size_t f = x / N;
        int i = (int)(x - f * N);
        bytes[x] = (words[f * nwords + (i >> 5)] >> (i & 31)) & 1u;
    }
}

cudaError_t launch_unpack_hard(const uint32_t* words, uint8_t* bytes, int nf, int N, int nwords, cudaStream_t s)
{
    size_t total = (size_t)nf * N;
    int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    if (grid < 1) grid = 1;
    unpack_hard_kernel<<<grid, 256, 0, s>>>(words, bytes, nf, N, nwords);
    return cudaGetLastError();
}

template <typename T>
__global__ void generate_llr_kernel(ChannelParams ch, int N, int nf, T* __restrict__ llr)
{
    size_t total = (size_t)nf * N;
    for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (size_t)gridDim.x * blockDim.x) {
        size_t f = x / N;
        int i = (int)(x - f * N);
        llr[x] = (T)channel_llr(ch, ch.first_frame + f, i);
    }
}

cudaError_t launch_generate_llr(const ChannelParams& ch, int N, int nf, void* llr, int llr_dtype, cudaStream_t s)
{
    size_t total = (size_t)nf * N;
    int grid = (int)((total + 255) / 256 < 148 * 32 ? (total + 255) / 256 : 148 * 32);
    if (grid < 1) grid = 1;
    if (llr_dtype == LDPCB200_F64) generate_llr_kernel<double><<<grid, 256, 0, s>>>(ch, N, nf, (double*)llr);
    else if (llr_dtype == LDPCB200_F32) generate_llr_kernel<float><<<grid, 256, 0, s>>>(ch, N, nf, (float*)llr);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

// Error accounting of decoded frames against a transmitted codeword (bp_simulation.cpp:731-743, 805-810) -- the fused decode
// kernels count against the all-zero codeword the reference sends; with a real codeword (ldpcb200_set_codeword) the decode
// writes packed decisions and this kernel compares them: one warp per frame.
__global__ void count_errors_kernel(const uint32_t* __restrict__ hard_words, const uint32_t* __restrict__ cw_words, const int* __restrict__ iters,
                                    int nf, int N, int R, int nwords, unsigned long long* counters, uint32_t* per_frame)
{
    const int lane = threadIdx.x & 31;
    for (int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; f < nf; f += (gridDim.x * blockDim.x) >> 5) {
        int e = 0, ei = 0;
        for (int w = lane; w < nwords; w += 32) {
            uint32_t x = hard_words[(size_t)f * nwords + w] ^ cw_words[w];
            if (32 * w + 32 > N) x &= (N - 32 * w >= 32) ? 0xffffffffu : ((1u << (N - 32 * w)) - 1u);
            e += __popc(x);
            const int lo = R - 32 * w;                                           // bits >= R are information bits (:738)
            ei += __popc(lo <= 0 ? x : (lo >= 32 ? 0u : (x >> lo) << lo));
        }
        for (int o = 16; o > 0; o >>= 1) { e += __shfl_xor_sync(0xffffffffu, e, o); ei += __shfl_xor_sync(0xffffffffu, ei, o); }
        if (lane == 0) {
            const int ret = iters[f];
            if (per_frame) per_frame[f] = (e ? 0x80000000u : 0u) | (ret >= 0 ? 0x40000000u : 0u) | (uint32_t)(ei < 0xFFFFFF ? ei : 0xFFFFFF);
            atomicAdd(&counters[0], 1ull);
            atomicAdd(&counters[4], (unsigned long long)(ret < 0 ? -ret : ret));
            if (e) {
                atomicAdd(&counters[1], 1ull);
                atomicAdd(&counters[2], (unsigned long long)ei);
                atomicAdd(&counters[5], (unsigned long long)e);
                if (ret >= 0) atomicAdd(&counters[3], 1ull);
            }
        }
    }
}

cudaError_t launch_count_errors(const uint32_t* hard_words, const uint32_t* cw_words, const int* iters, int nf, int N, int R, int nwords,
                                unsigned long long* counters, uint32_t* per_frame, cudaStream_t s)
{
    int grid = (nf + 7) / 8;
    if (grid > 148 * 8) grid = 148 * 8;
    if (grid < 1) grid = 1;
    count_errors_kernel<<<grid, 256, 0, s>>>(hard_words, cw_words, iters, nf, N, R, nwords, counters, per_frame);
    return cudaGetLastError();
}

// the generator's N(0,1) samples 0 .. ns-1 of frames [first_frame, first_frame + nf) (tests: the channel against the reference's modulator / demodulator)
__global__ void generate_noise_kernel(ChannelParams ch, int ns, int nf, float* __restrict__ out)
{
    size_t total = (size_t)nf * ns;
    for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (size_t)gridDim.x * blockDim.x) {
        size_t f = x / ns;
        out[x] = channel_noise(ch, ch.first_frame + f, (unsigned int)(x - f * ns));
    }
}

cudaError_t launch_generate_noise(const ChannelParams& ch, int ns, int nf, float* out, cudaStream_t s)
{
    size_t total = (size_t)nf * ns;
    int grid = (int)((total + 255) / 256 < 148 * 32 ? (total + 255) / 256 : 148 * 32);
    if (grid < 1) grid = 1;
    generate_noise_kernel<<<grid, 256, 0, s>>>(ch, ns, nf, out);
    return cudaGetLastError();
}

// IMS_DEC quantiser pre-pass (decoders.cpp:5472-5479): coef[f] = sqrt(N / sum_i y_i^2) with the sum taken in the
// reference's order, i = 0 .. N-1, in double.  Floating-point addition is not associative, so the chain of one frame
// cannot be split -- but frames are independent: one LANE per frame, 32 chains per warp, hundreds of thousands in
// flight per launch, instead of one thread of a decoding CTA crawling through N dependent additions while the other
// lanes of its CTA wait.  Buffer mode stages 32 frames x 32 values through shared memory so that global reads stay
// coalesced (each lane then walks its own frame's row, padded to 33 words: no bank conflicts).
__global__ void __launch_bounds__(128) ims_energy_kernel(FrameIO io, int N, double* __restrict__ coef)
{
    __shared__ double tile[4][32][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int f0 = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32; f0 < io.nf; f0 += warps * 32) {
        const int f = f0 + lane;
        double en = 0;
        if (io.ch.enabled) {
            if (f < io.nf) {
                const unsigned long long frame = io.ch.first_frame + (unsigned long long)f;
                int i = 0;
                if (io.ch.m <= 2 && !io.ch.perm_inv && !io.ch.cw) {                 // one Philox block -> four consecutive LLRs (the same values, a quarter of the work)
                    for (; i + 4 <= N; i += 4) {
                        float o[4];
                        int d[4];
                        channel_llr4_bpsk(io.ch, frame, i >> 2, o, d);
#pragma unroll
                        for (int b = 0; b < 4; b++) { const double v = (double)o[b]; en += v * v; }
                    }
                }
                for (; i < N; i++) {
                    const double v = (double)channel_llr(io.ch, frame, i);
                    en += v * v;
                }
            }
        } else {
            for (int base = 0; base < N; base += 32) {
                const int i = base + lane;
                for (int r = 0; r < 32; r++) {                     // row r = frame f0 + r, 32 consecutive values: one coalesced read
                    double v = 0;
                    if (f0 + r < io.nf && i < N) {
                        const size_t k = (size_t)(f0 + r) * N + i;
                        v = io.llr_dtype == LDPCB200_F64 ? ((const double*)io.llr)[k] : (double)((const float*)io.llr)[k];
                    }
                    tile[w][r][lane] = v * v;
                }
                __syncwarp();
                const int m = N - base < 32 ? N - base : 32;
                for (int j = 0; j < m; j++) en += tile[w][lane][j];
                __syncwarp();
            }
        }
        if (f < io.nf) coef[f] = sqrt(N / en);
    }
}

cudaError_t launch_ims_energy(const FrameIO& io, int N, double* coef, cudaStream_t s)
{
    int warps = (io.nf + 31) / 32;
    int grid = (warps + 3) / 4;
    if (grid > 148 * 16) grid = 148 * 16;
    if (grid < 1) grid = 1;
    ims_energy_kernel<<<grid, 128, 0, s>>>(io, N, coef);
    return cudaGetLastError();
}

// Demodulate(), QAM_demodulator.cpp:99-566, one thread per (symbol, component)
__global__ void demodulate_kernel(int m, int ns, double sigma, double T, int out_type,
                                  const double* __restrict__ x, double* __restrict__ res)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * ns) return;
    if (m == 2) {                                        // QAM-4, :114-122 (LLR output only)
        res[t] = 2.0 * x[t] / (sigma * sigma);
        return;
    }
    int half = m >> 1, sym = t >> 1, comp = t & 1;
    double o[4];
    pam_demod(x[t], 2.0 * sigma * sigma, T, m, out_type, o);
    for (int i = 0; i < half; i++) res[(size_t)sym * m + comp * half + i] = o[i];
}

cudaError_t launch_demodulate(int m, int ns, double sigma, double T, int out_type, const double* x, double* res, cudaStream_t s)
{
    demodulate_kernel<<<(2 * ns + 127) / 128, 128, 0, s>>>(m, ns, sigma, T, out_type, x, res);
    return cudaGetLastError();
}

// QAM_modulator(), QAM_modulator.cpp:142-194: bits MSB first, I half then Q half, natural index ->
// gray[] position -> 2 pos - (sqrt(Q) - 1)
__global__ void modulate_kernel(int m, int ns, const uint8_t* __restrict__ bits, double* __restrict__ out)
{
    const int gray[16] = { 0, 1, 3, 2, 7, 6, 4, 5, 15, 14, 12, 13, 8, 9, 11, 10 };   // QAM_modulator.cpp:127
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ns) return;
    int half = m >> 1, off = (1 << half) - 1, z1 = 0, z2 = 0;
    for (int i = 0; i < half; i++) {
        z1 = (z1 << 1) | (bits[(size_t)j * m + i] & 1);
        z2 = (z2 << 1) | (bits[(size_t)j * m + half + i] & 1);
    }
    out[2 * j] = 2 * gray[z1] - off;
    out[2 * j + 1] = 2 * gray[z2] - off;
}

cudaError_t launch_modulate(int m, int ns, const uint8_t* bits, double* out, cudaStream_t s)
{
    modulate_kernel<<<(ns + 127) / 128, 128, 0, s>>>(m, ns, bits, out);
    return cudaGetLastError();
}

} // namespace ldpcb200
