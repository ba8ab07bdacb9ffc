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
