// Helpers shared by the table-driven decoders: persistent frame loop, input load, output emission.
#pragma once
#include "kernels.h"
#include "channel.cuh"

namespace ldpcb200 {

enum { CNT_FRAMES = 0, CNT_FRAME_ERRORS, CNT_INFO_BIT_ERRORS, CNT_UNDETECTED, CNT_ITER_SUM, CNT_BIT_ERRORS };

__device__ __forceinline__ int wrapz(int x, int Z) { return x >= Z ? x - Z : x; }     // x in [0, 2Z)

// Channel LLR i of frame f as the decoder's working type.
__device__ __forceinline__ double load_llr(const FrameIO& io, int N, int f, int i)
{
    if (io.ch.enabled) return (double)channel_llr(io.ch, io.ch.first_frame + (unsigned long long)f, i);
    size_t k = (size_t)f * N + i;
    return io.llr_dtype == LDPCB200_F64 ? ((const double*)io.llr)[k] : (double)((const float*)io.llr)[k];
}

template <typename T>
__device__ __forceinline__ void store_post(const FrameIO& io, int N, int f, int i, T v)
{
    if (!io.post) return;
    size_t k = (size_t)f * N + i;
    switch (io.post_dtype) {
    case LDPCB200_F64: ((double*)io.post)[k] = (double)v; break;
    case LDPCB200_F32: ((float*)io.post)[k] = (float)v; break;
    case LDPCB200_I16: ((int16_t*)io.post)[k] = (int16_t)v; break;
    default:           ((uint16_t*)io.post)[k] = (uint16_t)v; break;
    }
}

// Next frame of the persistent grid (all threads of the CTA get the same value).
__device__ __forceinline__ int next_frame(const FrameIO& io)
{
    __shared__ int s_next;
    __syncthreads();
    if (threadIdx.x == 0) s_next = (int)atomicAdd(io.next_frame, 1u);
    __syncthreads();
    return s_next;
}

// Pack the hard decisions of frame f (hardbit(i) in {0,1}), count errors against the all-zero
// codeword the reference always transmits (bp_simulation.cpp:568, 731-743) and update the counters.
// Must be called by all threads of the CTA; blockDim.x is a multiple of 32.
template <class HardFn>
__device__ __forceinline__ void emit_frame(const QcDev& g, const FrameIO& io, int f, int ret, HardFn hardbit)
{
    __shared__ int s_err[2];
    if (threadIdx.x < 2) s_err[threadIdx.x] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    int nerr = 0, nerr_info = 0;
    const int nround = (g.N + 31) & ~31;
    for (int i = threadIdx.x; i < nround; i += blockDim.x) {
        int bit = i < g.N ? hardbit(i) : 0;
        unsigned w = __ballot_sync(0xffffffffu, bit);
        if (lane == 0) {
            if (io.hard_words) io.hard_words[(size_t)f * g.nwords + (i >> 5)] = w;
            nerr += __popc(w);
            // bits >= R are information bits (bp_simulation.cpp:738)
            int lo = g.R - i;                       // number of leading parity bits inside this word
            unsigned wi = lo <= 0 ? w : (lo >= 32 ? 0u : (w >> lo) << lo);
            nerr_info += __popc(wi);
        }
    }
    if (lane == 0 && nerr) { atomicAdd(&s_err[0], nerr); atomicAdd(&s_err[1], nerr_info); }
    __syncthreads();
    if (threadIdx.x == 0) {
        int e = s_err[0], ei = s_err[1];
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
    __syncthreads();
}

// carve a typed array out of the per-CTA workspace slice
template <typename T>
__device__ __forceinline__ T* carve(char*& p, size_t n)
{
    T* r = (T*)p;
    p += (n * sizeof(T) + 15) & ~(size_t)15;
    return r;
}
inline size_t carve_bytes(size_t n, size_t elem) { return (n * elem + 15) & ~(size_t)15; }

} // namespace ldpcb200
