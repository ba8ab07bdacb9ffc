// Issue-rate micro-benchmark of the packed 16-bit and three-input instructions the min-sum kernels choose between
// (sm_100a): each kernel runs a long chain-free stream of one instruction in 8 independent accumulators per thread,
// 1024 threads per SM; the result is warp instructions per clock and SM.   nvcc -arch=sm_100a -O3 -o pipe_rates pipe_rates.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
template <int OP>
__global__ void __launch_bounds__(1024) rate(unsigned* out, unsigned a0, unsigned b0)
{
    __align__(8) unsigned r[8];
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = a0 + threadIdx.x * 8 + i;
    unsigned b = b0 + threadIdx.x, c = b0 * 3 + 1;
    double db = 1.0 + 1e-9 * b0, dc = 1e-9 * threadIdx.x;
#pragma unroll 1
    for (int it = 0; it < ITERS / 4; it++) {
#pragma unroll
      for (int u = 0; u < 4; u++) {
        asm volatile("" : "+r"(b), "+r"(c), "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]));
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (OP == 0) asm volatile("min.f16x2 %0, %0, %1;" : "+r"(r[i]) : "r"(b));
            if (OP == 1) asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(r[i]) : "r"(b));
            if (OP == 2) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(b), "r"(c));
            if (OP == 3) r[i] = __vminu2(r[i], b);
            if (OP == 4) r[i] = __vimin3_u16x2(r[i], b, c);
            if (OP == 5) asm volatile("min.f32 %0, %0, %1;" : "+f"(*(float*)&r[i]) : "f"(*(float*)&b));
            if (OP == 6) asm volatile("{ .reg .f32 t; min.f32 t, %0, %1; min.f32 %0, t, %2; }" : "+f"(*(float*)&r[i]) : "f"(*(float*)&b), "f"(*(float*)&c));   // FMNMX3
            if (OP == 7) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(b), "r"(c));
            if (OP == 8) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(*(float*)&r[i]) : "f"(*(float*)&b), "f"(*(float*)&c));
            if (OP == 9) r[i] = __vadd2(r[i], b);
            if (OP == 10) r[i] = __vmins2(r[i], b);
            if (OP == 11) asm volatile("{ .reg .b32 t; abs.f16x2 t, %1; min.f16x2 %0, %0, t; }" : "+r"(r[i]) : "r"(b));
            if (OP == 12) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(*(float*)&r[i]) : "f"(*(float*)&b));
            if (OP == 13) asm volatile("max.f16x2 %0, %0, %1;" : "+r"(r[i]) : "r"(b));
            if (OP == 14) asm volatile("fma.rn.relu.f16x2 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(b), "r"(c));
            if (OP == 15) asm volatile("add.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(b));
            if (OP == 16) { double* dp = (double*)&r[i & ~1]; if ((i & 1) == 0) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(*dp) : "d"(db), "d"(dc)); }
            if (OP == 17) { double* dp = (double*)&r[i & ~1]; if ((i & 1) == 0) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(*dp) : "d"(db)); }
            if (OP == 18) { double* dp = (double*)&r[i & ~1]; if ((i & 1) == 0) asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(*dp)); }
        }
      }
    }
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= r[i];
    if (s == 0x12345678u) out[threadIdx.x] = s;
}

template <int OP>
static void run(const char* name, int sms, double ghz)
{
    unsigned* d;
    cudaMalloc(&d, 4096);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    rate<OP><<<sms, 1024>>>(d, 1, 2);
    cudaEventRecord(e0);
    rate<OP><<<sms, 1024>>>(d, 1, 2);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    // SASS instructions per source statement (cuobjdump of this file): ptxas fuses two dependent two-input minima / adds into
    // one three-input instruction (VHMNMX, VIMNMX3, FMNMX3, IADD3)
    const double mult = (OP >= 16) ? 0.5 : (OP == 0 || OP == 13 || OP == 11 || OP == 3 || OP == 10 || OP == 5 || OP == 15) ? 0.5 : 1.0;
    const double winst = 32.0 * ITERS * 8 * mult;                          // per SM: 32 warps
    printf("{\"op\": \"%s\", \"ms\": %.4f, \"warp_inst_per_clk_per_sm\": %.3f}\n", name, ms, winst / (ms * 1e-3 * ghz * 1e9));
    cudaFree(d);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz / 1e6;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_ghz\": %.3f}\n", p.name, sms, ghz);
    run<0>("min.f16x2 (pairs fused to VHMNMX)", sms, ghz);
    run<13>("max.f16x2 (pairs fused to VHMNMX)", sms, ghz);
    run<11>("min.f16x2 with abs operand (VHMNMX)", sms, ghz);
    run<1>("add.rn.f16x2 (HADD2)", sms, ghz);
    run<2>("fma.rn.f16x2 (HFMA2)", sms, ghz);
    run<14>("fma.rn.relu.f16x2", sms, ghz);
    run<3>("min.u16x2 (pairs fused to VIMNMX3.U16x2)", sms, ghz);
    run<10>("min.s16x2 (pairs fused to VIMNMX3.S16x2)", sms, ghz);
    run<4>("min3 u16x2 (VIMNMX3.U16x2)", sms, ghz);
    run<9>("add.s16x2 (VIADD.16x2)", sms, ghz);
    run<5>("min.f32 (pairs fused to FMNMX3)", sms, ghz);
    run<6>("min3 f32 (FMNMX3)", sms, ghz);
    run<7>("lop3", sms, ghz);
    run<8>("fma.rn.f32 (FFMA)", sms, ghz);
    run<12>("add.rn.f32 (FADD)", sms, ghz);
    run<15>("add.u32 (pairs fused to IADD3)", sms, ghz);
    run<16>("fma.rn.f64 (DFMA)", sms, ghz);
    run<17>("add.rn.f64 (DADD)", sms, ghz);
    run<18>("rcp.approx.ftz.f64 (MUFU.RCP64H)", sms, ghz);
    return 0;
}
