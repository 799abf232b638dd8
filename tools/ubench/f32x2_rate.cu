// Micro-benchmark (measurement only, not product code): issue rate of packed fp32 (FMUL2 / FFMA2) against the
// scalar FMUL / FADD it replaces in the triangle test, at the trace kernel's occupancy (1 x 1024 threads per SM)
// and at full occupancy.  Prints warp-instructions per cycle per SM sub-partition and scalar-equivalent Tflop/s.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o f32x2_rate f32x2_rate.cu && ./f32x2_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }

// MODE 0: scalar mul + add, 8 chains (16 instr / round, 16 flop)
// MODE 1: packed mul2 + fma2(x, ONE, a), 8 chains (16 instr / round, 32 flop)
// MODE 2: packed mul2 with a scalar-broadcast operand + fma2 (as in the kernel)
// MODE 3: scalar mix 8 FMUL + 8 FADD + 8 FSETP-ish (adds ALU pressure) -- not used
template <int MODE>
__global__ void __launch_bounds__(1024) rate_kernel(float *out, uint32_t rounds, float m, float a, f32x2 one)
{
    if (MODE == 0)
    {
        float v[8];
        for (int i = 0; i < 8; i++) v[i] = threadIdx.x * 1e-3f + i;
#pragma unroll 4
        for (uint32_t r = 0; r < rounds; r++)
        {
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = v[i] * m;
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = v[i] + a;
        }
        float s = 0; for (int i = 0; i < 8; i++) s += v[i];
        if (s == 12345.678f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    }
    else
    {
        f32x2 v[8];
        for (int i = 0; i < 8; i++) v[i] = pk2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
        const f32x2 m2 = pk2(m, m), a2 = pk2(a, a + 1e-9f);
#pragma unroll 4
        for (uint32_t r = 0; r < rounds; r++)
        {
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = mul2(v[i], MODE == 2 ? m2 : a2);
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = fma2(v[i], one, a2);
        }
        f32x2 s = 0; for (int i = 0; i < 8; i++) s ^= v[i];
        if (s == 0x1234567812345678ull) out[blockIdx.x * blockDim.x + threadIdx.x] = 1.0f;
    }
}

template <int MODE>
void run(const char *name, int sms, int ctas_per_sm, double clock_ghz)
{
    float *d_out; cudaMalloc(&d_out, (size_t) sms * 2 * 1024 * 4);
    const float one[2] = {1.0f, 1.0f}; f32x2 one2; memcpy(&one2, one, 8);
    const uint32_t rounds = 1u << 15;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int it = 0; it < 4; it++)
    {
        cudaEventRecord(e0);
        rate_kernel<MODE><<<sms * ctas_per_sm, 1024>>>(d_out, rounds, 1.0000001f, 1e-7f, one2);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (it) best = ms < best ? ms : best;
    }
    const double warp_instr = (double) sms * ctas_per_sm * 32 * rounds * 16.0;
    const double flops = warp_instr * 32 * (MODE == 0 ? 1 : 2);
    const double cycles = best * 1e-3 * clock_ghz * 1e9;
    printf("%-44s ctas/sm %d  %.3f ms  %.2f warp-instr/clk/SMSP  %.1f T scalar-equivalent flop/s\n", name, ctas_per_sm, best,
           warp_instr / cycles / (sms * 4), flops / (best * 1e-3) / 1e12);
    cudaFree(d_out);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz * 1e-6;
    printf("%s, %d SMs, %.3f GHz (max)\n", p.name, p.multiProcessorCount, ghz);
    for (int c = 1; c <= 2; c++)
    {
        run<0>("scalar FMUL + FADD", p.multiProcessorCount, c, ghz);
        run<1>("packed FMUL2 + FFMA2(x, ONE, a)", p.multiProcessorCount, c, ghz);
        run<2>("packed FMUL2(broadcast) + FFMA2(x, ONE, a)", p.multiProcessorCount, c, ghz);
    }
    return 0;
}
