// libcuda_trace_measure.so -- MEASUREMENT AND SELF-CHECK code, kept out of the product library
// (include/cuda_trace_measure.h; loaded by bench.py, tools/ and tests/ only):
//   * device ceilings the trace kernel is judged against, measured on the box (SURVEY.md section 8d asks for them
//     next to the driver's HBM / bf16 figures): FP32 throughput WITHOUT fused multiply-add -- the kernels are
//     built with -fmad=false, so a multiply and an add are two instructions -- and L2 read bandwidth on a buffer
//     that fits the L2 (the scenes of C1-C4 are L2-resident);
//   * an L2 flush between timed frames;
//   * a check of the range-check-free reciprocal / division / square root sequences of rt_device.cuh against the
//     IEEE intrinsics over many operands.
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/cuda_trace_measure.h"
#include "rt_device.cuh"

namespace rtm
{

namespace
{

// 8 independent chains of (mul, add) per thread: 16 FP32 instructions per round, no FMA (file built with
// -fmad=false; the SASS holds FMUL + FADD pairs)
__global__ void __launch_bounds__(1024) fp32_nonfma_kernel(float *out, uint32_t rounds, float m, float a)
{
    float v0 = threadIdx.x * 1e-3f, v1 = v0 + 1.0f, v2 = v0 + 2.0f, v3 = v0 + 3.0f;
    float v4 = v0 + 4.0f, v5 = v0 + 5.0f, v6 = v0 + 6.0f, v7 = v0 + 7.0f;
#pragma unroll 4
    for (uint32_t r = 0; r < rounds; r++)
    {
        v0 = v0 * m; v1 = v1 * m; v2 = v2 * m; v3 = v3 * m; v4 = v4 * m; v5 = v5 * m; v6 = v6 * m; v7 = v7 * m;
        v0 = v0 + a; v1 = v1 + a; v2 = v2 + a; v3 = v3 + a; v4 = v4 + a; v5 = v5 + a; v6 = v6 + a; v7 = v7 + a;
    }
    const float s = ((v0 + v1) + (v2 + v3)) + ((v4 + v5) + (v6 + v7));
    if (s == 12345.678f) // never true for the launch values; keeps the chains alive
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// every CTA streams the whole buffer `passes` times with 16-byte loads, starting at a different offset
__global__ void __launch_bounds__(1024) l2_read_kernel(const uint4 *__restrict__ buf, uint32_t n_vec, uint32_t passes,
                                                       uint32_t *out)
{
    uint32_t acc = 0;
    const uint32_t start = (uint32_t) (((uint64_t) blockIdx.x * n_vec) / gridDim.x);
    for (uint32_t p = 0; p < passes; p++)
        for (uint32_t i = threadIdx.x; i < n_vec; i += blockDim.x)
        {
            uint32_t k = start + i;
            k = k >= n_vec ? k - n_vec : k;
            const uint4 v = __ldcg(buf + k); // cache at L2 only: measures L2, not L1
            acc += v.x ^ v.y ^ v.z ^ v.w;
        }
    if (acc == 0x9E3779B9u)
        out[blockIdx.x] = acc;
}

} // namespace

} // namespace rtm

extern "C" int rtm_measure_peaks(int device, double *fp32_nonfma_tflops, double *l2_read_gbps)
{
    using namespace rtm;
    if (!fp32_nonfma_tflops || !l2_read_gbps)
        return RTM_MEASURE_ERR_ARG;
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess)
        return RTM_MEASURE_ERR_CUDA;
    const int sms = prop.multiProcessorCount;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float *d_out = nullptr;
    uint4 *d_buf = nullptr;
    const size_t buf_bytes = 48u << 20; // well inside the 126 MB L2
    int rc = RTM_MEASURE_OK;
    if (cudaMalloc(&d_out, (size_t) sms * 2 * 1024 * sizeof(float)) != cudaSuccess ||
        cudaMalloc(&d_buf, buf_bytes) != cudaSuccess || cudaMemset(d_buf, 1, buf_bytes) != cudaSuccess)
        rc = RTM_MEASURE_ERR_CUDA;
    if (rc == RTM_MEASURE_OK)
    {
        const uint32_t rounds = 1u << 16;
        const int blocks = sms * 2; // 2 x 1024 threads per SM = full occupancy
        double best = 0.0;
        for (int it = 0; it < 4; it++)
        {
            cudaEventRecord(e0);
            fp32_nonfma_kernel<<<blocks, 1024>>>(d_out, rounds, 1.0000001f, 1e-7f);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0.0f;
            cudaEventElapsedTime(&ms, e0, e1);
            const double flops = (double) blocks * 1024.0 * rounds * 16.0;
            if (it > 0 && ms > 0.0f)
                best = flops / (ms * 1e-3) / 1e12 > best ? flops / (ms * 1e-3) / 1e12 : best;
        }
        *fp32_nonfma_tflops = best;

        const uint32_t n_vec = (uint32_t) (buf_bytes / sizeof(uint4)), passes = 4;
        best = 0.0;
        for (int it = 0; it < 4; it++)
        {
            cudaEventRecord(e0);
            l2_read_kernel<<<sms, 1024>>>(d_buf, n_vec, passes, (uint32_t *) d_out);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0.0f;
            cudaEventElapsedTime(&ms, e0, e1);
            const double bytes = (double) sms * passes * (double) buf_bytes;
            if (it > 0 && ms > 0.0f)
                best = bytes / (ms * 1e-3) / 1e9 > best ? bytes / (ms * 1e-3) / 1e9 : best;
        }
        *l2_read_gbps = best;
        if (cudaGetLastError() != cudaSuccess)
            rc = RTM_MEASURE_ERR_CUDA;
    }
    cudaFree(d_out);
    cudaFree(d_buf);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return rc;
}

// ------------------------------------------------------------------------------------------------ L2 flush
extern "C" int rtm_measure_flush_l2(int device)
{
    static void *scratch[64] = {};
    const size_t bytes = 256u << 20; // twice the 126 MB L2
    if (device < 0 || device >= 64)
        return RTM_MEASURE_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess)
        return RTM_MEASURE_ERR_CUDA;
    if (!scratch[device] && cudaMalloc(&scratch[device], bytes) != cudaSuccess)
        return RTM_MEASURE_ERR_CUDA;
    if (cudaMemset(scratch[device], 0xA5, bytes) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess)
        return RTM_MEASURE_ERR_CUDA;
    return RTM_MEASURE_OK;
}

// ------------------------------------------------------------------------- fast arithmetic against the intrinsics
namespace rtm
{
namespace
{

__device__ __forceinline__ uint32_t mix32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// a float with a random sign and mantissa and an exponent in [lo, hi] (unbiased)
__device__ __forceinline__ float random_float(uint32_t bits, int lo, int hi)
{
    const uint32_t e = (uint32_t) (lo + 127) + (bits >> 24) % (uint32_t) (hi - lo + 1);
    return __uint_as_float((bits & 0x807FFFFFu) | (e << 23));
}

// mismatches[0] rcp_normal vs __frcp_rn, [1] div_normal vs __fdiv_rn, [2] sqrt_normal vs __fsqrt_rn
__global__ void check_fast_arith_kernel(uint64_t n, uint32_t seed, int exp_lo, int exp_hi, unsigned long long *mismatches)
{
    unsigned long long bad[3] = { 0, 0, 0 };
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x)
    {
        const uint32_t h0 = mix32((uint32_t) i ^ seed), h1 = mix32(h0 + (uint32_t) (i >> 32) + 0x9E3779B9u);
        const float b = random_float(h0, exp_lo, exp_hi);
        float a = random_float(h1, exp_lo, exp_hi);
        if ((h1 & 0xFFu) == 0)
            a = __uint_as_float(h1 & 0x80000000u) * 0.0f + ((h1 >> 8) & 1u ? 0.0f : a); // now and then a zero numerator
        const float r = rcp_normal(b);
        bad[0] += __float_as_uint(r) != __float_as_uint(__frcp_rn(b));
        const float q = div_normal(a, b, r), q_ref = __fdiv_rn(a, b);
        // a zero numerator may lose the sign of the zero quotient (documented in rt_device.cuh)
        bad[1] += (a == 0.0f) ? (q != 0.0f) : (__float_as_uint(q) != __float_as_uint(q_ref));
        const float x = fabsf(a == 0.0f ? b : a);
        bad[2] += __float_as_uint(sqrt_normal(x)) != __float_as_uint(__fsqrt_rn(x));
    }
    for (int k = 0; k < 3; k++)
        if (bad[k])
            atomicAdd(mismatches + k, bad[k]);
}

} // namespace
} // namespace rtm

extern "C" int rtm_measure_check_fast_arith(int device, uint64_t n, uint32_t seed, int exp_lo, int exp_hi,
                                            unsigned long long mismatches[3])
{
    if (!mismatches || exp_lo < -126 || exp_hi > 127 || exp_lo > exp_hi)
        return RTM_MEASURE_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess)
        return RTM_MEASURE_ERR_CUDA;
    unsigned long long *d_bad = nullptr;
    if (cudaMalloc(&d_bad, 3 * sizeof(unsigned long long)) != cudaSuccess || cudaMemset(d_bad, 0, 3 * sizeof(unsigned long long)) != cudaSuccess)
        return RTM_MEASURE_ERR_CUDA;
    rtm::check_fast_arith_kernel<<<148 * 8, 256>>>(n, seed, exp_lo, exp_hi, d_bad);
    const cudaError_t e = cudaMemcpy(mismatches, d_bad, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaFree(d_bad);
    return e == cudaSuccess ? RTM_MEASURE_OK : RTM_MEASURE_ERR_CUDA;
}
