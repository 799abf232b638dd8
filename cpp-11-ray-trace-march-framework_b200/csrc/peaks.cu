// Device ceilings the trace kernel is judged against, measured on the box (SURVEY.md section 8d asks for them
// next to the driver's HBM / bf16 figures): FP32 throughput WITHOUT fused multiply-add -- the kernels are built
// with -fmad=false, so a multiply and an add are two instructions -- and L2 read bandwidth on a buffer that fits
// the L2 (the scenes of C1-C4 are L2-resident).  Not on any product path.
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/cuda_trace.h"

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

extern "C" int cuda_trace_measure_peaks(int device, double *fp32_nonfma_tflops, double *l2_read_gbps)
{
    using namespace rtm;
    if (!fp32_nonfma_tflops || !l2_read_gbps)
        return CUDA_TRACE_ERR_ARG;
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess)
        return CUDA_TRACE_ERR_CUDA;
    const int sms = prop.multiProcessorCount;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float *d_out = nullptr;
    uint4 *d_buf = nullptr;
    const size_t buf_bytes = 48u << 20; // well inside the 126 MB L2
    int rc = CUDA_TRACE_OK;
    if (cudaMalloc(&d_out, (size_t) sms * 2 * 1024 * sizeof(float)) != cudaSuccess ||
        cudaMalloc(&d_buf, buf_bytes) != cudaSuccess || cudaMemset(d_buf, 1, buf_bytes) != cudaSuccess)
        rc = CUDA_TRACE_ERR_CUDA;
    if (rc == CUDA_TRACE_OK)
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
            rc = CUDA_TRACE_ERR_CUDA;
    }
    cudaFree(d_out);
    cudaFree(d_buf);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return rc;
}
