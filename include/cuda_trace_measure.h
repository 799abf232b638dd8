/* libcuda_trace_measure.so -- measurement and self-check helpers of the tile tracer.
 *
 * NOT part of the drop-in boundary (that is include/cuda_trace.h) and not linked by the product: nothing here
 * replaces a reference interface.  bench.py, tools/ and tests/ load it for the roofline denominators, the L2 flush
 * between timed frames and the arithmetic self-check.  Source: <package>/csrc/measure.cu.
 */
#ifndef CUDA_TRACE_MEASURE_H
#define CUDA_TRACE_MEASURE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { RTM_MEASURE_OK = 0, RTM_MEASURE_ERR_ARG = 1, RTM_MEASURE_ERR_CUDA = 2 };

/* Measured ceilings for the roofline figures: FP32 throughput without fused multiply-add (the
 * kernels are built -fmad=false) in T instr-flop/s, and L2 read bandwidth on an L2-resident buffer in GB/s.
 * A micro-benchmark (~0.2 s). */
int rtm_measure_peaks(int device, double *fp32_nonfma_tflops, double *l2_read_gbps);

/* Evict everything from the L2 of `device` by overwriting a 256 MiB scratch buffer; synchronises the device.
 * Called between timed frames, outside the timed region. */
int rtm_measure_flush_l2(int device);

/* Compare the range-check-free 1/x, a/b and sqrt(x) sequences the trace kernel uses for operands of ordinary
 * magnitude (csrc/rt_device.cuh: rcp_normal, div_normal, sqrt_normal) with __frcp_rn / __fdiv_rn / __fsqrt_rn on n
 * pseudo-random operands whose exponents lie in [exp_lo, exp_hi]; mismatches[0..2] = differing results. */
int rtm_measure_check_fast_arith(int device, uint64_t n, uint32_t seed, int exp_lo, int exp_hi,
                                 unsigned long long mismatches[3]);

#ifdef __cplusplus
}
#endif

#endif
