// Cost-ordered strip scheduling (longest-processing-time-first from frame-to-frame coherence).
//
// Strip costs are very uneven (killeroo: the cells where the ground plane meets the body hold up to
// 426 triangles; a strip there takes ~7x the average).  With strips handed out in image order the
// frame ends with a few warps grinding through expensive strips while the rest of the GPU idles --
// about 0.4 ms per launch, which is 3 % of a 12 ms frame but 20 % of the 2 ms share of an 8-way
// sharded one.  K1 records the cycles each strip took; after the frame this file turns them into
// the NEXT frame's visiting order: expensive strips (> 2x the mean) first, everything else behind
// them in image order (a stable partition, so L1 locality of neighbouring strips is kept).
// The order only affects scheduling, never results; a frame whose layout differs from the
// previous one simply runs in image order.
#include "trace_kernels.cuh"

namespace rtm
{

namespace
{

__global__ void cost_sum_kernel(const uint32_t *__restrict__ cycles, uint32_t n, unsigned long long *__restrict__ sum)
{
    unsigned long long s = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        s += cycles[i];
    for (int o = 16; o > 0; o >>= 1)
        s += __shfl_down_sync(0xFFFFFFFFu, s, o);
    if ((threadIdx.x & 31) == 0 && s)
        atomicAdd(sum, s);
}

// Stable two-way partition in three small kernels: per-block counts of expensive strips, a
// single-CTA scan of those counts, then a ballot-ranked scatter.  ~15 us for the 1 M strips of a
// 4K frame; runs after the frame, off the timed kernel.
constexpr uint32_t kOrderBlock = 1024;

__device__ __forceinline__ unsigned long long heavy_threshold(const unsigned long long *sum, uint32_t n)
{
    return 2ull * (*sum) / (n ? n : 1u);
}

__global__ void __launch_bounds__(kOrderBlock) order_count_kernel(const uint32_t *__restrict__ cycles, uint32_t n,
                                                                  const unsigned long long *__restrict__ sum,
                                                                  uint32_t *__restrict__ block_count)
{
    __shared__ uint32_t s_warp[kOrderBlock / 32];
    const uint32_t i = blockIdx.x * kOrderBlock + threadIdx.x;
    const bool heavy = i < n && cycles[i] > heavy_threshold(sum, n);
    const unsigned ballot = __ballot_sync(0xFFFFFFFFu, heavy);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = __popc(ballot);
    __syncthreads();
    if (threadIdx.x < 32)
    {
        uint32_t v = s_warp[threadIdx.x];
        for (int o = 16; o > 0; o >>= 1)
            v += __shfl_down_sync(0xFFFFFFFFu, v, o);
        if (threadIdx.x == 0) block_count[blockIdx.x] = v;
    }
}

// exclusive scan of block_count[nb] -> block_base[nb], total in block_base[nb]
__global__ void __launch_bounds__(1024) order_scan_kernel(const uint32_t *__restrict__ block_count, uint32_t nb,
                                                          uint32_t *__restrict__ block_base)
{
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t tile = 0; tile < nb; tile += 1024)
    {
        const uint32_t i = tile + threadIdx.x;
        const uint32_t v = i < nb ? block_count[i] : 0u;
        uint32_t inc = v;
        for (int o = 1; o < 32; o <<= 1)
        {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
            if (lane >= (uint32_t) o) inc += t;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        if (warp == 0)
        {
            uint32_t w = s_warp[lane];
            for (int o = 1; o < 32; o <<= 1)
            {
                const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, w, o);
                if (lane >= (uint32_t) o) w += t;
            }
            s_warp[lane] = w;
        }
        __syncthreads();
        const uint32_t carry = s_carry;
        if (i < nb)
            block_base[i] = carry + inc - v + (warp ? s_warp[warp - 1] : 0u);
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + s_warp[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) block_base[nb] = s_carry;
}

__global__ void __launch_bounds__(kOrderBlock) order_scatter_kernel(const uint32_t *__restrict__ cycles, uint32_t n,
                                                                    const unsigned long long *__restrict__ sum,
                                                                    const uint32_t *__restrict__ block_base, uint32_t nb,
                                                                    uint32_t *__restrict__ order)
{
    __shared__ uint32_t s_warp[kOrderBlock / 32];
    const uint32_t i = blockIdx.x * kOrderBlock + threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool heavy = i < n && cycles[i] > heavy_threshold(sum, n);
    const unsigned ballot = __ballot_sync(0xFFFFFFFFu, heavy);
    if (lane == 0) s_warp[warp] = __popc(ballot);
    __syncthreads();
    uint32_t before = 0;
    for (uint32_t w = 0; w < warp; w++) before += s_warp[w];
    if (i < n)
    {
        const uint32_t heavy_rank = block_base[blockIdx.x] + before + __popc(ballot & ((1u << lane) - 1u));
        order[heavy ? heavy_rank : block_base[nb] + (i - heavy_rank)] = i; // expensive first, both halves in image order
    }
}

} // namespace

size_t strip_order_scratch_words(uint32_t n)
{
    const size_t nb = (n + kOrderBlock - 1) / kOrderBlock;
    return 2 * nb + 2;
}

// cycles[n] (written by K1) -> order[n]; sum and scratch are work buffers.  4 kernels + one memset.
void launch_build_strip_order(const uint32_t *cycles, uint32_t n, unsigned long long *sum, uint32_t *scratch,
                              uint32_t *order, cudaStream_t stream)
{
    if (n == 0)
        return;
    const uint32_t nb = (n + kOrderBlock - 1) / kOrderBlock;
    uint32_t *block_count = scratch, *block_base = scratch + nb;
    cudaMemsetAsync(sum, 0, sizeof(unsigned long long), stream);
    cost_sum_kernel<<<64, 256, 0, stream>>>(cycles, n, sum);
    order_count_kernel<<<nb, kOrderBlock, 0, stream>>>(cycles, n, sum, block_count);
    order_scan_kernel<<<1, 1024, 0, stream>>>(block_count, nb, block_base);
    order_scatter_kernel<<<nb, kOrderBlock, 0, stream>>>(cycles, n, sum, block_base, nb, order);
}

} // namespace rtm
