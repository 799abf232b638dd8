// Cost-ordered strip scheduling (longest-processing-time-first from frame-to-frame coherence).
//
// Strip costs are very uneven (killeroo: the cells where the ground plane meets the body hold up to
// 426 triangles; a strip there takes ~7x the average).  With strips handed out in image order the
// frame ends with a few warps grinding through expensive strips while the rest of the GPU idles --
// about 0.4 ms per launch, which is 3 % of a 12 ms frame but 20 % of the 2 ms share of an 8-way
// sharded one.  K1 records the cycles each strip took; after the frame this file turns them into
// the NEXT frame's visiting order: expensive strips (> 2x the mean) first, everything else behind
// them in image order (a stable partition, so L1 locality of neighbouring strips is kept).
// An expensive strip is also SPLIT: it is visited as `parts` pieces of whole 32-ray rounds (whole
// pixels), taken by different warps -- the costliest killeroo strip runs 1.6 ms as one piece, most of
// an 8-way shard's 1.9 ms frame.  A visit entry = strip | (piece + 1) << 28 (0 on top = whole strip).
// The order only affects scheduling, never results; a frame whose layout differs from the
// previous one simply runs in image order.
#include "trace_kernels.cuh"

namespace rtm
{

namespace
{

// K1 records cycles per VISIT; a strip split into pieces was visited several times
__global__ void visit_costs_to_strips_kernel(const uint32_t *__restrict__ visit_cycles, const uint32_t *__restrict__ order,
                                             const uint32_t *__restrict__ visit_total, uint32_t n,
                                             uint32_t *__restrict__ strip_cycles)
{
    const uint32_t visits = order ? *visit_total : n;
    for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < visits; v += gridDim.x * blockDim.x)
    {
        const uint32_t c = visit_cycles[v];
        if (!order)
            strip_cycles[v] = c;
        else if (c)
            atomicAdd(&strip_cycles[order[v] & kVisitStripMask], c);
    }
}

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
                                                                    uint32_t parts, uint32_t *__restrict__ visit_total,
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
        // expensive first (each as `parts` consecutive pieces), both halves in image order
        const uint32_t n_heavy = block_base[nb];
        if (!heavy)
            order[n_heavy * parts + (i - heavy_rank)] = i;
        else if (parts == 1)
            order[heavy_rank] = i;
        else
            for (uint32_t k = 0; k < parts; k++)
                order[heavy_rank * parts + k] = i | ((k + 1u) << 28);
        if (i == 0)
            *visit_total = n + n_heavy * (parts - 1u);
    }
}

} // namespace

size_t strip_order_scratch_words(uint32_t n)
{
    const size_t nb = (n + kOrderBlock - 1) / kOrderBlock;
    return 2 * nb + 2;
}

// "expensive" = above twice the mean, so fewer than n / 2 strips are split
size_t strip_order_capacity(uint32_t n, uint32_t parts)
{
    return (size_t) n + ((size_t) n / 2 + 1) * (parts - 1);
}

// cycles[n] (written by K1) -> order[*visit_total]; sum and scratch are work buffers.  5 kernels + memsets.
void launch_build_strip_order(const uint32_t *visit_cycles, bool order_was_used, uint32_t *strip_cycles, uint32_t n,
                              uint32_t parts, unsigned long long *sum, uint32_t *scratch, uint32_t *visit_total,
                              uint32_t *order, cudaStream_t stream)
{
    if (n == 0)
        return;
    if (order_was_used)
        cudaMemsetAsync(strip_cycles, 0, sizeof(uint32_t) * n, stream);
    visit_costs_to_strips_kernel<<<296, 512, 0, stream>>>(visit_cycles, order_was_used ? order : nullptr, visit_total, n,
                                                          strip_cycles);
    const uint32_t *cycles = strip_cycles;
    const uint32_t nb = (n + kOrderBlock - 1) / kOrderBlock;
    uint32_t *block_count = scratch, *block_base = scratch + nb;
    cudaMemsetAsync(sum, 0, sizeof(unsigned long long), stream);
    cost_sum_kernel<<<64, 256, 0, stream>>>(cycles, n, sum);
    order_count_kernel<<<nb, kOrderBlock, 0, stream>>>(cycles, n, sum, block_count);
    order_scan_kernel<<<1, 1024, 0, stream>>>(block_count, nb, block_base);
    order_scatter_kernel<<<nb, kOrderBlock, 0, stream>>>(cycles, n, sum, block_base, nb, parts, visit_total, order);
}

} // namespace rtm
