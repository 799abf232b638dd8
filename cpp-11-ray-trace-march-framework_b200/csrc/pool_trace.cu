// K7 trace_pool: the traversal of K1 for INCOHERENT rays -- grids too large for a shared-memory occupancy map
// (the 50 M-triangle soup at 512^3), where the 32 rays of a warp part ways after a few cells.
//
// K1 keeps a warp's 32 rays in lock step between its two phases: lanes that have reached an occupied cell wait
// for the last walker (measured on the soup: 10.9 of 32 lanes walking, 18 of 32 threads per instruction overall).
// Here a warp owns a POOL of 64 rays in shared memory and picks, round by round, 32 rays that all need the same
// kind of work -- the active lanes are compacted by ballot / population count across the divergent DDA walks:
//
//   refill  32 new rays (the next round of the warp's current strip): ray generation, grid entry, DDA set-up
//   walk    32 rays standing in empty space step through the distance map (warp_trace.cuh, kOccGlobalDist) until
//           they reach an occupied cell (-> test), leave the grid (-> miss) or have used their look-ups
//   test    32 rays standing on occupied cells test their cells' pair records (test_pair_list of K1, unchanged)
//           -> hit, or back to walking
//
// Per-ray arithmetic is exactly K1's (same functions / the same included set-up), so every ray visits the same
// cells and tests the same triangles in the same order: results are bit-identical.  Finished rays leave only their
// hit record (triangle, t, u, v per sample); K1 instantiated with kVariantFromHits then shades, sums the samples
// in order, resolves and publishes the row bands exactly as for its own hits.
//
// Compiled with -fmad=false (see rt_device.cuh).
#include "trace_kernels.cuh"
#include "warp_trace.cuh"

namespace rtm
{

namespace
{

constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr uint32_t kSlotFree = 0, kSlotWalk = 1, kSlotTest = 2;
// look-ups of the distance map a ray gets per walk round (bounds how long finished lanes wait for the longest walk)
constexpr uint32_t kWalkLookups = 8;
constexpr uint32_t kMetaStepsMask = 0xFFu, kMetaNegX = 0x100u, kMetaNegY = 0x200u, kMetaNegZ = 0x400u;

// One warp's pool, structure of arrays over the slots
struct PoolWarp
{
    float n[3][kPoolSlots];   // next crossing per axis (grid.cpp:199-214)
    float dl[3][kPoolSlots];  // increment per cell
    float d[3][kPoolSlots];   // ray direction
    int pc[kPoolSlots];       // padded cell index
    uint32_t meta[kPoolSlots]; // bits 0-7: steps up to and including the next cell to look at (0: standing on a cell
                               // not looked at yet); bits 8-10: the x / y / z stride is negative
    uint32_t ray[kPoolSlots]; // index of the sample's hit record: (py * width + px) * spp + s
    uint32_t beg[kPoolSlots], len[kPoolSlots]; // pair records of the occupied cell a tester stands on
    uint32_t status[kPoolSlots];
    uint32_t list[32];        // the slots picked for this round, compacted
};
static_assert(sizeof(PoolWarp) == kPoolWarpBytes, "trace_kernels.cuh: kPoolWarpBytes");

// owners (lane l holds the status of slots l and l + 32) compact the slots of one kind into list[0 .. 32)
__device__ __forceinline__ void pick_slots(PoolWarp& pw, uint32_t lane, bool mine0, bool mine1, unsigned m0, unsigned m1)
{
    const unsigned lt = (1u << lane) - 1u;
    const uint32_t r0 = __popc(m0 & lt), r1 = __popc(m0) + __popc(m1 & lt);
    if (mine0 && r0 < 32u) pw.list[r0] = lane;
    if (mine1 && r1 < 32u) pw.list[r1] = lane + 32u;
    __syncwarp();
}

__device__ __forceinline__ void store_miss(const TraceParams& p, uint32_t k)
{
    p.hit_tri[k] = 0xFFFFFFFFu;
    p.hit_t[k] = 0.0f;
    p.hit_u[k] = 0.0f;
    p.hit_v[k] = 0.0f;
}

template <bool RCP_GUARD>
__global__ void __launch_bounds__(kTraceMaxThreads) trace_pool_kernel(const __grid_constant__ TraceParams p)
{
    // shared memory: [sample table, spp x float2, padded to 16 bytes] [one PoolWarp per warp]
    extern __shared__ float2 s_mem[];
    float2 *s_smp = s_mem;
    PoolWarp *pools = reinterpret_cast<PoolWarp *>(s_mem + ((p.spp + 1u) & ~1u));
    for (uint32_t i = threadIdx.x; i < p.spp; i += blockDim.x)
        s_smp[i] = p.smp[i];
    __syncthreads();

    const uint32_t lane = threadIdx.x & 31u;
    PoolWarp& pw = pools[threadIdx.x >> 5];
    pw.status[lane] = kSlotFree;
    pw.status[lane + 32u] = kSlotFree;
    __syncwarp();

    PackedUnits pku;
    pku.one = p.pk_one;
    pku.minus_one = p.pk_minus_one;
    const GridDev& g = p.grid;
    const bool fast_math = p.cam.fast_math != 0;
    const int stride_y = ((int) g.dim[0] + 2) * ((int) g.dim[2] + 2), stride_z = (int) g.dim[0] + 2;
    const uint32_t *__restrict__ pstart = g.ppair_start;
    const uint8_t *__restrict__ dist = g.pcell_dist;
    const float4 *__restrict__ recs = g.pair_recs;
    const float3 o = make_float3(p.cam.origin[0], p.cam.origin[1], p.cam.origin[2]); // primary rays share their origin

    // the warp's current strip (warp-uniform): pixel slots [pbase, slots) are still to be handed out
    const uint32_t slots = p.strip_w * p.strip_h, ppr = 32u / p.spp;
    const uint32_t pl = lane / p.spp, s = lane - pl * p.spp;
    uint32_t pbase = slots, bx0 = 0, by0 = 0, bw = 0, bh = 0;
    bool more = true;

    for (;;)
    {
        const uint32_t st0 = pw.status[lane], st1 = pw.status[lane + 32u];
        const unsigned w0 = __ballot_sync(kFull, st0 == kSlotWalk), w1 = __ballot_sync(kFull, st1 == kSlotWalk);
        const unsigned t0 = __ballot_sync(kFull, st0 == kSlotTest), t1 = __ballot_sync(kFull, st1 == kSlotTest);
        const uint32_t n_walk = __popc(w0) + __popc(w1), n_test = __popc(t0) + __popc(t1);
        const uint32_t n_free = (uint32_t) kPoolSlots - n_walk - n_test;
        // what 32 lanes can do together: a full round of tests, else a full round of walks, else take in new rays,
        // else (the pool is draining, or split three ways) whichever kind there is more of
        enum { kRefill, kWalk, kTest } action;
        if (n_test >= 32u) action = kTest;
        else if (n_walk >= 32u) action = kWalk;
        else if (more && n_free >= 32u) action = kRefill;
        else if (n_walk + n_test == 0u) break;
        else action = n_test >= n_walk ? kTest : kWalk;

        if (action == kRefill)
        {
            if (pbase >= slots)
            {
                // next strip of this shard: the scheduler and the strip -> pixel mapping of K1 (trace_kernels.cu)
                uint32_t visit = 0;
                bool cancel_seen = false;
                if (lane == 0)
                {
                    cancel_seen = *(volatile const uint32_t *) p.cancel == p.frame_seq;
                    visit = atomicAdd(p.strip_counter + kPoolCounterWord, 1u);
                    if (cancel_seen)
                        *(volatile uint32_t *) p.cancel_seen = p.frame_seq;
                }
                visit = __shfl_sync(kFull, visit, 0);
                if (__any_sync(kFull, cancel_seen) || visit >= p.shard_strips)
                {
                    more = false; // a cancelled frame drains what is in flight and stops
                    continue;
                }
                const uint32_t my_chunk = visit / p.shard_chunk, in_chunk = visit - my_chunk * p.shard_chunk;
                const uint64_t chunk_id = (uint64_t) my_chunk * p.shard_world +
                                          (p.shard_rank + p.shard_world - my_chunk % p.shard_world) % p.shard_world;
                const uint64_t strip64 = chunk_id * p.shard_chunk + in_chunk;
                if (strip64 >= p.total_strips)
                    continue;
                const uint32_t strip = (uint32_t) strip64;
                uint32_t lo = 0, hi = p.n_tiles;
                while (hi - lo > 1)
                {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (__ldg(&p.tile_strip_prefix[mid]) <= strip) lo = mid; else hi = mid;
                }
                const uint4 rect = __ldg(&p.tile_rects[lo]);
                const uint32_t local = strip - __ldg(&p.tile_strip_prefix[lo]);
                const uint32_t strips_x = (rect.z - rect.x + p.strip_w - 1) / p.strip_w;
                bx0 = rect.x + (local % strips_x) * p.strip_w;
                by0 = rect.y + (local / strips_x) * p.strip_h;
                bw = min(p.strip_w, rect.z - bx0);
                bh = min(p.strip_h, rect.w - by0);
                pbase = 0;
            }
            // one round of the strip: 32 / spp pixels in 2x2-quad order, lane = (pixel in round) * spp + sample
            const uint32_t slot = pbase + pl;
            pbase += ppr;
            const uint32_t ox = (slot & 1u) | ((slot >> 1) & 2u) | ((slot >> 2) & 4u);
            const uint32_t oy = ((slot >> 1) & 1u) | ((slot >> 2) & 2u);
            const bool valid = pl < ppr && slot < slots && ox < bw && oy < bh;
            if (!__any_sync(kFull, valid))
                continue;
            const uint32_t px = bx0 + ox, py = by0 + oy;
            const float2 off = s_smp[valid ? s : 0];
            float3 d, o_ray;
            generate_ray<false>(p.cam, px, py, off.x, off.y, o_ray, d);
            const uint32_t k = (py * p.width + px) * p.spp + s; // (< 2^32: checked by the launcher)
            bool active;
            float n0, n1, n2, dl0, dl1, dl2;
            int c0, c1, c2, pc;
#include "dda_setup.inc"
            if (valid && !active)
                store_miss(p, k); // the ray misses the grid's box (grid.cpp:182-184)
            // the i-th ray of this round takes the i-th free slot
            pick_slots(pw, lane, st0 == kSlotFree, st1 == kSlotFree, ~(w0 | t0), ~(w1 | t1));
            const unsigned need = __ballot_sync(kFull, active);
            if (active)
            {
                const uint32_t slot_id = pw.list[__popc(need & ((1u << lane) - 1u))];
                pw.n[0][slot_id] = n0; pw.n[1][slot_id] = n1; pw.n[2][slot_id] = n2;
                pw.dl[0][slot_id] = dl0; pw.dl[1][slot_id] = dl1; pw.dl[2][slot_id] = dl2;
                pw.d[0][slot_id] = d.x; pw.d[1][slot_id] = d.y; pw.d[2][slot_id] = d.z;
                pw.pc[slot_id] = pc;
                pw.meta[slot_id] = (c0 < 0 ? kMetaNegX : 0u) | (c1 < 0 ? kMetaNegY : 0u) | (c2 < 0 ? kMetaNegZ : 0u);
                pw.ray[slot_id] = k;
                pw.status[slot_id] = kSlotWalk;
            }
            __syncwarp();
        }
        else if (action == kWalk)
        {
            pick_slots(pw, lane, st0 == kSlotWalk, st1 == kSlotWalk, w0, w1);
            if (lane < min(n_walk, 32u))
            {
                const uint32_t slot_id = pw.list[lane];
                float n0 = pw.n[0][slot_id], n1 = pw.n[1][slot_id], n2 = pw.n[2][slot_id];
                const float dl0 = pw.dl[0][slot_id], dl1 = pw.dl[1][slot_id], dl2 = pw.dl[2][slot_id];
                int pc = pw.pc[slot_id];
                const uint32_t meta = pw.meta[slot_id];
                const int c0 = (meta & kMetaNegX) ? -1 : 1;
                const int c1 = (meta & kMetaNegY) ? -stride_y : stride_y;
                const int c2 = (meta & kMetaNegZ) ? -stride_z : stride_z;
                // phase A of K1 on the distance map (warp_trace.cuh): v = distance of the cell the ray stands on ->
                // the next v - 1 steps need no look-up; every step still does its own next_t += delta
                uint32_t steps = meta & kMetaStepsMask, looks = 0;
                bool stop = false;
                if (steps == 0)
                {
                    steps = __ldg(&dist[pc]);
                    stop = steps == 0;
                }
                while (!stop)
                {
                    for (; steps > 1; steps--)
                        dda_step(n0, n1, n2, dl0, dl1, dl2, pc, c0, c1, c2);
                    dda_step(n0, n1, n2, dl0, dl1, dl2, pc, c0, c1, c2);
                    steps = __ldg(&dist[pc]);
                    stop = steps == 0;
                    if (++looks >= kWalkLookups)
                        break;
                }
                uint32_t status = kSlotWalk;
                if (stop)
                {
                    // an occupied cell, or the padding: the ray has left the grid (grid.cpp:275-276)
                    const uint32_t beg = __ldg(&pstart[pc]), len = __ldg(&pstart[pc + 1]) - beg;
                    if (len == 0)
                    {
                        store_miss(p, pw.ray[slot_id]);
                        status = kSlotFree;
                    }
                    else
                    {
                        pw.beg[slot_id] = beg;
                        pw.len[slot_id] = len;
                        status = kSlotTest;
                    }
                }
                pw.n[0][slot_id] = n0; pw.n[1][slot_id] = n1; pw.n[2][slot_id] = n2;
                pw.pc[slot_id] = pc;
                pw.meta[slot_id] = (meta & ~kMetaStepsMask) | steps;
                pw.status[slot_id] = status;
            }
            __syncwarp();
        }
        else
        {
            pick_slots(pw, lane, st0 == kSlotTest, st1 == kSlotTest, t0, t1);
            const bool mine = lane < min(n_test, 32u);
            const uint32_t slot_id = pw.list[mine ? lane : 0u];
            const float3 d = make_float3(pw.d[0][slot_id], pw.d[1][slot_id], pw.d[2][slot_id]);
            const float n0 = pw.n[0][slot_id], n1 = pw.n[1][slot_id], n2 = pw.n[2][slot_id];
            const uint32_t beg = mine ? pw.beg[slot_id] : 0u, len = mine ? pw.len[slot_id] : 0u;
            const uint32_t max_len = __reduce_max_sync(kFull, len);
            // next_crossing_t[step_axis] (grid.cpp:236-239,260) -- the step axis rule of dda_step
            const bool a2 = (n2 <= n0) && (n2 <= n1);
            const bool a1 = !a2 && (n1 <= n0);
            float bound = a2 ? n2 : (a1 ? n1 : n0);
            float best_t = FLT_MAX;
            Hit hit;
            hit.t = hit.u = hit.v = 0.0f;
            hit.tri = 0xFFFFFFFFu;
            test_pair_list<false, RCP_GUARD>(recs, beg, len, len ? len - 1u : 0u, max_len, o, d, pku, bound, best_t, hit);
            if (mine)
            {
                if (best_t != FLT_MAX) // grid.cpp:270-271: the walk ends at the first cell with a hit
                {
                    const uint32_t k = pw.ray[slot_id];
                    p.hit_tri[k] = hit.tri;
                    p.hit_t[k] = hit.t;
                    p.hit_u[k] = hit.u;
                    p.hit_v[k] = hit.v;
                    pw.status[slot_id] = kSlotFree;
                }
                else
                {
                    pw.meta[slot_id] = (pw.meta[slot_id] & ~kMetaStepsMask) | 1u; // step, then look
                    pw.status[slot_id] = kSlotWalk;
                }
            }
            __syncwarp();
        }
    }
}

} // namespace

size_t trace_pool_smem_bytes(uint32_t spp, int threads)
{
    return sizeof(float2) * ((spp + 1u) & ~1u) + (size_t) kPoolWarpBytes * (threads / 32);
}

void launch_trace_pool(const TraceParams& p, int grid_blocks, int threads, cudaStream_t stream)
{
    const size_t smem = trace_pool_smem_bytes(p.spp, threads);
    static size_t opted_in[64][2] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    const int gi = p.rcp_guard ? 1 : 0;
    if (dev < 0 || dev >= 64 || opted_in[dev][gi] < smem)
    {
        if (gi)
            cudaFuncSetAttribute(trace_pool_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        else
            cudaFuncSetAttribute(trace_pool_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        if (dev >= 0 && dev < 64)
            opted_in[dev][gi] = smem;
    }
    if (gi)
        trace_pool_kernel<true><<<grid_blocks, threads, smem, stream>>>(p);
    else
        trace_pool_kernel<false><<<grid_blocks, threads, smem, stream>>>(p);
}

} // namespace rtm
