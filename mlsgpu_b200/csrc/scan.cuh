// Block-level scans and the decoupled look-back used by every device-wide prefix sum in the
// library (the role clogs::Scan plays in the reference: src/splat_tree_cl.cpp:102,
// src/marching.cpp:353-354).
#pragma once

#include <cuda_runtime.h>
#include <cstdint>

namespace mlsb
{

constexpr uint64_t kFlagAggregate = 1ull << 62;
constexpr uint64_t kFlagPrefix = 2ull << 62;
constexpr uint64_t kFlagMask = 3ull << 62;

__device__ __forceinline__ uint32_t laneId() { return threadIdx.x & 31; }

template <typename T> __device__ __forceinline__ T warpInclusiveScan(T v)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        T t = __shfl_up_sync(0xFFFFFFFFu, v, o);
        if (laneId() >= (uint32_t) o) v += t;
    }
    return v;
}

template <typename T> __device__ __forceinline__ T warpSum(T v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

// Exclusive scan across a thread block of up to 1024 threads (blockDim.x a multiple of 32).
// sWarp must hold 33 elements.  Returns the exclusive prefix; *total gets the block sum.
template <typename T> __device__ __forceinline__ T blockExclusiveScan(T v, T *sWarp, T *total)
{
    const uint32_t lane = laneId(), warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    T inc = warpInclusiveScan(v);
    __syncthreads();   // protect sWarp from a previous use
    if (lane == 31) sWarp[warp] = inc;
    __syncthreads();
    if (warp == 0)
    {
        T w = lane < nwarps ? sWarp[lane] : T(0);
        T winc = warpInclusiveScan(w);
        sWarp[lane] = winc - w;
        if (lane == 31) sWarp[32] = winc;
    }
    __syncthreads();
    if (total) *total = sWarp[32];
    return inc - v + sWarp[warp];
}

__device__ __forceinline__ uint64_t loadStatus(const volatile uint64_t *p) { return *p; }

// Decoupled look-back (single pass chained scan).  Must be called by all 32 lanes of one
// warp; `tile` comes from an atomic ticket so that every predecessor tile is already
// running.  Publishes `aggregate` for this tile and returns the sum of the aggregates of all
// earlier tiles.  Values use the low 62 bits (two packed 31-bit counters are fine as long as
// neither total overflows its field).
__device__ __forceinline__ uint64_t lookbackExclusive(volatile uint64_t *status, uint32_t tile, uint64_t aggregate)
{
    const uint32_t lane = laneId();
    if (tile == 0)
    {
        if (lane == 0) status[0] = kFlagPrefix | aggregate;
        return 0;
    }
    if (lane == 0) status[tile] = kFlagAggregate | aggregate;
    uint64_t exclusive = 0;
    int base = (int) tile - 1;
    while (true)
    {
        int t = base - (int) lane;
        uint64_t v = kFlagPrefix;   // virtual tiles before 0: empty prefix
        if (t >= 0)
        {
            do { v = loadStatus(status + t); } while ((v & kFlagMask) == 0);
        }
        unsigned pm = __ballot_sync(0xFFFFFFFFu, (v & kFlagMask) == kFlagPrefix);
        int first = pm ? __ffs(pm) - 1 : 32;
        uint64_t contrib = ((int) lane <= first) ? (v & ~kFlagMask) : 0ull;
        exclusive += warpSum(contrib);
        if (pm) break;
        base -= 32;
    }
    if (lane == 0) status[tile] = kFlagPrefix | (exclusive + aggregate);
    return exclusive;
}

// The same for a whole thread block (blockDim.x a multiple of 32, at most 1024): every hop inspects
// blockDim.x predecessors instead of 32, so a tile in a chain of thousands of short tiles needs a
// handful of round trips to global memory instead of a hundred (with all tiles resident at once the
// aggregates are there almost immediately, and what a tile waits for is its own walk back).
// Must be called by ALL threads of the block; sRed needs 33 elements, sFirst 33.
__device__ __forceinline__ uint64_t lookbackExclusiveBlock(volatile uint64_t *status, uint32_t tile, uint64_t aggregate,
                                                           uint64_t *sRed, uint32_t *sFirst)
{
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    if (tile == 0)
    {
        if (tid == 0) status[0] = kFlagPrefix | aggregate;
        return 0;
    }
    if (tid == 0) status[tile] = kFlagAggregate | aggregate;
    uint64_t exclusive = 0;
    int base = (int) tile - 1;
    while (true)
    {
        const int t = base - (int) tid;
        uint64_t v = kFlagPrefix;   // virtual tiles before 0: empty prefix
        if (t >= 0)
        {
            do { v = loadStatus(status + t); } while ((v & kFlagMask) == 0);
        }
        const unsigned pm = __ballot_sync(0xFFFFFFFFu, (v & kFlagMask) == kFlagPrefix);
        __syncthreads();   // sFirst / sRed free again
        if (lane == 0) sFirst[warp] = pm ? warp * 32 + __ffs(pm) - 1 : 0xFFFFFFFFu;
        __syncthreads();
        uint32_t first = 0xFFFFFFFFu;
        for (uint32_t w = 0; w < nwarps; w++) first = min(first, sFirst[w]);
        const uint64_t part = warpSum(tid <= first ? (v & ~kFlagMask) : 0ull);
        if (lane == 0) sRed[warp] = part;
        __syncthreads();
        for (uint32_t w = 0; w < nwarps; w++) exclusive += sRed[w];
        if (first != 0xFFFFFFFFu) break;
        base -= (int) blockDim.x;
    }
    if (tid == 0) status[tile] = kFlagPrefix | (exclusive + aggregate);
    return exclusive;
}

}  // namespace mlsb
