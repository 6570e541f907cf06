// Stable LSD radix sort of (u32 key, u32 value) pairs — the role of clogs::Radixsort in the
// reference (src/splat_tree_cl.cpp:308, src/marching.cpp:572) — written as a one-sweep sort:
// one histogram kernel for all digits, then one kernel per 8-bit digit that ranks a
// 4096-element tile in shared memory and obtains its global offsets with a decoupled
// look-back over the tiles' digit counts.  Per pass every key/value is read once and written
// once (16 B per pair), against 3 reads + 1 write for a histogram/scan/scatter pipeline.
#include "common.cuh"
#include "scan.cuh"

namespace mlsb
{

namespace
{

constexpr int kSortThreadsSmall = 256, kSortThreadsLarge = 1024;
constexpr uint32_t kSortLargeThreshold = 3u << 20;   // elements (host-side bound) from which large tiles are used
constexpr int kSortItems = 16;
constexpr int kRadix = 256;
constexpr uint32_t kStFlagAggregate = 1u << 30;
constexpr uint32_t kStFlagPrefix = 2u << 30;
constexpr uint32_t kStFlagMask = 3u << 30;

__global__ void __launch_bounds__(256) sortHistogramKernel(const uint32_t *__restrict__ keys, uint32_t nHost,
                                                           const uint32_t *__restrict__ nDev,
                                                           uint32_t *__restrict__ hist, uint32_t passes,
                                                           uint32_t bits)
{
    const uint32_t n = nDev ? min(*nDev, nHost) : nHost;
    __shared__ uint32_t sHist[4 * kRadix];
    for (uint32_t i = threadIdx.x; i < passes * kRadix; i += blockDim.x) sHist[i] = 0;
    __syncthreads();
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    {
        uint32_t k = keys[i];
        for (uint32_t p = 0; p < passes; p++)
        {
            uint32_t w = min(8u, bits - 8 * p);
            uint32_t d = (k >> (8 * p)) & ((1u << w) - 1);
            atomicAdd(&sHist[p * kRadix + d], 1u);
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < passes * kRadix; i += blockDim.x)
        if (sHist[i]) atomicAdd(&hist[i], sHist[i]);
}

// THREADS = 256 (4096-key tiles: plenty of CTAs for the ~1 M-element sorts of a typical bin) or
// 1024 (16384-key tiles, one CTA per SM: with hundreds of small tiles in flight the look-back chain
// — one status word per tile and digit — becomes the bottleneck of multi-million-element sorts).
// EMIT: the last pass of the octree's entry sort writes the command list instead of the sorted pairs
// (writeSplatIds, kernels/octree.cl:256-279): commands[1 + position + 2 * runIndex[key]] = value, with
// keysOut = the int32 command buffer and valsOut = runIndex.
#ifndef MLSB_SORT_MINBLOCKS
#define MLSB_SORT_MINBLOCKS 4   /* 64 registers: 357 against 354 Msplats/s with 12 workers — less of the register file taken from the MLS kernels */
#endif
template <int THREADS, bool EMIT> __global__ void __launch_bounds__(THREADS, THREADS == 256 ? MLSB_SORT_MINBLOCKS : 1)
sortOnesweepKernel(const uint32_t *__restrict__ keysIn, const uint32_t *__restrict__ valsIn,
                   uint32_t *__restrict__ keysOut, uint32_t *__restrict__ valsOut, uint32_t nHost,
                   const uint32_t *__restrict__ nDev, uint32_t shift, uint32_t mask, const uint32_t *__restrict__ hist,
                   volatile uint32_t *status, uint32_t *ticket)
{
    constexpr int kSortThreads = THREADS;
    constexpr int kSortTile = THREADS * kSortItems;
    extern __shared__ uint32_t sortSmem[];
    uint32_t(*sWarpCnt)[kRadix] = reinterpret_cast<uint32_t(*)[kRadix]>(sortSmem);
    uint32_t *sKeys = sortSmem + (THREADS / 32) * kRadix;
    uint32_t *sVals = sKeys + kSortTile;
    uint32_t *sDigitStart = sVals + kSortTile;
    uint32_t *sGlobalBase = sDigitStart + kRadix;
    __shared__ uint32_t sScan[33];
    __shared__ uint32_t sTile;

    const uint32_t n = nDev ? min(*nDev, nHost) : nHost;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) sTile = atomicAdd(ticket, 1u);
    for (uint32_t i = tid; i < (kSortThreads / 32) * kRadix; i += kSortThreads) (&sWarpCnt[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = sTile;
    const uint32_t base = tile * kSortTile;
    if (base >= n) return;

    // ---- load (warp-striped: ascending index order inside a warp's 512-element span) ----
    uint32_t key[kSortItems];
    uint16_t rank[kSortItems];
    const uint32_t warpBase = base + warp * (32 * kSortItems);
#pragma unroll
    for (int k = 0; k < kSortItems; k++)
    {
        uint32_t idx = warpBase + k * 32 + lane;
        key[k] = idx < n ? keysIn[idx] : 0xFFFFFFFFu;
    }

    // ---- stable rank inside the warp: match lanes with equal digits ----
#pragma unroll
    for (int k = 0; k < kSortItems; k++)
    {
        uint32_t idx = warpBase + k * 32 + lane;
        bool valid = idx < n;
        unsigned vmask = __ballot_sync(0xFFFFFFFFu, valid);
        if (valid)
        {
            uint32_t d = (key[k] >> shift) & mask;
            unsigned peers = __match_any_sync(vmask, d);
            int leader = __ffs(peers) - 1;
            uint32_t b = 0;
            if ((int) lane == leader)
            {
                b = sWarpCnt[warp][d];
                sWarpCnt[warp][d] = b + __popc(peers);
            }
            b = __shfl_sync(peers, b, leader);
            rank[k] = (uint16_t) (b + __popc(peers & ((1u << lane) - 1)));
        }
        __syncwarp();
    }
    __syncthreads();

    // ---- per-digit totals of the tile (thread d < 256 owns digit d) ----
    const uint32_t d = tid & (kRadix - 1);
    const bool owner = tid < kRadix;
    uint32_t tileCount = 0;
    if (owner)
    {
#pragma unroll 8
        for (int w = 0; w < kSortThreads / 32; w++)
        {
            uint32_t t = sWarpCnt[w][d];
            sWarpCnt[w][d] = tileCount;
            tileCount += t;
        }
    }

    // ---- decoupled look-back on this digit's chain ----
    uint32_t exclusive = 0;
    volatile uint32_t *st = status + (size_t) tile * kRadix + d;
    if (!owner)
        ;
    else if (tile == 0)
        *st = kStFlagPrefix | tileCount;
    else
    {
        *st = kStFlagAggregate | tileCount;
        // walk back over the predecessors, kLookBack status words in flight at a time: with every
        // tile of a small sort resident at once the aggregates are there within microseconds, and what
        // the last tile waits for is its own walk (132 tiles: 33 dependent round trips at 4 in flight)
        constexpr int kLookBack = 16;
        int t = (int) tile - 1;
        bool done = false;
        while (!done)
        {
            uint32_t v[kLookBack];
#pragma unroll
            for (int k = 0; k < kLookBack; k++) v[k] = t - k >= 0 ? status[(size_t) (t - k) * kRadix + d] : (2u << 30) /* empty prefix */;
#pragma unroll
            for (int k = 0; k < kLookBack; k++)
            {
                if (done) break;
                while ((v[k] & kStFlagMask) == 0) v[k] = status[(size_t) (t - k) * kRadix + d];
                exclusive += v[k] & ~kStFlagMask;
                done = (v[k] & kStFlagMask) == kStFlagPrefix;
            }
            t -= kLookBack;
        }
        *st = kStFlagPrefix | (exclusive + tileCount);
    }

    uint32_t digitStart = blockExclusiveScan(tileCount, sScan, (uint32_t *) nullptr);
    uint32_t histPrefix = blockExclusiveScan(owner ? hist[d] : 0u, sScan, (uint32_t *) nullptr);
    if (owner)
    {
        sDigitStart[d] = digitStart;
        sGlobalBase[d] = histPrefix + exclusive - digitStart;
    }
    __syncthreads();

    // ---- scatter into shared memory in sorted order, then stream out coalesced runs ----
#pragma unroll
    for (int k = 0; k < kSortItems; k++)
    {
        uint32_t idx = warpBase + k * 32 + lane;
        if (idx < n)
        {
            uint32_t dd = (key[k] >> shift) & mask;
            uint32_t pos = sDigitStart[dd] + sWarpCnt[warp][dd] + rank[k];
            sKeys[pos] = key[k];
            sVals[pos] = valsIn ? valsIn[idx] : idx;
        }
    }
    __syncthreads();
    const uint32_t count = min((uint32_t) kSortTile, n - base);
    for (uint32_t i = tid; i < count; i += kSortThreads)
    {
        uint32_t k = sKeys[i];
        uint32_t dst = sGlobalBase[(k >> shift) & mask] + i;
        if (EMIT)
        {
            if (k != 0xFFFFFFFFu) reinterpret_cast<int32_t *>(keysOut)[1 + dst + 2 * __ldg(valsOut + k)] = (int32_t) sVals[i];
        }
        else
        {
            keysOut[dst] = k;
            valsOut[dst] = sVals[i];
        }
    }
}

}  // namespace

size_t sortScratchBytes(size_t maxItems, uint32_t maxPasses)
{
    const size_t tiles = divUp(maxItems, (size_t) (kSortThreadsSmall * kSortItems)) + 1;
    return ((size_t) maxPasses * kRadix + 64 + (size_t) maxPasses * tiles * kRadix) * sizeof(uint32_t);
}

int sortScratchAlloc(SortScratch &s, size_t maxItems, uint32_t maxPasses, size_t prefixBytes, void **prefix)
{
    s.maxItems = maxItems;
    s.maxPasses = maxPasses;
    size_t tiles = divUp(maxItems, (size_t) (kSortThreadsSmall * kSortItems)) + 1;
    s.statusWordsPerPass = tiles * kRadix;
    const int largeSmem = (int) (((kSortThreadsLarge / 32) * kRadix + 2 * kSortThreadsLarge * kSortItems + 2 * kRadix) * sizeof(uint32_t));
    MLSB_CUDA(cudaFuncSetAttribute(sortOnesweepKernel<kSortThreadsLarge, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, largeSmem));
    MLSB_CUDA(cudaFuncSetAttribute(sortOnesweepKernel<kSortThreadsLarge, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, largeSmem));
    // one allocation [prefix | hist | tickets | status] so that a sort (and its caller) needs a single memset
    prefixBytes = roundUp(prefixBytes, (size_t) 256);
    const size_t words = (size_t) maxPasses * kRadix + 64 + (size_t) maxPasses * s.statusWordsPerPass;
    MLSB_CUDA(cudaMalloc(&s.base, prefixBytes + words * sizeof(uint32_t)));
    if (prefix) *prefix = s.base;
    s.hist = reinterpret_cast<uint32_t *>(static_cast<char *>(s.base) + prefixBytes);
    s.tickets = s.hist + (size_t) maxPasses * kRadix;
    s.status = s.tickets + 64;
    return MLSB_OK;
}

void sortScratchFree(SortScratch &s)
{
    cudaFree(s.base);
    s = SortScratch();
}

namespace
{

struct SortPlan
{
    bool large;
    int threads;
    uint32_t tiles;
    size_t smem, statusStride;
};

SortPlan planSort(uint32_t n)
{
    SortPlan p;
    p.large = n >= kSortLargeThreshold;
    p.threads = p.large ? kSortThreadsLarge : kSortThreadsSmall;
    p.tiles = divUp(n, (uint32_t) (p.threads * kSortItems));
    p.smem = ((size_t) (p.threads / 32) * kRadix + 2 * (size_t) p.threads * kSortItems + 2 * kRadix) * sizeof(uint32_t);
    p.statusStride = (size_t) p.tiles * kRadix;   // status of pass q lives at s.status + q * tiles * 256: compact
    return p;
}

template <bool EMIT>
void launchPass(mlsb_worker *w, const SortPlan &pl, const uint32_t *kin, const uint32_t *vin, uint32_t *kout, uint32_t *vout,
                uint32_t n, const uint32_t *nDev, uint32_t p, uint32_t wbits)
{
    SortScratch &s = w->sort;
    if (pl.large)
        sortOnesweepKernel<kSortThreadsLarge, EMIT><<<pl.tiles, kSortThreadsLarge, pl.smem, w->stream>>>(
            kin, vin, kout, vout, n, nDev, 8 * p, (1u << wbits) - 1, s.hist + p * kRadix, s.status + p * pl.statusStride, s.tickets + p);
    else
        sortOnesweepKernel<kSortThreadsSmall, EMIT><<<pl.tiles, kSortThreadsSmall, pl.smem, w->stream>>>(
            kin, vin, kout, vout, n, nDev, 8 * p, (1u << wbits) - 1, s.hist + p * kRadix, s.status + p * pl.statusStride, s.tickets + p);
}

}  // namespace

size_t sortScratchBytesUsed(const SortScratch &s, uint32_t n, uint32_t passes)
{
    return ((size_t) s.maxPasses * kRadix + 64 + passes * planSort(n).statusStride) * sizeof(uint32_t);
}

int radixSortToCommands(mlsb_worker *w, uint32_t *keys[2], uint32_t *values[2], uint32_t n, const uint32_t *nDev,
                        uint32_t bits, int32_t *commands, const uint32_t *runIndex, const char *statName)
{
    SortScratch &s = w->sort;
    const uint32_t passes = divUp(bits, 8u);
    MLSB_REQUIRE(n > 0 && passes >= 1 && n <= s.maxItems && passes <= s.maxPasses && passes <= 4, MLSB_ERR_LENGTH,
                 "sort scratch too small");
    ScopedStat stat(w, statName, passes);
    const SortPlan pl = planSort(n);
    int cur = 0;
    for (uint32_t p = 0; p < passes; p++)
    {
        const uint32_t wbits = bits - 8 * p < 8 ? bits - 8 * p : 8;
        if (p + 1 == passes)
            launchPass<true>(w, pl, keys[cur], values[cur], reinterpret_cast<uint32_t *>(commands), const_cast<uint32_t *>(runIndex),
                             n, nDev, p, wbits);
        else
            launchPass<false>(w, pl, keys[cur], values[cur], keys[cur ^ 1], values[cur ^ 1], n, nDev, p, wbits);
        cur ^= 1;
    }
    w->launches += passes;
    MLSB_CUDA(cudaGetLastError());
    return MLSB_OK;
}

int radixSortPairs(mlsb_worker *w, uint32_t *keys[2], uint32_t *values[2], uint32_t n, uint32_t bits,
                   int *result, const char *statName)
{
    return radixSortPairsDeviceCount(w, keys, values, n, nullptr, bits, result, statName, false);
}

int radixSortPairsDeviceCount(mlsb_worker *w, uint32_t *keys[2], uint32_t *values[2], uint32_t n,
                              const uint32_t *nDev, uint32_t bits, int *result, const char *statName,
                              bool identityValues)
{
    *result = 0;
    if (n == 0 || bits == 0) return MLSB_OK;
    SortScratch &s = w->sort;
    uint32_t passes = divUp(bits, 8u);
    MLSB_REQUIRE(n <= s.maxItems && passes <= s.maxPasses && passes <= 4, MLSB_ERR_LENGTH, "sort scratch too small");
    ScopedStat stat(w, statName, 1 + passes);
    const SortPlan pl = planSort(n);
    MLSB_CUDA(cudaMemsetAsync(s.hist, 0, sortScratchBytesUsed(s, n, passes), w->stream));
    uint32_t histBlocks = min(divUp(n, 256u * 8u), 148u * 8u);
    sortHistogramKernel<<<histBlocks, 256, 0, w->stream>>>(keys[0], n, nDev, s.hist, passes, bits);
    int cur = 0;
    for (uint32_t p = 0; p < passes; p++)
    {
        uint32_t wbits = bits - 8 * p < 8 ? bits - 8 * p : 8;
        const uint32_t *vin = (p == 0 && identityValues) ? nullptr : values[cur];
        launchPass<false>(w, pl, keys[cur], vin, keys[cur ^ 1], values[cur ^ 1], n, nDev, p, wbits);
        cur ^= 1;
    }
    w->launches += 1 + passes;
    MLSB_CUDA(cudaGetLastError());
    *result = cur;
    return MLSB_OK;
}

}  // namespace mlsb
