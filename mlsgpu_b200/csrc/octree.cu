// GPU octree build: SplatTreeCL::enqueueBuild (src/splat_tree_cl.cpp:269-335) and
// kernels/octree.cl, re-designed for one resident bin:
//
//   writeEntries  splat -> 8 (key, id) slots                         kernels/octree.cl:159-214
//   radix sort    stable, low 3*(maxShift-minShift)+1 key bits       src/splat_tree_cl.cpp:308
//   findRuns      first/last sorted position of every key            (replaces countCommands)
//   scanRuns      prefix count of non-empty keys over the KEY space  (replaces the 8N-element
//                 clogs scan: commandMap[i] = 1 + i + 2*runIndex[key[i]], so only numStart
//                 elements need scanning, not 8N)
//   emitCommands  commands[commandMap[i]] = id[i]                    kernels/octree.cl:256-279
//   linkLevels    start[] / jump slots for all levels in one launch  kernels/octree.cl:292-341
//                 (each cell finds its nearest non-empty ancestor directly, so the six
//                 level-serial launches of src/splat_tree_cl.cpp:320-331 collapse into one)
//
// The resulting start[] and commands[] arrays are bit-identical to the reference's.
#include "common.cuh"
#include "scan.cuh"

namespace mlsb
{

namespace
{

struct LevelInfo
{
    uint32_t offsets[kMaxLevels + 22];   // indexed by shift; valid for minShift..maxShift
    uint32_t minShift, maxShift;
};

// kernels/octree.cl:121-135 via bit spreading: x in bit 0, y in bit 1, z in bit 2 of each triple
__device__ __forceinline__ uint32_t spread3(uint32_t v)
{
    v &= 0x3FF;
    v = (v | (v << 16)) & 0x030000FF;
    v = (v | (v << 8)) & 0x0300F00F;
    v = (v | (v << 4)) & 0x030C30C3;
    v = (v | (v << 2)) & 0x09249249;
    return v;
}

__device__ __forceinline__ uint32_t makeCode(int x, int y, int z)
{
    return spread3((uint32_t) x) | (spread3((uint32_t) y) << 1) | (spread3((uint32_t) z) << 2);
}

// kernels/octree.cl:60-66 with dot() evaluated left to right, unfused (the oracle's choice)
__device__ __forceinline__ float pointBoxDist2(float px, float py, float pz, float lx, float ly, float lz,
                                                float hx, float hy, float hz)
{
    float dx = __fsub_rn(fmaxf(lx, fminf(hx, px)), px);
    float dy = __fsub_rn(fmaxf(ly, fminf(hy, py)), py);
    float dz = __fsub_rn(fmaxf(lz, fminf(hz, pz)), pz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// prepare() + the 2x2x2 candidate loop of writeEntries, kernels/octree.cl:79-110,184-212.
// The candidates are ilo + {0,1}^3, so the per-axis box distances (and Morton bit spreads) are
// computed once for the two options of each axis and combined per candidate; the squared
// distance is still summed as (dx*dx + dy*dy) + dz*dz, one rounding per operation, exactly like
// pointBoxDist2() above.  CODES = false only counts the accepted candidates.
template <bool CODES>
__device__ __forceinline__ uint32_t splatEntries(const float4 pr, const int3 bias, const LevelInfo &lv, uint32_t k[8])
{
    const float p[3] = {pr.x, pr.y, pr.z};
    const int b[3] = {bias.x, bias.y, bias.z};
    int lo[3], big = 0;
#pragma unroll
    for (int a = 0; a < 3; a++)
    {
        lo[a] = __float2int_rd(__fsub_rn(p[a], pr.w));
        big = max(big, __float2int_rd(__fadd_rn(p[a], pr.w)) - lo[a]);
    }
    int shift = big > 1 ? 32 - __clz(big - 1) : 0;   // levelShift(), :50-55
    shift = min(max(shift, (int) lv.minShift), (int) lv.maxShift);
    const float radius2 = __fmul_rn(__fmul_rn(pr.w, pr.w), 1.00001f);
    const int bound = 1 << (lv.maxShift - shift);
    float sq[3][2];
    uint32_t bits[3][2];
    bool inside[3][2];
#pragma unroll
    for (int a = 0; a < 3; a++)
    {
        const int il = max(lo[a] - b[a], 0) >> shift;
#pragma unroll
        for (int o = 0; o < 2; o++)
        {
            const int c = il + o;
            // goodEntry(), :100-110: box [c << shift, (c+1) << shift] + bias along this axis
            const float blo = (float) ((int) ((uint32_t) c << shift) + b[a]);
            const float bhi = (float) ((int) ((uint32_t) (c + 1) << shift) + b[a]);
            const float d = __fsub_rn(fmaxf(blo, fminf(bhi, p[a])), p[a]);
            sq[a][o] = __fmul_rn(d, d);
            inside[a][o] = c < bound;
            if (CODES) bits[a][o] = spread3((uint32_t) c) << a;
        }
    }
    const uint32_t levelOffset = CODES ? lv.offsets[shift] : 0u;
    uint32_t cnt = 0;
#pragma unroll
    for (int o = 0; o < 8; o++)
    {
        const int ox = o & 1, oy = (o >> 1) & 1, oz = o >> 2;
        const float dist2 = __fadd_rn(__fadd_rn(sq[0][ox], sq[1][oy]), sq[2][oz]);
        const bool isect = dist2 < radius2 && inside[0][ox] && inside[1][oy] && inside[2][oz];
        if (CODES) k[o] = isect ? (bits[0][ox] | bits[1][oy] | bits[2][oz]) + levelOffset : 0xFFFFFFFFu;
        cnt += isect;
    }
    return cnt;
}

// The reference writes exactly 8 (key, id) slots per splat, UINT_MAX in the rejected ones, and sorts
// all 8N.  Here only the accepted entries are written, packed in (splat, slot) order, and their
// number goes to *numValid: the sort, the run detection and the command emission then touch 2-4x
// fewer elements, with identical results (rejected slots sort to the end and are skipped by
// writeSplatIds, kernels/octree.cl:266).  Three launches, none of which waits on another CTA:
// EMIT = false counts per 256-splat tile, scanCountsKernel turns the counts into bases, EMIT = true
// recomputes the (cheap) candidates and writes them.
template <bool EMIT>
__global__ void __launch_bounds__(256) writeEntriesKernel(uint32_t *__restrict__ keys, uint32_t *__restrict__ values,
                                                          mlsb_splat *__restrict__ splats, int3 bias, LevelInfo lv,
                                                          uint32_t firstSplat, uint32_t numSplats,
                                                          uint32_t *__restrict__ tileCounts)
{
    __shared__ uint32_t sScan[33];
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t k[8];
    uint32_t cnt = 0;
    const uint32_t gid = g + firstSplat;
    if (g < numSplats)
    {
        const float4 pr = *reinterpret_cast<const float4 *>(&splats[gid]);
        cnt = splatEntries<EMIT>(pr, bias, lv, k);
        if (EMIT) splats[gid].radius = __fdiv_rn(1.0f, __fmul_rn(pr.w, pr.w));   // kernels/octree.cl:193
    }
    uint32_t total;
    const uint32_t excl = blockExclusiveScan(cnt, sScan, &total);
    if (!EMIT)
    {
        if (threadIdx.x == 0) tileCounts[blockIdx.x] = total;
        return;
    }
    uint32_t dst = tileCounts[blockIdx.x] + excl;
    if (cnt)
    {
#pragma unroll
        for (int o = 0; o < 8; o++)
            if (k[o] != 0xFFFFFFFFu)
            {
                keys[dst] = k[o];
                values[dst] = gid;
                dst++;
            }
    }
}

// In-place exclusive scan of n per-tile counts by one CTA; the grand total goes to *total.
__global__ void __launch_bounds__(1024) scanCountsKernel(uint32_t *counts, uint32_t n, uint32_t *total)
{
    __shared__ uint32_t sScan[33];
    const uint32_t per = (n + 1023) / 1024;
    const uint32_t first = threadIdx.x * per, last = min(first + per, n);
    uint32_t sum = 0;
    for (uint32_t i = first; i < last; i++) sum += counts[i];
    uint32_t all;
    uint32_t base = blockExclusiveScan(sum, sScan, &all);
    for (uint32_t i = first; i < last; i++)
    {
        const uint32_t c = counts[i];
        counts[i] = base;
        base += c;
    }
    if (threadIdx.x == 0) *total = all;
}

// First and last sorted position of each key (boundaries of equal-key runs).
__global__ void __launch_bounds__(256) findRunsKernel(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ numValid,
                                                      int32_t *__restrict__ runFirst, int32_t *__restrict__ runLast)
{
    const uint32_t numEntries = *numValid;
    for (uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x; pos < numEntries; pos += gridDim.x * blockDim.x)
    {
        uint32_t cur = keys[pos];
        if (cur == 0xFFFFFFFFu) continue;
        uint32_t prev = pos > 0 ? keys[pos - 1] : 0xFFFFFFFFu;
        uint32_t next = pos + 1 < numEntries ? keys[pos + 1] : 0xFFFFFFFFu;
        if (prev != cur) runFirst[cur] = (int32_t) pos;
        if (cur != next) runLast[cur] = (int32_t) pos;
    }
}

// Device-wide exclusive count of non-empty keys, then the command positions of each run:
//   runStart[key] = first + 2*runIndex            (slot holding the `end` pointer)
//   jumpPos[key]  = last + 2*runIndex + 2         (slot holding the jump to the ancestor)
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
__global__ void __launch_bounds__(kScanThreads)
scanRunsKernel(const int32_t *__restrict__ runFirst, const int32_t *__restrict__ runLast, uint32_t numStart,
               uint32_t *__restrict__ runIndex, int32_t *__restrict__ runStart, int32_t *__restrict__ jumpPos,
               volatile uint64_t *status, uint32_t *ticket)
{
    __shared__ uint32_t sScan[33];
    __shared__ uint32_t sTile;
    __shared__ uint64_t sPrefix;
    if (threadIdx.x == 0) sTile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = sTile;
    const uint32_t base = tile * (kScanThreads * kScanItems) + threadIdx.x * kScanItems;
    int32_t first[kScanItems];
    uint32_t cnt = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; i++)
    {
        first[i] = base + i < numStart ? runFirst[base + i] : -1;
        cnt += first[i] >= 0;
    }
    uint32_t total;
    uint32_t excl = blockExclusiveScan(cnt, sScan, &total);
    if (threadIdx.x < 32)
    {
        uint64_t p = lookbackExclusive(status, tile, total);
        if (threadIdx.x == 0) sPrefix = p;
    }
    __syncthreads();
    uint32_t run = (uint32_t) sPrefix + excl;
#pragma unroll
    for (int i = 0; i < kScanItems; i++)
    {
        uint32_t key = base + i;
        if (key < numStart)
        {
            runIndex[key] = run;
            if (first[i] >= 0)
            {
                runStart[key] = first[i] + 2 * (int32_t) run;
                jumpPos[key] = runLast[key] + 2 * (int32_t) run + 2;
                run++;
            }
            else
            {
                runStart[key] = -1;
                jumpPos[key] = -1;
            }
        }
    }
}

// writeSplatIds, kernels/octree.cl:256-279, with commandMap computed in place.
__global__ void __launch_bounds__(256) emitCommandsKernel(int32_t *__restrict__ commands, const uint32_t *__restrict__ keys,
                                                          const uint32_t *__restrict__ ids,
                                                          const uint32_t *__restrict__ runIndex,
                                                          const uint32_t *__restrict__ numValid)
{
    const uint32_t numEntries = *numValid;
    for (uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x; pos < numEntries; pos += gridDim.x * blockDim.x)
    {
        uint32_t key = keys[pos];
        if (key == 0xFFFFFFFFu) continue;
        commands[1 + pos + 2 * runIndex[key]] = (int32_t) ids[pos];
    }
}

// writeStart / writeStartTop for every level at once (kernels/octree.cl:292-341).
__global__ void __launch_bounds__(256) linkLevelsKernel(int32_t *__restrict__ start, int32_t *__restrict__ commands,
                                                        const int32_t *__restrict__ runStart,
                                                        const int32_t *__restrict__ jumpPos, LevelInfo lv,
                                                        uint32_t numStart)
{
    uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= numStart) return;
    // which level does this key belong to?  offsets grow from the finest level (minShift)
    uint32_t level = lv.minShift;
    while (level < lv.maxShift && pos >= lv.offsets[level + 1]) level++;
    uint32_t code = pos - lv.offsets[level];
    // nearest non-empty strict ancestor
    int32_t prev = -1;
    uint32_t l = level, c = code;
    while (l < lv.maxShift)
    {
        l++;
        c >>= 3;
        int32_t s = runStart[lv.offsets[l] + c];
        if (s >= 0) { prev = s; break; }
    }
    int32_t jp = jumpPos[pos];
    if (jp >= 0)
    {
        int32_t s = runStart[pos];
        commands[jp] = prev;
        commands[s] = jp;
        start[pos] = s;
    }
    else
        start[pos] = prev;
}

}  // namespace

int octreeAlloc(mlsb_worker *w)
{
    const mlsb_worker_config &c = w->cfg;
    // src/splat_tree_cl.cpp:114-129
    uint64_t maxStart = (1ull << (3 * c.maxLevels)) / 7;
    uint64_t maxEntries = 8ull * c.maxSplats;
    uint64_t maxRanges = maxStart < maxEntries ? maxStart : maxEntries;
    MLSB_REQUIRE(maxStart < (1ull << 31) && maxEntries + 2 * maxRanges < (1ull << 31), MLSB_ERR_LENGTH,
                 "octree too large");
    w->numStartMax = (uint32_t) maxStart;
    w->numCommandsMax = (uint32_t) (maxEntries + 2 * maxRanges);
    MLSB_CUDA(cudaMalloc(&w->start, maxStart * sizeof(int32_t)));
    MLSB_CUDA(cudaMalloc(&w->runStart, maxStart * sizeof(int32_t)));
    MLSB_CUDA(cudaMalloc(&w->jumpPos, maxStart * sizeof(int32_t)));
    MLSB_CUDA(cudaMalloc(&w->runFirst, maxStart * sizeof(int32_t)));
    MLSB_CUDA(cudaMalloc(&w->runLast, maxStart * sizeof(int32_t)));
    MLSB_CUDA(cudaMalloc(&w->runIndex, maxStart * sizeof(uint32_t)));
    MLSB_CUDA(cudaMalloc(&w->commands, (size_t) w->numCommandsMax * sizeof(int32_t)));
    for (int i = 0; i < 2; i++)
    {
        MLSB_CUDA(cudaMalloc(&w->entryKeys[i], (maxEntries + 8) * sizeof(uint32_t)));
        MLSB_CUDA(cudaMalloc(&w->entryValues[i], (maxEntries + 8) * sizeof(uint32_t)));
    }
    return MLSB_OK;
}

void octreeFree(mlsb_worker *w)
{
    cudaFree(w->start);
    cudaFree(w->runStart);
    cudaFree(w->jumpPos);
    cudaFree(w->runFirst);
    cudaFree(w->runLast);
    cudaFree(w->runIndex);
    cudaFree(w->commands);
    for (int i = 0; i < 2; i++)
    {
        cudaFree(w->entryKeys[i]);
        cudaFree(w->entryValues[i]);
    }
}

int octreeBuild(mlsb_worker *w, uint32_t firstSplat, uint32_t numSplats, const int32_t offset[3],
                uint32_t subsamplingShift)
{
    const uint32_t maxLevels = w->cfg.maxLevels;
    uint32_t maxShift = maxLevels + subsamplingShift - 1;
    uint32_t minShift = subsamplingShift < maxShift ? subsamplingShift : maxShift;
    MLSB_REQUIRE(maxShift < 32, MLSB_ERR_INVALID_ARGUMENT, "levels + subsampling too large");
    LevelInfo lv{};
    lv.minShift = minShift;
    lv.maxShift = maxShift;
    uint32_t pos = 0;
    for (uint32_t i = minShift; i <= maxShift; i++)
    {
        lv.offsets[i] = pos;
        pos += 1u << (3 * (maxShift - i));
    }
    const uint32_t numStart = pos;
    MLSB_REQUIRE(numStart <= w->numStartMax, MLSB_ERR_LENGTH, "start array too small");
    w->numStart = numStart;
    cudaStream_t s = w->stream;
    const uint32_t numEntries = numSplats * 8;

    if (numSplats > 0)
    {
        {
            ScopedStat st(w, "kernel.octree.writeEntries.time");
            const uint32_t tiles = divUp(numSplats, 256u);
            MLSB_REQUIRE(tiles <= 2 * w->scanStatusWords, MLSB_ERR_LENGTH, "scan scratch too small");
            uint32_t *tileCounts = reinterpret_cast<uint32_t *>(w->scanStatus);
            const int3 bias = make_int3(offset[0], offset[1], offset[2]);
            writeEntriesKernel<false><<<tiles, 256, 0, s>>>(nullptr, nullptr, w->splats, bias, lv, firstSplat, numSplats, tileCounts);
            scanCountsKernel<<<1, 1024, 0, s>>>(tileCounts, tiles, w->scanTicket + 2);
            writeEntriesKernel<true><<<tiles, 256, 0, s>>>(w->entryKeys[0], w->entryValues[0], w->splats, bias, lv, firstSplat,
                                                          numSplats, tileCounts);
        }
    }
    int cur = 0;
    if (numSplats == 0) MLSB_CUDA(cudaMemsetAsync(w->scanTicket, 0, 4 * sizeof(uint32_t), s));
    const uint32_t *numValid = w->scanTicket + 2;   // device-side count of accepted entries
    MLSB_TRY(radixSortPairsDeviceCount(w, w->entryKeys, w->entryValues, numEntries, numValid,
                                       3 * (maxShift - minShift) + 1, &cur, "kernel.octree.sort.time", false));
    {
        ScopedStat st(w, "kernel.octree.fill.time", 0);
        MLSB_CUDA(cudaMemsetAsync(w->runFirst, 0xFF, (size_t) numStart * sizeof(int32_t), s));
    }
    if (numEntries > 0)
    {
        ScopedStat st(w, "kernel.octree.countCommands.time");
        findRunsKernel<<<min(divUp(numEntries, 256u), 148u * 16u), 256, 0, s>>>(w->entryKeys[cur], numValid, w->runFirst, w->runLast);
    }
    {
        ScopedStat st(w, "kernel.octree.scan.time");
        uint32_t tiles = divUp(numStart, (uint32_t) (kScanThreads * kScanItems));
        MLSB_REQUIRE(tiles <= w->scanStatusWords, MLSB_ERR_LENGTH, "scan status too small");
        MLSB_CUDA(cudaMemsetAsync(w->scanStatus, 0, (size_t) tiles * sizeof(uint64_t), s));
        MLSB_CUDA(cudaMemsetAsync(w->scanTicket, 0, sizeof(uint32_t), s));   // ticket 0 only: [2] holds numValid
        scanRunsKernel<<<tiles, kScanThreads, 0, s>>>(w->runFirst, w->runLast, numStart, w->runIndex, w->runStart,
                                                      w->jumpPos, w->scanStatus, w->scanTicket);
    }
    if (numEntries > 0)
    {
        ScopedStat st(w, "kernel.octree.writeSplatIds.time");
        emitCommandsKernel<<<min(divUp(numEntries, 256u), 148u * 16u), 256, 0, s>>>(w->commands, w->entryKeys[cur], w->entryValues[cur],
                                                                   w->runIndex, numValid);
    }
    {
        ScopedStat st(w, "kernel.octree.writeStart.time");
        linkLevelsKernel<<<divUp(numStart, 256u), 256, 0, s>>>(w->start, w->commands, w->runStart, w->jumpPos, lv, numStart);
    }
    w->launches += (numSplats > 0 ? 5 : 0) + 2;
    MLSB_CUDA(cudaGetLastError());
    return MLSB_OK;
}

}  // namespace mlsb
