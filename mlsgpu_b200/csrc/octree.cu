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
// With the default six levels the key space (37 449 cells) fits the shared memory of one cluster ("planned"
// paths below).  The default is the BUCKET path further down (five launches, one memset, no CTA ever waits on
// another); MLSB_OCTREE_SORT=onesweep selects the one-sweep path described here (four launches, one memset):
//
//   writeEntriesFused   splats read ONCE: candidates, tile base by decoupled look-back, packed
//                       (key, id) pairs, 1/r^2 and the MLS hit threshold per splat, and the number of
//                       entries of every key (warp-aggregated atomics)
//   planKernel          one thread-block cluster (8 CTAs, distributed shared memory) over the key counts: run starts, run indices, start[], the
//                       end/jump slots of every range AND the ancestor links, plus the sort's digit
//                       histograms — none of it needs the sorted entries
//   sort pass(es)       the last one writes the command list directly (no sorted pairs, no emit pass)
//
// Deeper octrees (levels > 6: the reference's tests use 8 and 9) take the general path further down.
// The resulting start[] and commands[] arrays are bit-identical to the reference's.
#include "common.cuh"
#include "scan.cuh"
#include <cooperative_groups.h>
#include <algorithm>
#include <cstdlib>
#include <cstring>

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
// number goes to counters[1]: the sort and the command emission then touch 2-4x fewer elements, with
// identical results (rejected slots sort to the end and are skipped by writeSplatIds,
// kernels/octree.cl:266).  One launch: a tile of 256 splats computes its candidates, obtains its base
// with a decoupled look-back over the tiles in front of it (tiles are numbered by an atomic ticket,
// so every predecessor is running) and writes.  keyCount != nullptr: entries per key as well, one
// atomic per distinct key and warp.
__global__ void __launch_bounds__(256) writeEntriesFusedKernel(uint32_t *__restrict__ keys, uint32_t *__restrict__ values,
                                                               mlsb_splat *__restrict__ splats, float *__restrict__ thresholds,
                                                               int3 bias, LevelInfo lv, uint32_t firstSplat, uint32_t numSplats,
                                                               volatile uint64_t *status, uint32_t *counters,
                                                               uint32_t *__restrict__ keyCount)
{
    __shared__ uint32_t sScan[33];
    __shared__ uint64_t sRed[33];
    __shared__ uint32_t sFirst[33];
    __shared__ uint32_t sTile;
    if (threadIdx.x == 0) sTile = atomicAdd(&counters[0], 1u);
    __syncthreads();
    const uint32_t tile = sTile;
    const uint32_t g = tile * blockDim.x + threadIdx.x;
    uint32_t k[8];
    uint32_t cnt = 0;
    const uint32_t gid = g + firstSplat;
#pragma unroll
    for (int o = 0; o < 8; o++) k[o] = 0xFFFFFFFFu;
    if (g < numSplats)
    {
        const float4 pr = *reinterpret_cast<const float4 *>(&splats[gid]);
        cnt = splatEntries<true>(pr, bias, lv, k);
        const float invR2 = __fdiv_rn(1.0f, __fmul_rn(pr.w, pr.w));
        splats[gid].radius = invR2;   // kernels/octree.cl:193
        thresholds[gid] = hitThreshold(invR2);
    }
    if (keyCount != nullptr)
    {
        const uint32_t lane = threadIdx.x & 31;
#pragma unroll
        for (int o = 0; o < 8; o++)
        {
            const unsigned peers = __match_any_sync(0xFFFFFFFFu, k[o]);
            if (k[o] != 0xFFFFFFFFu && lane == (uint32_t) __ffs(peers) - 1) atomicAdd(&keyCount[k[o]], (uint32_t) __popc(peers));
        }
    }
    uint32_t total;
    const uint32_t excl = blockExclusiveScan(cnt, sScan, &total);
    const uint32_t base = (uint32_t) lookbackExclusiveBlock(status, tile, total, sRed, sFirst);
    if ((tile + 1) * blockDim.x >= numSplats && tile * blockDim.x < numSplats && threadIdx.x == 0) counters[1] = base + total;
    uint32_t dst = base + excl;
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

// Everything the command list needs except the splat ids, from the per-key entry counts alone, by one
// cluster that holds the key space in (distributed) shared memory (numStart <= kPlanMaxKeys):
//   runStart[key] = (entries of smaller keys) + 2 * (non-empty smaller keys)     slot of the range's `end` pointer
//   jump slot     = runStart + count + 1                                        kernels/octree.cl:256-279
//   start[], commands[runStart] = jump slot, commands[jump slot] = nearest non-empty ancestor's
//   runStart or -1                                                              kernels/octree.cl:292-341
//   runIndex[key] (the sort's last pass places id i of the sorted order at 1 + i + 2 * runIndex[key])
//   digit histograms of the radix passes, hist[p][256]
// One SM cannot do it fast enough (160 k warp instructions: 20 us of issue slots on one SM), so the
// key space is spread over a thread-block CLUSTER of eight CTAs: every CTA keeps the run starts of
// its 1/8 of the keys in its own shared memory, the exclusive totals of the CTAs in front and the
// run starts of ancestors in other CTAs' ranges are read through distributed shared memory, and
// the three phases are separated by the cluster's hardware barrier — no global scratch, no
// second launch.
constexpr int kPlanCluster = 8;
constexpr int kPlanThreads = 1024;
constexpr int kPlanPer = 5;                                                     // keys per thread
constexpr uint32_t kPlanChunkMax = kPlanThreads * kPlanPer;                     // keys per CTA
constexpr uint32_t kPlanMaxKeys = kPlanCluster * kPlanChunkMax;                 // 40 960 >= 37 449 (six levels)
__global__ void __cluster_dims__(kPlanCluster, 1, 1) __launch_bounds__(kPlanThreads)
planKernel(uint32_t *__restrict__ keyCount, uint32_t numStart, LevelInfo lv, uint32_t bits, uint32_t *__restrict__ runIndex,
           int32_t *__restrict__ start, int32_t *__restrict__ commands, uint32_t *__restrict__ hist, uint32_t *numValid)
{
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ uint32_t sRun[kPlanChunkMax];   // entry counts, then the run start of every key of this CTA (0xFFFFFFFF = empty)
    __shared__ uint32_t sHist[2 * 256];
    __shared__ uint64_t sScan[33];
    __shared__ uint64_t sTotal;                // this CTA's (entries | runs << 32), read by the CTAs behind it
    const uint32_t tid = threadIdx.x, rank = cluster.block_rank();
    const uint32_t chunk = (numStart + kPlanCluster - 1) / kPlanCluster;   // keys per CTA
    const uint32_t base = min(rank * chunk, numStart), count = min(base + chunk, numStart) - base;
    if (tid < 512) sHist[tid] = 0;
    // counts: key base + i * 1024 + tid in register c[i] — coalesced, all loads in flight at once
    uint32_t c[kPlanPer];
#pragma unroll
    for (int i = 0; i < kPlanPer; i++)
    {
        const uint32_t k = i * kPlanThreads + tid;
        c[i] = k < count ? keyCount[base + k] : 0u;
    }
#pragma unroll
    for (int i = 0; i < kPlanPer; i++)
    {
        const uint32_t k = i * kPlanThreads + tid;
        if (k < count) sRun[k] = c[i];
    }
    __syncthreads();
    // the scan runs over consecutive keys per thread, out of shared memory
    const uint32_t per = (count + kPlanThreads - 1) / kPlanThreads;
    const uint32_t first = min(tid * per, count), last = min(first + per, count);
    const uint32_t mask0 = bits >= 8 ? 0xFFu : (1u << bits) - 1;
    const uint32_t mask1 = bits > 8 ? (1u << min(bits - 8, 8u)) - 1 : 0u;
    uint32_t entries = 0, runs = 0;
    for (uint32_t k = first; k < last; k++)
    {
        const uint32_t n = sRun[k], key = base + k;
        if (n)
        {
            entries += n;
            runs++;
            atomicAdd(&sHist[key & mask0], n);
            if (bits > 8) atomicAdd(&sHist[256 + ((key >> 8) & mask1)], n);
        }
    }
    uint64_t total;
    const uint64_t excl = blockExclusiveScan((uint64_t) entries | ((uint64_t) runs << 32), sScan, &total);
    if (tid == 0) sTotal = total;
    cluster.sync();
    uint64_t before = 0, all = 0;
    for (uint32_t r = 0; r < kPlanCluster; r++)
    {
        const uint64_t t = *cluster.map_shared_rank(&sTotal, r);
        all += t;
        if (r < rank) before += t;
    }
    uint32_t e = (uint32_t) (excl + before), r = (uint32_t) ((excl + before) >> 32);
    for (uint32_t k = first; k < last; k++)
    {
        const uint32_t n = sRun[k];
        if (n)
        {
            runIndex[base + k] = r;   // only read for keys that have entries: a few thousand scattered stores
            sRun[k] = e + 2 * r;
            e += n;
            r++;
        }
        else
            sRun[k] = 0xFFFFFFFFu;
    }
    cluster.sync();
#pragma unroll
    for (int i = 0; i < kPlanPer; i++)
    {
        const uint32_t k = i * kPlanThreads + tid;
        if (k >= count) continue;
        const uint32_t pos = base + k;
        // which level does this key belong to?  offsets grow from the finest level (minShift)
        uint32_t level = lv.minShift;
        while (level < lv.maxShift && pos >= lv.offsets[level + 1]) level++;
        // nearest non-empty strict ancestor; the coarser levels live in the last CTAs of the cluster
        int32_t prev = -1;
        uint32_t l = level, code = pos - lv.offsets[level];
        while (l < lv.maxShift)
        {
            l++;
            code >>= 3;
            const uint32_t idx = lv.offsets[l] + code, owner = idx / chunk;
            const uint32_t a = cluster.map_shared_rank(sRun, owner)[idx - owner * chunk];
            if (a != 0xFFFFFFFFu) { prev = (int32_t) a; break; }
        }
        if (c[i])
        {
            const uint32_t s = sRun[k];
            const int32_t jp = (int32_t) (s + c[i] + 1);
            commands[jp] = prev;
            commands[s] = jp;
            start[pos] = (int32_t) s;
            keyCount[pos] = 0;   // clean for the next build: no memset of the counts
        }
        else
            start[pos] = prev;
    }
    if (tid < 512 && sHist[tid]) atomicAdd(&hist[tid], sHist[tid]);   // hist was zeroed with the sort's scratch
    if (rank == 0 && tid == 0) *numValid = (uint32_t) all;
    cluster.sync();   // no CTA may leave while its shared memory can still be read
}

// ---- the entry sort without any CTA waiting on another (MLSB_OCTREE_SORT=bucket) -------------------
// The one-sweep passes above spend most of their life in the look-back: every tile waits for the
// digit counts of the tiles in front of it, 25 us per pass whatever the size, and a CTA that waits
// holds its share of the SM all the while.  But the number of entries of every KEY is
// known before anything is sorted (keyCount), and with it every key's final run of command slots
// (start[]).  So the sort becomes two scatters whose destinations are computed, not negotiated:
//
//   countEntriesKernel          (replaces the counting half of the fused entry kernel) per tile of 256
//                               splats: entries per key (-> keyCount) and per BUCKET of 256 keys
//   planKernel                  as before; its high-byte histogram is the size of every bucket
//   bucketPrefixKernel          one CTA per bucket: base of every tile's entries inside the bucket
//   writeEntriesBucketedKernel  candidates again (cheap), ranked inside the tile, written straight to
//                               their bucket in (tile, splat, slot) order: a stable partition by high byte
//   bucketSortEmitKernel        one CTA per bucket: stable ranking by low byte with running counters,
//                               id written to start[key] + 1 + rank — the command list
//
// Five launches instead of four, none of which waits.  (A bucket is walked by ONE CTA in chunks of
// 2048 entries: an input with millions of entries in one 256-key bucket serialises there.)
constexpr uint32_t kBucketBits = 8;
constexpr uint32_t kMaxBuckets = kPlanMaxKeys >> kBucketBits;   // 160

__global__ void __launch_bounds__(256) countEntriesKernel(const mlsb_splat *__restrict__ splats, int3 bias, LevelInfo lv,
                                                          uint32_t firstSplat, uint32_t numSplats, uint32_t numBuckets,
                                                          uint32_t tileStride, uint32_t *__restrict__ keyCount,
                                                          uint32_t *__restrict__ tileBucket)
{
    __shared__ uint32_t sBucket[kMaxBuckets];
    const uint32_t tile = blockIdx.x, lane = threadIdx.x & 31;
    for (uint32_t i = threadIdx.x; i < numBuckets; i += blockDim.x) sBucket[i] = 0;
    __syncthreads();
    const uint32_t g = tile * blockDim.x + threadIdx.x;
    uint32_t k[8];
#pragma unroll
    for (int o = 0; o < 8; o++) k[o] = 0xFFFFFFFFu;
    if (g < numSplats) splatEntries<true>(*reinterpret_cast<const float4 *>(&splats[g + firstSplat]), bias, lv, k);
#pragma unroll
    for (int o = 0; o < 8; o++)
    {
        const unsigned peers = __match_any_sync(0xFFFFFFFFu, k[o]);
        if (k[o] != 0xFFFFFFFFu && lane == (uint32_t) __ffs(peers) - 1)
        {
            const uint32_t c = (uint32_t) __popc(peers);
            atomicAdd(&keyCount[k[o]], c);
            atomicAdd(&sBucket[k[o] >> kBucketBits], c);
        }
    }
    __syncthreads();
    // [bucket][tile], zeroed by the build's memset: a tile of coherent input touches a handful of buckets
    for (uint32_t i = threadIdx.x; i < numBuckets; i += blockDim.x)
        if (sBucket[i]) tileBucket[(size_t) i * tileStride + tile] = sBucket[i];
}

// bucketSizes = the high-byte digit histogram planKernel leaves (hist[256 ..]).  Turns tileBucket[t][b] (entries of
// tile t in bucket b) into the position of the tile's first entry of that bucket in the bucketed layout.
__global__ void __launch_bounds__(256) bucketPrefixKernel(uint32_t *__restrict__ tileBucket, uint32_t tiles, uint32_t tileStride,
                                                          const uint32_t *__restrict__ bucketSizes)
{
    __shared__ uint32_t sScan[33];
    const uint32_t b = blockIdx.x, tid = threadIdx.x;
    uint32_t bucketStart;
    blockExclusiveScan(tid < b ? bucketSizes[tid] : 0u, sScan, &bucketStart);   // numBuckets <= 160 < 256 threads
    if (bucketSizes[b] == 0) return;
    // the bucket's row, 256 consecutive tiles at a time (coalesced), carrying the running total
    uint32_t *const row = tileBucket + (size_t) b * tileStride;
    uint32_t carry = bucketStart;
    for (uint32_t t0 = 0; t0 < tiles; t0 += blockDim.x)
    {
        const uint32_t t = t0 + tid;
        const uint32_t c = t < tiles ? row[t] : 0u;
        uint32_t total;
        const uint32_t excl = blockExclusiveScan(c, sScan, &total);
        if (c) row[t] = carry + excl;
        carry += total;
    }
}

constexpr int kTileEntries = 8 * 256;   // at most 8 entries per splat
__global__ void __launch_bounds__(256) writeEntriesBucketedKernel(uint32_t *__restrict__ keysOut, uint32_t *__restrict__ idsOut,
                                                                  mlsb_splat *__restrict__ splats, float *__restrict__ thresholds,
                                                                  int3 bias, LevelInfo lv, uint32_t firstSplat, uint32_t numSplats,
                                                                  uint32_t numBuckets, uint32_t tileStride,
                                                                  const uint32_t *__restrict__ tileBucketBase)
{
    __shared__ uint32_t sScan[33];
    __shared__ uint32_t sKey[kTileEntries], sId[kTileEntries];
    __shared__ uint32_t sCnt[8][kMaxBuckets];   // per warp and bucket: entries so far, then the base of the warp's entries
    const uint32_t tile = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t i = tid; i < 8 * kMaxBuckets; i += blockDim.x) (&sCnt[0][0])[i] = 0;
    const uint32_t g = tile * blockDim.x + tid, gid = g + firstSplat;
    uint32_t k[8];
    uint32_t cnt = 0;
#pragma unroll
    for (int o = 0; o < 8; o++) k[o] = 0xFFFFFFFFu;
    if (g < numSplats)
    {
        const float4 pr = *reinterpret_cast<const float4 *>(&splats[gid]);
        cnt = splatEntries<true>(pr, bias, lv, k);
        const float invR2 = __fdiv_rn(1.0f, __fmul_rn(pr.w, pr.w));
        splats[gid].radius = invR2;   // kernels/octree.cl:193
        thresholds[gid] = hitThreshold(invR2);
    }
    // the tile's entries in (splat, slot) order, in shared memory
    uint32_t n;
    uint32_t p = blockExclusiveScan(cnt, sScan, &n);
#pragma unroll
    for (int o = 0; o < 8; o++)
        if (k[o] != 0xFFFFFFFFu)
        {
            sKey[p] = k[o];
            sId[p] = gid;
            p++;
        }
    __syncthreads();
    // stable rank inside the tile, per bucket: a warp takes a contiguous span, lanes stride through it
    const uint32_t span = ((n + 7) / 8 + 31) / 32 * 32;   // <= 256
    const uint32_t spanEnd = min((warp + 1) * span, n);
    uint32_t key[8], rank[8];
#pragma unroll
    for (int j = 0; j < 8; j++)
    {
        const uint32_t idx = warp * span + j * 32 + lane;
        const bool valid = (uint32_t) j * 32 < span && idx < spanEnd;
        key[j] = valid ? sKey[idx] : 0xFFFFFFFFu;
        const unsigned vmask = __ballot_sync(0xFFFFFFFFu, valid);
        if (valid)
        {
            const uint32_t b = key[j] >> kBucketBits;
            const unsigned peers = __match_any_sync(vmask, b);
            const int leader = __ffs(peers) - 1;
            uint32_t base = 0;
            if ((int) lane == leader)
            {
                base = sCnt[warp][b];
                sCnt[warp][b] = base + __popc(peers);
            }
            base = __shfl_sync(peers, base, leader);
            rank[j] = base + __popc(peers & ((1u << lane) - 1));
        }
        __syncwarp();
    }
    __syncthreads();
    for (uint32_t b = tid; b < numBuckets; b += blockDim.x)
    {
        uint32_t any = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) any |= sCnt[w][b];
        if (!any) continue;   // the tile has nothing in this bucket: its base was never computed
        uint32_t run = tileBucketBase[(size_t) b * tileStride + tile];
#pragma unroll
        for (int w = 0; w < 8; w++)
        {
            const uint32_t t = sCnt[w][b];
            sCnt[w][b] = run;
            run += t;
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; j++)
        if (key[j] != 0xFFFFFFFFu)
        {
            const uint32_t idx = warp * span + j * 32 + lane;
            const uint32_t dst = sCnt[warp][key[j] >> kBucketBits] + rank[j];
            keysOut[dst] = key[j];
            idsOut[dst] = sId[idx];
        }
}

constexpr int kBucketThreads = 1024, kBucketWarps = kBucketThreads / 32;
constexpr int kBucketChunk = 8 * kBucketThreads;
#ifdef MLSB_BUCKET_MINBLOCKS
__global__ void __launch_bounds__(kBucketThreads, MLSB_BUCKET_MINBLOCKS) bucketSortEmitKernel(
#else
__global__ void __launch_bounds__(kBucketThreads) bucketSortEmitKernel(
#endif
    const uint32_t *__restrict__ keys, const uint32_t *__restrict__ ids,
                                                            const uint32_t *__restrict__ bucketSizes, uint32_t numStart,
                                                            const int32_t *__restrict__ start, int32_t *__restrict__ commands)
{
    __shared__ uint32_t sScan[33];
    __shared__ uint32_t sCnt[kBucketWarps][256];   // per warp and low byte: entries of this chunk so far, then the warp's base
    __shared__ uint32_t sRun[256];      // entries of each low byte in the chunks already done
    __shared__ int32_t sStart[256];     // start[] of the bucket's keys: the slot in front of the key's first id
    const uint32_t b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t bucketStart;
    blockExclusiveScan(tid < b ? bucketSizes[tid] : 0u, sScan, &bucketStart);
    const uint32_t size = bucketSizes[b];
    if (size == 0) return;
    if (tid < 256)
    {
        sRun[tid] = 0;
        const uint32_t key = (b << kBucketBits) | tid;
        sStart[tid] = key < numStart ? start[key] : -1;
    }
    for (uint32_t chunk = 0; chunk < size; chunk += kBucketChunk)
    {
        for (uint32_t i = tid; i < kBucketWarps * 256; i += blockDim.x) (&sCnt[0][0])[i] = 0;
        __syncthreads();
        const uint32_t n = min((uint32_t) kBucketChunk, size - chunk);
        const uint32_t *const ck = keys + bucketStart + chunk, *const ci = ids + bucketStart + chunk;
        uint32_t key[8], id[8], rank[8];
#pragma unroll
        for (int j = 0; j < 8; j++)
        {
            const uint32_t idx = warp * 256 + j * 32 + lane;
            key[j] = idx < n ? ck[idx] : 0xFFFFFFFFu;
            id[j] = idx < n ? ci[idx] : 0u;
        }
#pragma unroll
        for (int j = 0; j < 8; j++)
        {
            const bool valid = key[j] != 0xFFFFFFFFu;
            const unsigned vmask = __ballot_sync(0xFFFFFFFFu, valid);
            if (valid)
            {
                const uint32_t d = key[j] & 255u;
                const unsigned peers = __match_any_sync(vmask, d);
                const int leader = __ffs(peers) - 1;
                uint32_t base = 0;
                if ((int) lane == leader)
                {
                    base = sCnt[warp][d];
                    sCnt[warp][d] = base + __popc(peers);
                }
                base = __shfl_sync(peers, base, leader);
                rank[j] = base + __popc(peers & ((1u << lane) - 1));
            }
            __syncwarp();
        }
        __syncthreads();
        if (tid < 256)
        {
            uint32_t run = sRun[tid];
#pragma unroll 8
            for (int w = 0; w < kBucketWarps; w++)
            {
                const uint32_t t = sCnt[w][tid];
                sCnt[w][tid] = run;
                run += t;
            }
            sRun[tid] = run;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 8; j++)
            if (key[j] != 0xFFFFFFFFu)
            {
                const uint32_t d = key[j] & 255u;
                commands[sStart[d] + 1 + (int32_t) (sCnt[warp][d] + rank[j])] = (int32_t) id[j];   // writeSplatIds, kernels/octree.cl:256-279
            }
        __syncthreads();
    }
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

// What octreeAlloc allocates, for mlsb_worker_resource_usage (SplatTreeCL::resourceUsage, src/splat_tree_cl.cpp:68-92).
uint64_t octreeDeviceBytes(const mlsb_worker_config &c, uint64_t *maxAllocation)
{
    const uint64_t maxStart = (1ull << (3 * c.maxLevels)) / 7;
    const uint64_t maxEntries = 8ull * c.maxSplats;
    const uint64_t maxRanges = maxStart < maxEntries ? maxStart : maxEntries;
    const bool planned = maxStart <= kPlanMaxKeys;
    const uint64_t entryArray = (maxEntries + 8) * sizeof(uint32_t);
    const uint64_t commands = (maxEntries + 2 * maxRanges) * sizeof(int32_t);
    const uint64_t statusBytes = roundUp((uint64_t) (divUp(c.maxSplats, 256u) + 1) * sizeof(uint64_t), (uint64_t) 256);
    const uint64_t countBytes = planned ? roundUp(maxStart * sizeof(uint32_t), (uint64_t) 256) : 0;
    const uint64_t scratch = roundUp(countBytes + statusBytes + 256, (uint64_t) 256) + sortScratchBytes(maxEntries, 4);
    const uint64_t tileBucket = planned ? (512 + (uint64_t) roundUp(divUp(c.maxSplats, 256u) + 1, 32u) * kMaxBuckets) * sizeof(uint32_t) : 0;
    uint64_t total = maxStart * 4 * (planned ? 2 : 6) + commands + (uint64_t) c.maxSplats * sizeof(float) + 4 * entryArray + scratch + tileBucket;
    if (maxAllocation) *maxAllocation = std::max(std::max(commands, entryArray), scratch);
    return total;
}

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
    const bool planned = maxStart <= kPlanMaxKeys;
    MLSB_CUDA(cudaMalloc(&w->start, maxStart * sizeof(int32_t)));
    MLSB_CUDA(cudaMalloc(&w->runIndex, maxStart * sizeof(uint32_t)));
    if (!planned)
    {
        MLSB_CUDA(cudaMalloc(&w->runStart, maxStart * sizeof(int32_t)));
        MLSB_CUDA(cudaMalloc(&w->jumpPos, maxStart * sizeof(int32_t)));
        MLSB_CUDA(cudaMalloc(&w->runFirst, maxStart * sizeof(int32_t)));
        MLSB_CUDA(cudaMalloc(&w->runLast, maxStart * sizeof(int32_t)));
    }
    MLSB_CUDA(cudaMalloc(&w->commands, (size_t) w->numCommandsMax * sizeof(int32_t)));
    MLSB_CUDA(cudaMalloc(&w->hitThresholds, (size_t) c.maxSplats * sizeof(float)));
    for (int i = 0; i < 2; i++)
    {
        MLSB_CUDA(cudaMalloc(&w->entryKeys[i], (maxEntries + 8) * sizeof(uint32_t)));
        MLSB_CUDA(cudaMalloc(&w->entryValues[i], (maxEntries + 8) * sizeof(uint32_t)));
    }
    // [key counts | entry tile status | {ticket, numValid}] sit in front of the sort's scratch, in its allocation;
    // everything from the entry status on is zeroed by one memset per build, the key counts once, here
    w->entryTilesMax = divUp(c.maxSplats, 256u) + 1;
    const size_t statusBytes = roundUp((size_t) w->entryTilesMax * sizeof(uint64_t), (size_t) 256);
    const size_t countBytes = planned ? roundUp((size_t) maxStart * sizeof(uint32_t), (size_t) 256) : 0;
    void *prefix = nullptr;
    MLSB_TRY(sortScratchAlloc(w->sort, maxEntries, 4, countBytes + statusBytes + 256, &prefix));
    w->keyCount = planned ? static_cast<uint32_t *>(prefix) : nullptr;
    w->entryStatus = reinterpret_cast<uint64_t *>(static_cast<char *>(prefix) + countBytes);
    w->entryCounters = reinterpret_cast<uint32_t *>(static_cast<char *>(prefix) + countBytes + statusBytes);
    if (planned)
    {
        // [digit histograms of the bucket path: 512 words | tileBucket rows]: zeroed by one memset per build
        MLSB_CUDA(cudaMalloc(&w->tileBucket, (512 + (size_t) roundUp(w->entryTilesMax, 32u) * kMaxBuckets) * sizeof(uint32_t)));
        MLSB_CUDA(cudaMemset(w->keyCount, 0, countBytes));
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
    cudaFree(w->hitThresholds);
    cudaFree(w->tileBucket);
    for (int i = 0; i < 2; i++)
    {
        cudaFree(w->entryKeys[i]);
        cudaFree(w->entryValues[i]);
    }
    sortScratchFree(w->sort);
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
    const uint32_t bits = 3 * (maxShift - minShift) + 1;
    const uint32_t passes = divUp(bits, 8u);
    const bool planned = w->keyCount != nullptr && numStart <= kPlanMaxKeys && passes <= 2;
    const uint32_t tiles = divUp(numSplats, 256u);
    MLSB_REQUIRE(tiles <= w->entryTilesMax, MLSB_ERR_LENGTH, "entry scratch too small");
    const int3 bias = make_int3(offset[0], offset[1], offset[2]);
    uint32_t *numValid = w->entryCounters + 1;   // device-side count of accepted entries

    static const bool bucketSort = [] { const char *e = getenv("MLSB_OCTREE_SORT"); return e == nullptr || strcmp(e, "onesweep") != 0; }();
    if (planned && bucketSort && bits > kBucketBits && numSplats > 0)
    {
        const uint32_t numBuckets = divUp(numStart, 1u << kBucketBits);
        const uint32_t tileStride = roundUp(tiles, 32u);
        uint32_t *const bucketHist = w->tileBucket, *const tileBucket = w->tileBucket + 512;
        {
            // one memset: the digit histograms planKernel adds to, and the (sparse) per-tile bucket counts
            ScopedStat st(w, "kernel.octree.fill.time", 0);
            MLSB_CUDA(cudaMemsetAsync(bucketHist, 0, (512 + (size_t) numBuckets * tileStride) * sizeof(uint32_t), s));
        }
        {
            ScopedStat st(w, "kernel.octree.writeEntries.time");
            countEntriesKernel<<<tiles, 256, 0, s>>>(w->splats, bias, lv, firstSplat, numSplats, numBuckets, tileStride, w->keyCount,
                                                     tileBucket);
        }
        {
            ScopedStat st(w, "kernel.octree.scan.time");
            planKernel<<<kPlanCluster, kPlanThreads, 0, s>>>(w->keyCount, numStart, lv, bits, w->runIndex, w->start, w->commands,
                                                              bucketHist, numValid);
            bucketPrefixKernel<<<numBuckets, 256, 0, s>>>(tileBucket, tiles, tileStride, bucketHist + 256);
        }
        {
            ScopedStat st(w, "kernel.octree.writeEntries.time");
            writeEntriesBucketedKernel<<<tiles, 256, 0, s>>>(w->entryKeys[0], w->entryValues[0], w->splats, w->hitThresholds, bias, lv,
                                                             firstSplat, numSplats, numBuckets, tileStride, tileBucket);
        }
        {
            ScopedStat st(w, "kernel.octree.sort.time");
            bucketSortEmitKernel<<<numBuckets, kBucketThreads, 0, s>>>(w->entryKeys[0], w->entryValues[0], bucketHist + 256, numStart, w->start,
                                                                       w->commands);
        }
        w->launches += 5;
        MLSB_CUDA(cudaGetLastError());
        return MLSB_OK;
    }
    if (planned)
    {
        {
            // one memset: the entry tiles' look-back words and counters and, behind them in the allocation,
            // the sort's histograms, tickets and status.  The key counts are cleaned by planKernel itself.
            ScopedStat st(w, "kernel.octree.fill.time", 0);
            const size_t prefixBytes = reinterpret_cast<char *>(w->sort.hist) - reinterpret_cast<char *>(w->entryStatus);
            const size_t sortBytes = numEntries ? sortScratchBytesUsed(w->sort, numEntries, passes) : 0;
            MLSB_CUDA(cudaMemsetAsync(w->entryStatus, 0, prefixBytes + sortBytes, s));
        }
        if (numSplats > 0)
        {
            ScopedStat st(w, "kernel.octree.writeEntries.time");
            writeEntriesFusedKernel<<<tiles, 256, 0, s>>>(w->entryKeys[0], w->entryValues[0], w->splats, w->hitThresholds, bias, lv,
                                                          firstSplat, numSplats, w->entryStatus, w->entryCounters, w->keyCount);
        }
        {
            ScopedStat st(w, "kernel.octree.scan.time");
            planKernel<<<kPlanCluster, kPlanThreads, 0, s>>>(w->keyCount, numStart, lv, bits, w->runIndex, w->start, w->commands,
                                                              w->sort.hist, numValid);
        }
        if (numSplats > 0)
            MLSB_TRY(radixSortToCommands(w, w->entryKeys, w->entryValues, numEntries, numValid, bits, w->commands, w->runIndex,
                                         "kernel.octree.sort.time"));
        w->launches += (numSplats > 0 ? 1 : 0) + 1;
        MLSB_CUDA(cudaGetLastError());
        return MLSB_OK;
    }

    // ---- general path: any number of levels ----
    {
        ScopedStat st(w, "kernel.octree.writeEntries.time");
        const size_t prefixBytes = reinterpret_cast<char *>(w->entryCounters + 64) - reinterpret_cast<char *>(w->entryStatus);
        MLSB_CUDA(cudaMemsetAsync(w->entryStatus, 0, prefixBytes, s));
        if (numSplats > 0)
            writeEntriesFusedKernel<<<tiles, 256, 0, s>>>(w->entryKeys[0], w->entryValues[0], w->splats, w->hitThresholds, bias, lv,
                                                          firstSplat, numSplats, w->entryStatus, w->entryCounters, nullptr);
    }
    int cur = 0;
    MLSB_TRY(radixSortPairsDeviceCount(w, w->entryKeys, w->entryValues, numEntries, numValid,
                                       bits, &cur, "kernel.octree.sort.time", false));
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
        uint32_t scanTiles = divUp(numStart, (uint32_t) (kScanThreads * kScanItems));
        MLSB_REQUIRE(scanTiles <= w->scanStatusWords, MLSB_ERR_LENGTH, "scan status too small");
        MLSB_CUDA(cudaMemsetAsync(w->scanStatus, 0, (size_t) scanTiles * sizeof(uint64_t), s));
        MLSB_CUDA(cudaMemsetAsync(w->scanTicket, 0, sizeof(uint32_t), s));
        scanRunsKernel<<<scanTiles, kScanThreads, 0, s>>>(w->runFirst, w->runLast, numStart, w->runIndex, w->runStart,
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
    w->launches += (numSplats > 0 ? 3 : 0) + 2;
    MLSB_CUDA(cudaGetLastError());
    return MLSB_OK;
}

}  // namespace mlsb
