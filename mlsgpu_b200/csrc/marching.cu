// Marching tetrahedra + vertex welding + scale/bias: Marching::generate / addSlices / shipOut
// (src/marching.cpp:500-824), kernels/marching.cl and kernels/scale_bias.cl, re-designed around
// a field that is resident for the whole bin (no swathes, no copySlice):
//
//   generateCells   genOccupied + scanElements + generateElements in ONE pass: a tile of cells
//                   is classified, its (vertex, index) counts are block-scanned, the tile's base
//                   comes from a decoupled look-back, and occupied cells emit their vertices,
//                   compact 32-bit keys and local indices straight away (kernels/marching.cl:84-258)
//   radix sort      (compact key, unwelded index) pairs; the 64-bit key of the reference only has
//                   3 x <=10 live coordinate bits + the external flag, so it packs into <= 31 bits
//                   and the 16-byte vertex payload is not moved by the sort at all
//   weldVertices    countUniqueVertices + scan + compactVertices + scaleBiasVertices in one pass
//                   (kernels/marching.cl:271-326, kernels/scale_bias.cl:32-40)
//   reindex         kernels/marching.cl:334-340
//
// Cells are visited in x-fastest linear order (the reference's atomic order is arbitrary), so the
// output is deterministic: vertices sorted by key (internal first), triangles in cell order.
#include "common.cuh"
#include "scan.cuh"
#include <algorithm>
#include <cstring>

namespace mlsb
{

namespace
{

constexpr int kNumEdges = 19;
constexpr int kGenThreads = 256;
constexpr int kWeldThreads = 256;
constexpr int kWeldItems = 4;

// counters in mlsb_worker::marchCounters
enum
{
    CTR_VERTS = 0,      // unwelded vertices
    CTR_INDICES = 1,
    CTR_OVERFLOW = 2,
    CTR_WELDED = 3,
    CTR_FIRST_EXTERNAL = 4,
    CTR_CELLS = 5,
    CTR_GEN_TICKET = 8,
    CTR_WELD_TICKET = 9,
    CTR_COUNT = 16
};

// src/marching.cpp:50-81
const uint8_t hEdgeIndices[kNumEdges][2] = {
    {0, 1}, {0, 2}, {0, 3}, {1, 3}, {2, 3}, {0, 4}, {0, 5}, {1, 5}, {4, 5}, {0, 6},
    {2, 6}, {4, 6}, {0, 7}, {1, 7}, {2, 7}, {3, 7}, {4, 7}, {5, 7}, {6, 7}};
const uint8_t hTetrahedra[6][4] = {{0, 7, 1, 3}, {0, 7, 3, 2}, {0, 7, 2, 6}, {0, 7, 6, 4}, {0, 7, 4, 5}, {0, 7, 5, 1}};

// Device layout of the lookup tables (one allocation, see marchingAlloc)
struct Tables
{
    uint8_t count[256][2];
    uint16_t start[257][2];
    uint8_t data[8192];
    uint8_t edge[kNumEdges][2];   // corner ids of each edge
};

struct HostTables
{
    Tables t;
    uint32_t key[2432][3];
    bool built = false;
};
HostTables gTables;

int edgeByCorners(int v0, int v1)
{
    if (v0 > v1) std::swap(v0, v1);
    for (int i = 0; i < kNumEdges; i++)
        if (hEdgeIndices[i][0] == v0 && hEdgeIndices[i][1] == v1) return i;
    return -1;
}

int inversions(const int v[4])
{
    int p = 0;
    for (int i = 0; i < 4; i++)
        for (int j = i + 1; j < 4; j++) p ^= v[i] > v[j];
    return p;
}

// Marching::makeTables, src/marching.cpp:109-252
void buildTables()
{
    if (gTables.built) return;
    std::vector<uint8_t> vertexTable, indexTable;
    Tables &T = gTables.t;
    for (int code = 0; code < 256; code++)
    {
        T.start[code][0] = (uint16_t) vertexTable.size();
        T.start[code][1] = (uint16_t) indexTable.size();
        std::vector<int> tris;
        for (int j = 0; j < 6; j++)
        {
            int v[4];
            bool outside[8] = {};
            int nOut = 0;
            for (int k = 0; k < 4; k++)
            {
                v[k] = hTetrahedra[j][k];
                outside[v[k]] = (code >> v[k]) & 1;
                nOut += outside[v[k]];
            }
            int parity = inversions(v);
            if (nOut > 2)
            {
                parity ^= 1;
                for (int k = 0; k < 4; k++) outside[v[k]] = !outside[v[k]];
            }
            std::sort(v, v + 4);
            do
            {
                if (inversions(v) != parity) continue;
                int mask = 0;
                for (int k = 0; k < 4; k++) mask |= (int) outside[v[k]] << k;
                if (mask == 0) break;
                if (mask == 1)
                {
                    for (int o : {1, 3, 2}) tris.push_back(edgeByCorners(v[0], v[o]));
                    break;
                }
                if (mask == 3)
                {
                    const int pairs[6][2] = {{0, 2}, {1, 2}, {1, 3}, {1, 3}, {0, 3}, {0, 2}};
                    for (auto &p : pairs) tris.push_back(edgeByCorners(v[p[0]], v[p[1]]));
                    break;
                }
            } while (std::next_permutation(v, v + 4));
        }
        int compact[kNumEdges];
        int pool = 0;
        for (int e = 0; e < kNumEdges; e++)
            if (std::find(tris.begin(), tris.end(), e) != tris.end())
            {
                compact[e] = pool++;
                for (int axis = 0; axis < 3; axis++)
                    gTables.key[vertexTable.size()][axis] = ((hEdgeIndices[e][0] >> axis) & 1) + ((hEdgeIndices[e][1] >> axis) & 1);
                vertexTable.push_back((uint8_t) e);
            }
        for (int e : tris) indexTable.push_back((uint8_t) compact[e]);
        T.count[code][0] = (uint8_t) (vertexTable.size() - T.start[code][0]);
        T.count[code][1] = (uint8_t) (indexTable.size() - T.start[code][1]);
    }
    T.start[256][0] = (uint16_t) vertexTable.size();
    T.start[256][1] = (uint16_t) indexTable.size();
    for (int i = 0; i <= 256; i++) T.start[i][1] = (uint16_t) (T.start[i][1] + vertexTable.size());
    memset(T.data, 0, sizeof(T.data));
    memcpy(T.data, vertexTable.data(), vertexTable.size());
    memcpy(T.data + vertexTable.size(), indexTable.data(), indexTable.size());
    memcpy(T.edge, hEdgeIndices, sizeof(T.edge));
    gTables.built = true;
}

struct GenParams
{
    const float *field;
    uint32_t pitch, zStride;
    uint32_t cellsX, cellsY;       // W-1, H-1
    uint32_t zFirst, zLast;        // cells between slices zFirst..zLast
    uint32_t numCells;
    uint3 gridOffset;              // keyOffset: local -> global grid coordinates
    uint3 top;                     // kernels/marching.cl:151
    uint32_t keyBits;              // bits per axis of the compact key
    uint32_t vertexSpace, indexSpace;
    float4 *unwelded;
    uint32_t *keys, *vidx, *indices;
    uint32_t *counters;
    volatile uint64_t *status;
    const Tables *tables;
};

__device__ __forceinline__ bool finiteF(float x) { return (__float_as_uint(x) & 0x7F800000u) != 0x7F800000u; }

// Tile = kGenRows rows of cells of one z slab (x-fastest linear order is preserved because tiles
// are numbered z-major, then by row group).  The 2 x (rows+1) corner rows the tile needs are
// staged in shared memory with coalesced float4 loads, so every field value is read from HBM
// about 2.25 times per bin instead of 8, and a tile takes ONE ticket / look-back for up to
// 8 x 511 cells.  Occupied cells (a few percent) are then emitted warp-cooperatively: one lane
// per vertex / index, so the stores of a cell are contiguous.
constexpr int kGenRows = 8;
constexpr int kGenMaxVRows = 16;   // rows x ceil(cellsX / 256) virtual rows per tile

__device__ __forceinline__ uint32_t cellCode(const float *s0, const float *s1, uint32_t ps, uint32_t x, bool *valid)
{
    // s0 / s1: corner row y of slices z / z+1 in shared memory; ps = row pitch
    const float iso[8] = {s0[x], s0[x + 1], s0[ps + x], s0[ps + x + 1], s1[x], s1[x + 1], s1[ps + x], s1[ps + x + 1]};
    uint32_t code = 0;
    bool ok = true;
#pragma unroll
    for (int i = 0; i < 8; i++)
    {
        code |= (iso[i] >= 0.0f ? 1u : 0u) << i;
        ok = ok && finiteF(iso[i]);
    }
    *valid = ok;
    return code;
}

__global__ void __launch_bounds__(kGenThreads) generateCellsKernel(GenParams P)
{
    extern __shared__ float sField[];   // [2 slices][kGenRows + 1 rows][ps]
    __shared__ uint64_t sWarpTot[kGenMaxVRows * (kGenThreads / 32)];
    __shared__ uint64_t sScan[33];
    __shared__ uint32_t sTile;
    __shared__ uint64_t sPrefix;

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) sTile = atomicAdd(&P.counters[CTR_GEN_TICKET], 1u);
    __syncthreads();
    const uint32_t tile = sTile;
    const uint32_t rowGroups = (P.cellsY + kGenRows - 1) / kGenRows;
    const uint32_t z = P.zFirst + tile / rowGroups;
    const uint32_t y0 = (tile % rowGroups) * kGenRows;
    const uint32_t nRows = min((uint32_t) kGenRows, P.cellsY - y0);
    const uint32_t ps = (P.cellsX + 1 + 3) & ~3u;          // staged row pitch (floats), <= P.pitch
    const uint32_t chunks = (P.cellsX + kGenThreads - 1) / kGenThreads;
    const uint32_t vRows = nRows * chunks;

    // ---- stage the corner rows (float4, coalesced) ----
    {
        const uint32_t ps4 = ps / 4, rowsPerSlice = nRows + 1;
        const uint32_t total4 = 2 * rowsPerSlice * ps4;
        for (uint32_t i = tid; i < total4; i += kGenThreads)
        {
            uint32_t col4 = i % ps4, r = i / ps4;
            uint32_t slice = r / rowsPerSlice, row = r % rowsPerSlice;
            const float4 *src = reinterpret_cast<const float4 *>(P.field + ((size_t) (z + slice) * P.zStride + y0 + row) * P.pitch) + col4;
            reinterpret_cast<float4 *>(sField + (slice * (kGenRows + 1) + row) * ps)[col4] = __ldg(src);
        }
    }
    __syncthreads();
    const float *sl0 = sField, *sl1 = sField + (kGenRows + 1) * ps;

    // ---- genOccupied + tile-local scan (fields: vertices | indices << 21 | cells << 42) ----
    // inside a warp the three counts fit 10 + 12 + 6 bits of one 32-bit word
    uint32_t excl[kGenMaxVRows];
#pragma unroll
    for (int v = 0; v < kGenMaxVRows; v++)
    {
        excl[v] = 0;
        if ((uint32_t) v < vRows)
        {
            const uint32_t r = v / chunks, x = (v % chunks) * kGenThreads + tid;
            uint32_t packed = 0;
            if (x < P.cellsX)
            {
                bool valid;
                uint32_t code = cellCode(sl0 + r * ps, sl1 + r * ps, ps, x, &valid);
                if (valid && code != 0 && code != 255)
                    packed = (uint32_t) P.tables->count[code][0] | ((uint32_t) P.tables->count[code][1] << 10) | (1u << 22);
            }
            uint32_t inc = warpInclusiveScan(packed);
            excl[v] = inc - packed;
            if (lane == 31)
                sWarpTot[v * (kGenThreads / 32) + warp] =
                    (uint64_t) (inc & 0x3FF) | ((uint64_t) ((inc >> 10) & 0xFFF) << 21) | ((uint64_t) (inc >> 22) << 42);
        }
    }
    __syncthreads();
    {
        uint64_t t = tid < vRows * (kGenThreads / 32) ? sWarpTot[tid] : 0ull;
        uint64_t total;
        uint64_t e = blockExclusiveScan(t, sScan, &total);
        if (tid < vRows * (kGenThreads / 32)) sWarpTot[tid] = e;
        if (tid < 32)
        {
            const uint64_t agg = (total & 0x1FFFFF) | (((total >> 21) & 0x1FFFFF) << 31);
            uint64_t p = lookbackExclusive(P.status, tile, agg);
            if (tid == 0)
            {
                sPrefix = p;
                if (tile == gridDim.x - 1)
                {
                    P.counters[CTR_VERTS] = (uint32_t) ((p + agg) & 0x7FFFFFFFu);
                    P.counters[CTR_INDICES] = (uint32_t) ((p + agg) >> 31);
                }
                if (total >> 42) atomicAdd(&P.counters[CTR_CELLS], (uint32_t) (total >> 42));
            }
        }
    }
    __syncthreads();
    const uint32_t vBase = (uint32_t) (sPrefix & 0x7FFFFFFFu), iBase = (uint32_t) (sPrefix >> 31);

    // ---- generateElements for the occupied cells, one warp per cell ----
    const Tables &T = *P.tables;
#pragma unroll
    for (int v = 0; v < kGenMaxVRows; v++)
    {
        if ((uint32_t) v >= vRows) break;
        const uint32_t r = v / chunks, x = (v % chunks) * kGenThreads + tid;
        const float *s0 = sl0 + r * ps, *s1 = sl1 + r * ps;
        uint32_t code = 0;
        bool occupied = false;
        if (x < P.cellsX)
        {
            bool valid;
            code = cellCode(s0, s1, ps, x, &valid);
            occupied = valid && code != 0 && code != 255;
        }
        unsigned mask = __ballot_sync(0xFFFFFFFFu, occupied);
        if (mask == 0) continue;
        const uint64_t wbase = sWarpTot[v * (kGenThreads / 32) + warp];
        const uint32_t myV = vBase + (uint32_t) (wbase & 0x1FFFFF) + (excl[v] & 0x3FF);
        const uint32_t myI = iBase + (uint32_t) ((wbase >> 21) & 0x1FFFFF) + ((excl[v] >> 10) & 0xFFF);
        while (mask)
        {
            const int src = __ffs(mask) - 1;
            mask &= mask - 1;
            const uint32_t cx = __shfl_sync(0xFFFFFFFFu, x, src);
            const uint32_t cc = __shfl_sync(0xFFFFFFFFu, code, src);
            const uint32_t vNext = __shfl_sync(0xFFFFFFFFu, myV, src);
            const uint32_t iNext = __shfl_sync(0xFFFFFFFFu, myI, src);
            const uint32_t nv = T.count[cc][0], ni = T.count[cc][1];
            if (vNext + nv > P.vertexSpace || iNext + ni > P.indexSpace)
            {
                if (lane == 0) P.counters[CTR_OVERFLOW] = 1;
                continue;
            }
            const uint32_t sx = T.start[cc][0], sy = T.start[cc][1];
            const uint32_t y = y0 + r;
            if (lane < nv)
            {
                const uint32_t e = T.data[sx + lane];
                const uint32_t a = T.edge[e][0], b = T.edge[e][1];
                const uint32_t oax = a & 1, oay = (a >> 1) & 1, oaz = a >> 2;
                const uint32_t obx = b & 1, oby = (b >> 1) & 1, obz = b >> 2;
                const uint32_t dx = obx - oax, dy = oby - oay, dz = obz - oaz;
                const float isoA = (oaz ? s1 : s0)[oay * ps + cx + oax];
                const float isoB = (obz ? s1 : s0)[oby * ps + cx + obx];
                // interp(), kernels/marching.cl:130-138: no contraction, explicit fma
                const float inv = __fdiv_rn(1.0f, __fsub_rn(isoA, isoB));
                const float t = __fmul_rn(isoA, inv);
                float4 pos;
                pos.x = __fmaf_rn(t, (float) dx, (float) (cx + P.gridOffset.x + oax));
                pos.y = __fmaf_rn(t, (float) dy, (float) (y + P.gridOffset.y + oay));
                pos.z = __fmaf_rn(t, (float) dz, (float) (z + P.gridOffset.z + oaz));
                pos.w = 0.0f;
                // computeKey(), :148-154, on the compact layout (flag | z | y | x)
                const uint32_t kx = 2 * cx + 2 * oax + dx, ky = 2 * y + 2 * oay + dy, kz = 2 * z + 2 * oaz + dz;
                uint32_t key = (kz << (2 * P.keyBits)) | (ky << P.keyBits) | kx;
                if (kx == 0 || ky == 0 || kx == P.top.x || ky == P.top.y || kz == P.top.z) key |= 1u << (3 * P.keyBits);
                P.unwelded[vNext + lane] = pos;
                P.keys[vNext + lane] = key;
                P.vidx[vNext + lane] = vNext + lane;
            }
            for (uint32_t i = lane; i < ni; i += 32) P.indices[iNext + i] = vNext + T.data[sy + i];
        }
    }
}

struct WeldParams
{
    const uint32_t *keys, *vidx;
    const float4 *unwelded;
    float *welded;
    uint64_t *weldedKeys;
    uint32_t *indexRemap;
    uint32_t *counters;
    volatile uint64_t *status;
    uint32_t keyBits;
    uint32_t zMax2;            // 2 * zMax: z coordinates >= this are external (src/marching.cpp:593)
    uint64_t keyOffset;        // src/marching.cpp:594-599
    float scale, bx, by, bz;   // ScaleBiasFilter
};

__global__ void __launch_bounds__(kWeldThreads) weldVerticesKernel(WeldParams P)
{
    __shared__ uint64_t sScan[33];
    __shared__ uint32_t sTile;
    __shared__ uint64_t sPrefix;
    const uint32_t n = P.counters[CTR_VERTS];
    if (threadIdx.x == 0) sTile = atomicAdd(&P.counters[CTR_WELD_TICKET], 1u);
    __syncthreads();
    const uint32_t tile = sTile;
    const uint32_t tileBase = tile * (kWeldThreads * kWeldItems);
    if (tileBase >= n) return;   // ticket order: every later tile is out of range too
    const uint32_t base = tileBase + threadIdx.x * kWeldItems;

    uint32_t key[kWeldItems + 1];
#pragma unroll
    for (int i = 0; i <= kWeldItems; i++) key[i] = base + i < n ? P.keys[base + i] : 0xFFFFFFFFu;
    const uint32_t coordMask = (1u << P.keyBits) - 1;
    const uint32_t flagBit = 1u << (3 * P.keyBits);
    uint32_t nUnique = 0, nInternal = 0;
    bool last[kWeldItems], ext[kWeldItems];
#pragma unroll
    for (int i = 0; i < kWeldItems; i++)
    {
        bool in = base + i < n;
        last[i] = in && key[i] != key[i + 1];   // countUniqueVertices, :271-279 (sentinel = UINT_MAX)
        uint32_t z2 = (key[i] >> (2 * P.keyBits)) & coordMask;
        ext[i] = (key[i] & flagBit) || z2 >= P.zMax2;
        nUnique += last[i];
        nInternal += last[i] && !ext[i];
    }
    uint64_t packed = (uint64_t) nUnique | ((uint64_t) nInternal << 31);
    uint64_t total;
    uint64_t excl = blockExclusiveScan(packed, sScan, &total);
    if (threadIdx.x < 32)
    {
        uint64_t p = lookbackExclusive(P.status, tile, total);
        if (threadIdx.x == 0) sPrefix = p;
    }
    __syncthreads();
    uint32_t u = (uint32_t) ((sPrefix + excl) & 0x7FFFFFFFu);
    if (tileBase + kWeldThreads * kWeldItems >= n && threadIdx.x == kWeldThreads - 1)
    {
        uint64_t end = sPrefix + total;
        P.counters[CTR_WELDED] = (uint32_t) (end & 0x7FFFFFFFu);
        P.counters[CTR_FIRST_EXTERNAL] = (uint32_t) (end >> 31);
    }

    // ---- compactVertices (:295-326) + scaleBiasVertices ----
#pragma unroll
    for (int i = 0; i < kWeldItems; i++)
    {
        if (base + i >= n) break;
        const uint32_t orig = P.vidx[base + i];
        P.indexRemap[orig] = u;
        if (last[i])
        {
            const float4 v = P.unwelded[orig];
            P.welded[3 * (size_t) u + 0] = __fmaf_rn(v.x, P.scale, P.bx);
            P.welded[3 * (size_t) u + 1] = __fmaf_rn(v.y, P.scale, P.by);
            P.welded[3 * (size_t) u + 2] = __fmaf_rn(v.z, P.scale, P.bz);
            if (ext[i])
            {
                const uint64_t kx = key[i] & coordMask, ky = (key[i] >> P.keyBits) & coordMask,
                               kz = (key[i] >> (2 * P.keyBits)) & coordMask;
                P.weldedKeys[u] = (kx | (ky << kKeyAxisBits) | (kz << (2 * kKeyAxisBits))) + P.keyOffset;
            }
            u++;
        }
    }
}

__global__ void __launch_bounds__(256) reindexKernel(uint32_t *__restrict__ indices, const uint32_t *__restrict__ indexRemap,
                                                     const uint32_t *__restrict__ counters)
{
    const uint32_t n = counters[CTR_INDICES];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        indices[i] = indexRemap[indices[i]];
}

}  // namespace

int marchingTablesHost(uint8_t *count, uint16_t *start, uint8_t *data, uint32_t *key)
{
    buildTables();
    if (count) memcpy(count, gTables.t.count, sizeof(gTables.t.count));
    if (start) memcpy(start, gTables.t.start, sizeof(gTables.t.start));
    if (data) memcpy(data, gTables.t.data, sizeof(gTables.t.data));
    if (key) memcpy(key, gTables.key, sizeof(gTables.key));
    return MLSB_OK;
}

int marchingAlloc(mlsb_worker *w)
{
    const mlsb_worker_config &c = w->cfg;
    const uint32_t corners = c.maxCells + 1;
    // src/marching.cpp:311-315 (the reference's MAX_DIMENSION is 8192; the compact sort key
    // used here carries 10 bits per axis)
    MLSB_REQUIRE(corners >= 2 && corners <= 512, MLSB_ERR_INVALID_ARGUMENT, "maxCells + 1 must be in [2, 512]");
    MLSB_REQUIRE(c.meshBytes >= (uint64_t) c.maxCells * c.maxCells * kMaxCellBytes, MLSB_ERR_INVALID_ARGUMENT,
                 "meshBytes smaller than one slice of worst-case cells");
    buildTables();
    w->fieldDim = roundUp(corners, (uint32_t) kMlsWgs);
    const uint64_t meshCells = c.meshBytes / kMaxCellBytes;
    w->vertexSpace = meshCells * kMaxCellVertices;
    w->indexSpace = meshCells * kMaxCellIndices;
    MLSB_REQUIRE(w->vertexSpace < (1ull << 31) && w->indexSpace < (1ull << 31), MLSB_ERR_LENGTH, "meshBytes too large");
    MLSB_CUDA(cudaMalloc(&w->field, (size_t) w->fieldDim * w->fieldDim * w->fieldDim * sizeof(float)));
    MLSB_CUDA(cudaMalloc(&w->unwelded, (w->vertexSpace + 1) * sizeof(float4)));
    for (int i = 0; i < 2; i++)
    {
        MLSB_CUDA(cudaMalloc(&w->vertKeys[i], (w->vertexSpace + 8) * sizeof(uint32_t)));
        MLSB_CUDA(cudaMalloc(&w->vertIdx[i], (w->vertexSpace + 8) * sizeof(uint32_t)));
    }
    MLSB_CUDA(cudaMalloc(&w->indexRemap, (w->vertexSpace + 1) * sizeof(uint32_t)));
    MLSB_CUDA(cudaMalloc(&w->indices, (w->indexSpace + 1) * sizeof(uint32_t)));
    MLSB_CUDA(cudaMalloc(&w->welded, (w->vertexSpace + 1) * 3 * sizeof(float)));
    MLSB_CUDA(cudaMalloc(&w->weldedKeys, (w->vertexSpace + 1) * sizeof(uint64_t)));
    MLSB_CUDA(cudaMalloc(&w->marchCounters, CTR_COUNT * sizeof(uint32_t)));
    MLSB_CUDA(cudaMalloc(&w->tables, sizeof(Tables)));
    MLSB_CUDA(cudaMemcpy(w->tables, &gTables.t, sizeof(Tables), cudaMemcpyHostToDevice));
    mlsb_mesh_sizes worst{w->vertexSpace, w->indexSpace / 3, 0};
    w->hostBlobBytes = mlsb_mesh_host_bytes(&worst) + 64;
    MLSB_CUDA(cudaMallocHost(&w->hostBlob, w->hostBlobBytes));
    MLSB_CUDA(cudaMallocHost(&w->hostCounters, CTR_COUNT * sizeof(uint32_t)));
    return MLSB_OK;
}

void marchingFree(mlsb_worker *w)
{
    cudaFree(w->field);
    cudaFree(w->unwelded);
    for (int i = 0; i < 2; i++)
    {
        cudaFree(w->vertKeys[i]);
        cudaFree(w->vertIdx[i]);
    }
    cudaFree(w->indexRemap);
    cudaFree(w->indices);
    cudaFree(w->welded);
    cudaFree(w->weldedKeys);
    cudaFree(w->marchCounters);
    cudaFree(w->tables);
    cudaFreeHost(w->hostBlob);
    cudaFreeHost(w->hostCounters);
}

namespace
{

uint32_t bitLength(uint32_t v)
{
    uint32_t b = 0;
    while (v) { b++; v >>= 1; }
    return b;
}

// One shipped-out block: the cells between corner slices zFirst..zLast (Marching::addSlices +
// shipOut with zTop = zFirst, zMax = zLast).  Returns MLSB_ERR_OVERFLOW if the block does not
// fit in the worker's mesh memory (the caller then splits the z range, as
// src/marching.cpp:651-704 does with its per-slice histogram).
int shipBlock(mlsb_worker *w, uint32_t pitch, uint32_t zStride, const uint32_t size[3], uint32_t zFirst, uint32_t zLast,
              const uint32_t keyOffset[3], mlsb_mesh_sink sink, void *user)
{
    cudaStream_t s = w->stream;
    const uint32_t cellsX = size[0] - 1, cellsY = size[1] - 1;
    const uint64_t numCells64 = (uint64_t) cellsX * cellsY * (zLast - zFirst);
    if (numCells64 == 0) return MLSB_OK;
    MLSB_REQUIRE(numCells64 < (1ull << 32) - kGenThreads, MLSB_ERR_LENGTH, "too many cells");
    const uint32_t numCells = (uint32_t) numCells64;
    const uint32_t tiles = (zLast - zFirst) * divUp(cellsY, (uint32_t) kGenRows);
    const uint32_t stagedPitch = (cellsX + 1 + 3) & ~3u;
    const size_t genSmem = 2 * (kGenRows + 1) * (size_t) stagedPitch * sizeof(float);
    MLSB_REQUIRE(divUp(cellsX, (uint32_t) kGenThreads) * kGenRows <= (uint32_t) kGenMaxVRows, MLSB_ERR_LENGTH, "bin too wide");
    const uint32_t weldTilesMax = (uint32_t) divUp(w->vertexSpace, (uint64_t) (kWeldThreads * kWeldItems)) + 1;
    MLSB_REQUIRE(std::max(tiles, weldTilesMax) <= w->scanStatusWords, MLSB_ERR_LENGTH, "scan status too small");

    GenParams G;
    G.field = w->field;
    G.pitch = pitch;
    G.zStride = zStride;
    G.cellsX = cellsX;
    G.cellsY = cellsY;
    G.zFirst = zFirst;
    G.zLast = zLast;
    G.numCells = numCells;
    G.gridOffset = make_uint3(keyOffset[0], keyOffset[1], keyOffset[2]);
    G.top = make_uint3(2 * cellsX, 2 * cellsY, 2 * zFirst);
    G.keyBits = bitLength(std::max(std::max(2 * cellsX, 2 * cellsY), 2 * zLast));
    G.vertexSpace = (uint32_t) w->vertexSpace;
    G.indexSpace = (uint32_t) w->indexSpace;
    G.unwelded = w->unwelded;
    G.keys = w->vertKeys[0];
    G.vidx = w->vertIdx[0];
    G.indices = w->indices;
    G.counters = w->marchCounters;
    G.status = w->scanStatus;
    G.tables = reinterpret_cast<const Tables *>(w->tables);
    MLSB_REQUIRE(3 * G.keyBits + 1 <= 32, MLSB_ERR_LENGTH, "bin too large for the compact vertex key");

    {
        ScopedStat st(w, "kernel.marching.zero.time", 0);
        MLSB_CUDA(cudaMemsetAsync(w->marchCounters, 0, CTR_COUNT * sizeof(uint32_t), s));
        MLSB_CUDA(cudaMemsetAsync(w->scanStatus, 0, (size_t) tiles * sizeof(uint64_t), s));
    }
    {
        ScopedStat st(w, "kernel.marching.generateElements.time");
        generateCellsKernel<<<tiles, kGenThreads, genSmem, s>>>(G);
        w->launches++;
    }
    MLSB_CUDA(cudaGetLastError());

    // Sizes are only known on the device; the sort runs on the device-side count.
    int cur = 0;
    MLSB_TRY(radixSortPairsDeviceCount(w, w->vertKeys, w->vertIdx, (uint32_t) w->vertexSpace,
                                       w->marchCounters + CTR_VERTS, 3 * G.keyBits + 1, &cur,
                                       "kernel.marching.sortVertices.time"));

    WeldParams W;
    W.keys = w->vertKeys[cur];
    W.vidx = w->vertIdx[cur];
    W.unwelded = w->unwelded;
    W.welded = w->welded;
    W.weldedKeys = w->weldedKeys;
    W.indexRemap = w->indexRemap;
    W.counters = w->marchCounters;
    W.status = w->scanStatus;
    W.keyBits = G.keyBits;
    W.zMax2 = 2 * zLast;
    W.keyOffset = ((uint64_t) keyOffset[2] << (2 * kKeyAxisBits + 1)) | ((uint64_t) keyOffset[1] << (kKeyAxisBits + 1))
                  | ((uint64_t) keyOffset[0] << 1);
    W.scale = w->scale;
    W.bx = w->bias[0];
    W.by = w->bias[1];
    W.bz = w->bias[2];
    {
        ScopedStat st(w, "kernel.marching.compactVertices.time");
        MLSB_CUDA(cudaMemsetAsync(w->scanStatus, 0, (size_t) weldTilesMax * sizeof(uint64_t), s));
        weldVerticesKernel<<<weldTilesMax, kWeldThreads, 0, s>>>(W);
        w->launches++;
    }
    {
        ScopedStat st(w, "kernel.marching.reindex.time");
        reindexKernel<<<148 * 8, 256, 0, s>>>(w->indices, w->indexRemap, w->marchCounters);
        w->launches++;
    }
    MLSB_CUDA(cudaGetLastError());

    // ---- readback of the sizes (src/marching.cpp:587-617), then the mesh itself ----
    {
        ScopedStat st(w, "kernel.marching.readback.time", 0);
        MLSB_CUDA(cudaMemcpyAsync(w->hostCounters, w->marchCounters, CTR_COUNT * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    }
    MLSB_CUDA(cudaStreamSynchronize(s));
    const uint32_t *hc = w->hostCounters;
    if (hc[CTR_OVERFLOW] || hc[CTR_VERTS] > w->vertexSpace || hc[CTR_INDICES] > w->indexSpace) return MLSB_ERR_OVERFLOW;
    if (hc[CTR_VERTS] == 0) return MLSB_OK;

    mlsb_mesh_sizes sizes{hc[CTR_WELDED], hc[CTR_INDICES] / 3, hc[CTR_FIRST_EXTERNAL]};
    const uint64_t numExternal = sizes.numVertices - sizes.numInternalVertices;
    // HostKeyMesh layout, src/mesh.cpp:51-60
    uint8_t *blob = static_cast<uint8_t *>(w->hostBlob);
    uint8_t *hKeys = blob;
    uint8_t *hVerts = hKeys + numExternal * sizeof(uint64_t);
    uint8_t *hTris = hVerts + sizes.numVertices * 3 * sizeof(float);
    // enqueueReadMesh, src/mesh.cpp:62-102
    MLSB_CUDA(cudaMemcpyAsync(hTris, w->indices, sizes.numTriangles * 3 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    if (numExternal)
        MLSB_CUDA(cudaMemcpyAsync(hKeys, w->weldedKeys + sizes.numInternalVertices, numExternal * sizeof(uint64_t),
                                  cudaMemcpyDeviceToHost, s));
    MLSB_CUDA(cudaMemcpyAsync(hVerts, w->welded, sizes.numVertices * 3 * sizeof(float), cudaMemcpyDeviceToHost, s));
    MLSB_CUDA(cudaStreamSynchronize(s));
    if (sink) sink(&sizes, blob, user);
    return MLSB_OK;
}

int shipRange(mlsb_worker *w, uint32_t pitch, uint32_t zStride, const uint32_t size[3], uint32_t zFirst, uint32_t zLast,
              const uint32_t keyOffset[3], mlsb_mesh_sink sink, void *user)
{
    int st = shipBlock(w, pitch, zStride, size, zFirst, zLast, keyOffset, sink, user);
    if (st != MLSB_ERR_OVERFLOW) return st;
    if (zLast - zFirst <= 1)
    {
        setError("a single slice of cells does not fit in meshBytes");
        return MLSB_ERR_OVERFLOW;
    }
    // Too big for the mesh memory: split the slice range and ship the halves separately; the
    // seam is welded by the consumer through the external vertex keys.
    uint32_t mid = zFirst + (zLast - zFirst) / 2;
    MLSB_TRY(shipRange(w, pitch, zStride, size, zFirst, mid, keyOffset, sink, user));
    return shipRange(w, pitch, zStride, size, mid, zLast, keyOffset, sink, user);
}

}  // namespace

int marchingGenerate(mlsb_worker *w, bool useMls, const float *hostField, const uint32_t size[3],
                     const uint32_t keyOffset[3], mlsb_mesh_sink sink, void *user)
{
    // src/marching.cpp:767-769
    for (int i = 0; i < 3; i++)
        MLSB_REQUIRE(size[i] >= 1 && size[i] <= w->cfg.maxCells + 1, MLSB_ERR_LENGTH, "size exceeds the worker's maximum");
    for (int i = 0; i < 3; i++)
        MLSB_REQUIRE((uint64_t) keyOffset[i] + size[i] <= (1u << (kKeyAxisBits - 1)) - 1, MLSB_ERR_LENGTH,
                     "global coordinates exceed MAX_GLOBAL_DIMENSION");
    const uint32_t pitch = roundUp(size[0], (uint32_t) kMlsWgs);
    const uint32_t zStride = roundUp(size[1], (uint32_t) kMlsWgs);
    if (useMls)
    {
        mlsb_swathe sw{size[0], size[1], zStride, 0, 0, size[2] - 1};
        MLSB_TRY(mlsCorners(w, w->field, pitch, sw));
    }
    else
    {
        MLSB_REQUIRE(hostField != nullptr, MLSB_ERR_INVALID_ARGUMENT, "null field");
        for (uint32_t z = 0; z < size[2]; z++)
            MLSB_CUDA(cudaMemcpy2DAsync(w->field + (size_t) z * zStride * pitch, pitch * sizeof(float),
                                        hostField + (size_t) z * size[1] * size[0], size[0] * sizeof(float),
                                        size[0] * sizeof(float), size[1], cudaMemcpyHostToDevice, w->stream));
    }
    if (size[0] < 2 || size[1] < 2 || size[2] < 2)
    {
        MLSB_CUDA(cudaStreamSynchronize(w->stream));
        return MLSB_OK;
    }
    return shipRange(w, pitch, zStride, size, 0, size[2] - 1, keyOffset, sink, user);
}

}  // namespace mlsb
