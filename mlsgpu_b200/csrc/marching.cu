// Marching tetrahedra + vertex welding + scale/bias: Marching::generate / addSlices / shipOut
// (src/marching.cpp:500-824), kernels/marching.cl and kernels/scale_bias.cl, re-designed around
// a field that is resident for the whole bin (no swathes, no copySlice) and around the fact that
// a vertex key IS a position on the half-cell lattice of edge midpoints (computeKey,
// kernels/marching.cl:148-154: 2 * cell + the offsets of the edge's two corners).  The reference
// emits every vertex once per cell that uses it, sorts (key, vertex) pairs, drops duplicates and
// re-indexes; here the welded vertex set is a bitmask over that lattice and a vertex's welded index
// is a prefix popcount, so there is no per-cell vertex, no sort and no re-index pass:
//
//   generateCells<classify>  genOccupied (:84-120) bit-parallel on ballot masks; every warp row
//                   that holds surface also ORs the lattice bits of its valid cells' crossed edges
//                   (19 edges = 9 lattice rows x {even, odd} columns, a handful of logic ops each)
//   (tile scan)     per-tile index bases (the role of clogs::Scan at src/marching.cpp:721): one extra
//                   CTA of the scanLattice launch, so it costs no launch and no time of its own
//   scanLattice     countUniqueVertices + scan + compactVertices + scaleBiasVertices (:271-326,
//                   kernels/scale_bias.cl:32-40) with generateElements' interpolation (:130-138):
//                   one decoupled look-back scan over the lattice words in (z, y, x) order = key
//                   order gives every word its exclusive vertex count, split into unflagged /
//                   flagged (the external flag is the top key bit, so flagged vertices sort last),
//                   and the thread that holds a non-empty word writes its vertices straight to
//                   their final positions
//   generateCells<emit>  one thread per index: cell -> edge -> lattice point -> rank (reindex, :334-340)
//
// tests/lattice_model.py states the same identity in numpy and tests/test_lattice_model.py checks
// it against the oracle bit for bit.  Cells are visited in x-fastest linear order (the reference's
// atomic order is arbitrary), so the output is deterministic: vertices sorted by key (internal
// first), triangles in cell order.
#include "common.cuh"
#include "scan.cuh"
#include <algorithm>
#include <cstring>

namespace mlsb
{

namespace
{

// Measured on the 70-bin C2 step (one worker, ms per step; profiles/r2_summary.md): the index pass is
// bound by latency and by how many independent loads it has in flight, not by instruction count —
// a thread per INDEX (binary search, table walk and record load for every index) takes 4.2 ms, a thread
// per occupied CELL that ranks each vertex once (a third of the instructions) 4.9-5.5 ms; having all
// corner rows of a warp in flight before the first ballot helps the index pass (4.5 -> 4.2 ms, its CTAs
// start from a cold field) and not the classification (3.1 -> 3.2 ms).
#ifndef MLSB_MC_PREFETCH
#define MLSB_MC_PREFETCH EMIT      /* a C++ expression in the kernel's scope, not a preprocessor value */
#endif
// CTAs of 256 threads per SM that the register allocation of the classification and of the index pass aims for.
// An explicit bound of 1 makes ptxas spend MORE registers than no bound at all (classification 30 -> 40, index pass
// 64 -> 70, and with the same treatment the lattice pass 44 -> 48 and the octree's bucket sort 47 -> 64: the
// 12-worker step went from 26.4 to 28.0 ms, 379 -> 357 Msplats/s), so the bounds restate what ptxas chose on its
// own: 8 (<= 32 registers) and 4 (<= 64).  Capping the index pass harder does not pay: 5 (48 registers) and 6 (40,
// 96 bytes of spills) measured 377 and 377-378 Msplats/s against 380 at 4; capping the lattice pass or the bucket
// sort at 32 registers is within the noise as well.
#ifndef MLSB_MC_CLASSIFY_MINBLOCKS
#define MLSB_MC_CLASSIFY_MINBLOCKS 8
#endif
#ifndef MLSB_MC_EMIT_MINBLOCKS
#define MLSB_MC_EMIT_MINBLOCKS 4
#endif
#ifndef MLSB_MC_EMIT
#define MLSB_MC_EMIT 0
#endif
constexpr int kNumEdges = 19;
constexpr int kGenThreads = 256;

// counters in mlsb_worker::marchCounters
enum
{
    CTR_VERTS = 0,      // per-cell (unwelded) vertices: what the reference sizes its buffers by
    CTR_INDICES = 1,
    CTR_OVERFLOW = 2,
    CTR_WELDED = 3,
    CTR_FIRST_EXTERNAL = 4,
    CTR_CELLS = 5,
    CTR_TILES = 6,      // non-empty tiles
    CTR_UNFLAGGED = 7,  // welded vertices without the external flag (internal + the top z plane)
    CTR_GEN_TICKET = 8,
    CTR_WELD_TICKET = 9,
    CTR_PACKED = 10,    // host copy only: the pack kernel has written the block into the pinned blob
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
    uint32_t zFirst;               // first corner slice of the block: tile 0 holds the cells between zFirst and zFirst + 1
    uint32_t *indices;
    uint32_t *counters;
    // lattice of edge midpoints of this block: [2 * (zLast - zFirst) + 1][latRows][latStride] words,
    // bit x2 of row (y2, z2) = the edge with midpoint (x2, y2, z2) / 2 carries a vertex
    uint32_t *latMask;
    const uint4 *latRecord;        // per non-empty word: (mask, vertices before it: unflagged, flagged, -)
    uint32_t latStride, latRows;
    uint64_t *tileTotals;          // per tile: classify writes (v | i << 21 | cells << 42), scanTiles turns it
                                   // into the tile's exclusive base (v | i << 31)
    uint32_t *tileList;            // non-empty tiles in order; count in counters[CTR_TILES]
    const uint8_t *blockActive;    // optional [bz][by][bx] hint: 0 = every corner of that 8x8x8 block is NaN
    uint32_t activeDim;
    const Tables *tables;
};

__device__ __forceinline__ bool finiteF(float x) { return (__float_as_uint(x) & 0x7F800000u) != 0x7F800000u; }

// Tile = kGenRows rows of cells of one z slab; tiles are numbered z-major, then by row group, so
// the x-fastest linear cell order of the output is preserved.
//
// Classification is bit-parallel: for every corner row a warp turns 32 loaded field values into
// two ballot masks (sign >= 0, finite); the occupancy of 32 cells is then a handful of logic
// operations on the 16 masks of the cells' 8 corners, i.e. about one instruction per cell, and
// the field is read straight from global memory, coalesced, exactly once per (tile, row) plus the
// shared rows between tiles.  Only lanes that own an occupied cell (a few percent) assemble the
// 8-bit cube code and look up its counts.
constexpr int kGenRows = 8;
constexpr int kGenWarps = kGenThreads / 32;

struct RowMasks
{
    uint32_t s0, s1;   // sign bits (iso >= 0) at corner x / x+1
    uint32_t f0, f1;   // finite bits at corner x / x+1
};

// The 33 corner values of one row of one slice that the 32 cells starting at x0 need: lane L holds
// corner x0 + L, lane 0 also the 33rd.  Loading and ballot are separate so that a warp can have the
// loads of ALL its rows in flight before it consumes the first one (nine dependent round trips to
// memory otherwise).
struct RowValues
{
    float v, e;
};

__device__ __forceinline__ RowValues loadRowValues(const float *__restrict__ row, uint32_t x0, uint32_t lane, uint32_t numCorners)
{
    const uint32_t x = x0 + lane;
    RowValues r;
    r.v = x < numCorners ? __ldg(row + x) : __uint_as_float(0x7FC00000u);
    r.e = __uint_as_float(0x7FC00000u);
    if (lane == 0 && x0 + 32 < numCorners) r.e = __ldg(row + x0 + 32);
    return r;
}

// Masks of one corner row of one slice for the 32 cells starting at cell x0.
__device__ __forceinline__ RowMasks rowMasks(const RowValues rv)
{
    const float v = rv.v, e = rv.e;
    const uint32_t s = __ballot_sync(0xFFFFFFFFu, v >= 0.0f);
    const uint32_t f = __ballot_sync(0xFFFFFFFFu, finiteF(v));
    const uint32_t eb = __shfl_sync(0xFFFFFFFFu, (uint32_t) (e >= 0.0f) | ((uint32_t) finiteF(e) << 1), 0);
    RowMasks m;
    m.s0 = s;
    m.f0 = f;
    m.s1 = (s >> 1) | ((eb & 1u) << 31);
    m.f1 = (f >> 1) | ((eb >> 1) << 31);
    return m;
}

// 16 bits -> the even bit positions of 32
__device__ __forceinline__ uint32_t spread16(uint32_t v)
{
    v &= 0xFFFFu;
    v = (v | (v << 8)) & 0x00FF00FFu;
    v = (v | (v << 4)) & 0x0F0F0F0Fu;
    v = (v | (v << 2)) & 0x33333333u;
    v = (v | (v << 1)) & 0x55555555u;
    return v;
}

// Which bits of lattice word (wcol, y2, z2) carry the external flag of computeKey
// (kernels/marching.cl:148-154): x2 == 0, y2 == 0, z2 == top.z (the block's first plane),
// x2 == top.x, y2 == top.y.  The block's LAST plane is external too (key >= minExternalKey,
// :304) but unflagged: it sorts behind the internal vertices on its own.
__device__ __forceinline__ uint32_t latticeFlagMask(uint32_t wcol, uint32_t y2, uint32_t z2, uint32_t topX, uint32_t topY)
{
    if (z2 == 0 || y2 == 0 || y2 == topY) return 0xFFFFFFFFu;
    uint32_t m = wcol == 0 ? 1u : 0u;
    if (wcol == (topX >> 5)) m |= 1u << (topX & 31);
    return m;
}

// Welded index of the vertex at lattice point (x2, y2, z2) of the block.
__device__ __forceinline__ uint32_t latticeRank(const uint4 *__restrict__ record, uint32_t stride, uint32_t rows, uint32_t x2,
                                                uint32_t y2, uint32_t z2, uint32_t topX, uint32_t topY, uint32_t numUnflagged)
{
    const uint32_t wcol = x2 >> 5, bit = x2 & 31;
    const size_t word = ((size_t) z2 * rows + y2) * stride + wcol;
    const uint4 rec = __ldg(record + word);   // one 16-byte load: mask and both ranks
    const uint32_t fm = latticeFlagMask(wcol, y2, z2, topX, topY);
    const uint32_t below = (1u << bit) - 1;
    return (fm >> bit) & 1 ? numUnflagged + rec.z + __popc(rec.x & fm & below) : rec.y + __popc(rec.x & ~fm & below);
}

// EMIT = false: classify the tile's cells and write its (vertices, indices, cells) totals; warp
//               rows that hold surface also OR their crossed edges into the lattice mask.
// EMIT = true:  re-classify a NON-EMPTY tile (blockIdx.x indexes tileList), list its occupied cells
//               in shared memory and emit one index per thread at the base that the tile scan
//               computed, so that all stores are contiguous.
// Splitting the pass keeps every CTA free of inter-CTA waits; only the few percent of tiles that
// hold surface are classified twice.  CHUNKS = ceil(cellsX / 256).
template <bool EMIT, int CHUNKS> __global__ void __launch_bounds__(kGenThreads, EMIT ? MLSB_MC_EMIT_MINBLOCKS : MLSB_MC_CLASSIFY_MINBLOCKS) generateCellsKernel(GenParams P)
{
    constexpr int kVRows = kGenRows * CHUNKS;
    extern __shared__ uint32_t sList[];   // EMIT: [2][kVRows * 256]: cell, index base; then [13][256] vertex ranks
    __shared__ uint64_t sWarpTot[kVRows * kGenWarps];
    __shared__ uint64_t sScan[33];

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (EMIT && blockIdx.x >= P.counters[CTR_TILES]) return;
    const uint32_t tile = EMIT ? P.tileList[blockIdx.x] : blockIdx.x;
    const uint32_t rowGroups = (P.cellsY + kGenRows - 1) / kGenRows;
    const uint32_t z = P.zFirst + tile / rowGroups;
    const uint32_t y0 = (tile % rowGroups) * kGenRows;
    const uint32_t nRows = min((uint32_t) kGenRows, P.cellsY - y0);
    const uint32_t numCorners = P.cellsX + 1;
    const Tables &T = *P.tables;
    const float *slice0 = P.field + ((size_t) z * P.zStride + y0) * P.pitch;
    const float *slice1 = slice0 + (size_t) P.zStride * P.pitch;
    if (EMIT)
    {
        // (row, warp) totals are only written where a warp finds an occupied cell
        for (uint32_t i = tid; i < kVRows * kGenWarps; i += kGenThreads) sWarpTot[i] = 0;
        __syncthreads();
    }

    // ---- genOccupied (kernels/marching.cl:84-120), 32 cells per warp and row at a time ----
    // per virtual row v = r * CHUNKS + c: this lane's packed count (vertices | indices << 10 |
    // cell << 22, exclusive over the warp) and cube code (0 = not occupied)
    uint32_t excl[kVRows], codes[kVRows];
    uint64_t mine = 0;
#pragma unroll
    for (int c = 0; c < CHUNKS; c++)
    {
        const uint32_t x0 = c * kGenThreads + warp * 32;
        bool active = x0 < P.cellsX;     // warp-uniform
        // a cell whose lowest corner lies in a block without splats is invalid: if none of the four
        // blocks this warp spans has any, skip its rows altogether (tiles are aligned to block rows)
        if (active && P.blockActive != nullptr)
            active = *reinterpret_cast<const uint32_t *>(P.blockActive + ((size_t) (z >> 3) * P.activeDim + (y0 >> 3)) * P.activeDim + (x0 >> 3)) != 0;
        RowMasks p0 = {0, 0, 0, 0}, p1 = {0, 0, 0, 0};
        constexpr bool kPrefetch = MLSB_MC_PREFETCH;
        RowValues rv0[kGenRows + 1], rv1[kGenRows + 1];
        if (kPrefetch && active)
        {
#pragma unroll
            for (int r = 0; r <= kGenRows; r++)
                if ((uint32_t) r <= nRows)
                {
                    rv0[r] = loadRowValues(slice0 + (size_t) r * P.pitch, x0, lane, numCorners);
                    rv1[r] = loadRowValues(slice1 + (size_t) r * P.pitch, x0, lane, numCorners);
                }
            p0 = rowMasks(rv0[0]);
            p1 = rowMasks(rv1[0]);
        }
        if (!kPrefetch && active)
        {
            p0 = rowMasks(loadRowValues(slice0, x0, lane, numCorners));
            p1 = rowMasks(loadRowValues(slice1, x0, lane, numCorners));
        }
        const uint32_t inRange = P.cellsX - x0 >= 32 ? 0xFFFFFFFFu : (active ? (1u << (P.cellsX - x0)) - 1 : 0u);
        // per corner row, combined over its four corners per cell column (x, x+1 of both slices):
        // all finite / all outside / any outside.  A cell row then needs five more operations.
        uint32_t pv = p0.f0 & p0.f1 & p1.f0 & p1.f1, pa = p0.s0 & p0.s1 & p1.s0 & p1.s1, po = p0.s0 | p0.s1 | p1.s0 | p1.s1;
#pragma unroll
        for (int r = 0; r < kGenRows; r++)
        {
            const int v = r * CHUNKS + c;
            uint32_t packed = 0, code = 0;
            if (active && (uint32_t) r < nRows)
            {
                const RowMasks c0 = rowMasks(kPrefetch ? rv0[r + 1] : loadRowValues(slice0 + (size_t) (r + 1) * P.pitch, x0, lane, numCorners));
                const RowMasks c1 = rowMasks(kPrefetch ? rv1[r + 1] : loadRowValues(slice1 + (size_t) (r + 1) * P.pitch, x0, lane, numCorners));
                const uint32_t cv = c0.f0 & c0.f1 & c1.f0 & c1.f1, ca = c0.s0 & c0.s1 & c1.s0 & c1.s1, co = c0.s0 | c0.s1 | c1.s0 | c1.s1;
                const uint32_t occ = (pv & cv) & (po | co) & ~(pa & ca) & inRange;
                if (occ)
                {
                    if ((occ >> lane) & 1)
                    {
                        // corner k sits at offset (k&1, (k>>1)&1, k>>2): kernels/marching.cl:98-106
                        code = ((p0.s0 >> lane) & 1) | (((p0.s1 >> lane) & 1) << 1) | (((c0.s0 >> lane) & 1) << 2)
                               | (((c0.s1 >> lane) & 1) << 3) | (((p1.s0 >> lane) & 1) << 4) | (((p1.s1 >> lane) & 1) << 5)
                               | (((c1.s0 >> lane) & 1) << 6) | (((c1.s1 >> lane) & 1) << 7);
                        packed = (uint32_t) T.count[code][0] | ((uint32_t) T.count[code][1] << 10) | (1u << 22);
                    }
                    if (EMIT)
                    {
                        const uint32_t inc = warpInclusiveScan(packed);
                        excl[v] = inc - packed;
                        const uint32_t tot = __shfl_sync(0xFFFFFFFFu, inc, 31);
                        if (lane == 0)
                            sWarpTot[v * kGenWarps + warp] =
                                (uint64_t) (tot & 0x3FF) | ((uint64_t) ((tot >> 10) & 0xFFF) << 21) | ((uint64_t) (tot >> 22) << 42);
                    }
                    else
                    {
                        mine += (uint64_t) (packed & 0x3FF) | ((uint64_t) ((packed >> 10) & 0xFFF) << 21) | ((uint64_t) (packed >> 22) << 42);
                        // ---- lattice bits of this row of cells.  With the sign masks a = corner
                        // (x, y, z), b = (x, y+1, z), c = (x, y, z+1), d = (x, y+1, z+1) and their
                        // x+1 twins, the 19 edges of a cell fall on 9 lattice rows (oy, oz) in
                        // {0,1,2}^2 relative to (2y, 2z): odd columns 2x+1 hold the 9 edges with an
                        // x component, even columns 2x / 2x+2 the 10 without.  Contributions from
                        // neighbouring cells, rows and slabs meet by OR. ----
                        const uint32_t a = p0.s0, a1 = p0.s1, b = c0.s0, b1 = c0.s1, cc = p1.s0, cc1 = p1.s1, d = c1.s0, d1 = c1.s1;
                        const uint32_t valid = pv & cv & inRange;
                        // lane L handles row (oy, oz) = (L % 3, L / 3): its edges run from corner row
                        // (y + ly, z + lz) to (y + hy, z + hz), where an offset of 0 / 1 / 2 means
                        // (0,0) / (0,1) / (1,1); when both offsets are even lo == hi and the even
                        // columns are empty by themselves
                        const uint32_t oy = lane % 3, oz = (lane / 3) % 3;
                        const bool ly = oy == 2, hy = oy >= 1, lz = oz == 2, hz = oz >= 1;
                        const uint32_t lo = lz ? (ly ? d : cc) : (ly ? b : a), lo1 = lz ? (ly ? d1 : cc1) : (ly ? b1 : a1);
                        const uint32_t hi = hz ? (hy ? d : cc) : (hy ? b : a), hi1 = hz ? (hy ? d1 : cc1) : (hy ? b1 : a1);
                        uint32_t odd = lo ^ hi1, even = lo ^ hi;
                        const uint32_t evenNext = lo1 ^ hi1;
                        if (lane < 9)
                        {
                            odd &= valid;
                            // corner x belongs to cells x-1 and x; the one beyond this warp's 32 goes
                            // to bit 0 of the next word pair
                            even &= valid | (valid << 1);
                            const uint32_t carry = (evenNext & valid) >> 31;
                            const uint32_t y2 = 2 * (y0 + r) + oy, z2 = 2 * (z - P.zFirst) + oz;
                            uint32_t *const row = P.latMask + ((size_t) z2 * P.latRows + y2) * P.latStride + (x0 >> 4);
                            const uint32_t w0 = spread16(even) | (spread16(odd) << 1);
                            const uint32_t w1 = spread16(even >> 16) | (spread16(odd >> 16) << 1);
                            if (w0) atomicOr(row, w0);
                            if (w1) atomicOr(row + 1, w1);
                            if (carry) atomicOr(row + 2, 1u);
                        }
                    }
                }
                p0 = c0;
                p1 = c1;
                pv = cv;
                pa = ca;
                po = co;
            }
            if (EMIT)
            {
                codes[v] = code;
                if (code == 0) excl[v] = 0;
            }
        }
    }

    if (!EMIT)
    {
        mine = warpSum(mine);
        if (lane == 0) sWarpTot[warp] = mine;
        __syncthreads();
        if (tid == 0)
        {
            uint64_t total = 0;
            for (int i = 0; i < kGenWarps; i++) total += sWarpTot[i];
            P.tileTotals[tile] = total;
        }
        return;
    }

    // ---- scanElements inside the tile: exclusive bases of every (virtual row, warp) ----
    __syncthreads();
    uint64_t total;
    {
        uint64_t t = tid < kVRows * kGenWarps ? sWarpTot[tid] : 0ull;
        uint64_t e = blockExclusiveScan(t, sScan, &total);
        if (tid < kVRows * kGenWarps) sWarpTot[tid] = e;
    }
    __syncthreads();
    const uint32_t numCells = (uint32_t) (total >> 42);
    const uint32_t numIdx = (uint32_t) ((total >> 21) & 0x1FFFFF);
    uint32_t *const sCell = sList, *const sIloc = sList + kVRows * kGenThreads;
#pragma unroll
    for (int v = 0; v < kVRows; v++)
    {
        if (codes[v])
        {
            const uint64_t b = sWarpTot[v * kGenWarps + warp];
            const uint32_t idx = (uint32_t) (b >> 42) + (excl[v] >> 22);
            const uint32_t x = (v % CHUNKS) * kGenThreads + tid, r = v / CHUNKS;
            sCell[idx] = x | (r << 16) | (codes[v] << 24);
            sIloc[idx] = (uint32_t) ((b >> 21) & 0x1FFFFF) + ((excl[v] >> 10) & 0xFFF);
        }
    }
    __syncthreads();

    const uint32_t iBase = (uint32_t) (P.tileTotals[tile] >> 31);
    if (P.counters[CTR_OVERFLOW]) return;   // the block did not fit: the host splits it and starts over

    // ---- generateElements' index half + reindex (kernels/marching.cl:240-258, :334-340).  MLSB_MC_EMIT:
    // 0 = one thread per index (binary search for its cell, table walk and record load for every index),
    // 1 / 2 = one thread per occupied cell: it ranks the cell's <= 13 vertices once into its own column
    // of shared memory (2: unrolled and predicated) and then writes the cell's indices from there ----
    const uint32_t numUnflagged = P.counters[CTR_UNFLAGGED];
    const uint32_t topX = 2 * P.cellsX, topY = 2 * P.cellsY;
#if MLSB_MC_EMIT == 0
    for (uint32_t j = tid; j < numIdx; j += kGenThreads)
    {
        // last cell whose first index is <= j
        uint32_t lo = 0, hi = numCells;
        while (hi - lo > 1)
        {
            const uint32_t mid = (lo + hi) >> 1;
            if (sIloc[mid] <= j) lo = mid; else hi = mid;
        }
        const uint32_t cell = sCell[lo], k = j - sIloc[lo];
        const uint32_t x = cell & 0xFFFF, r = (cell >> 16) & 0xFF, code = cell >> 24;
        const uint32_t e = T.data[T.start[code][0] + T.data[T.start[code][1] + k]];
        const uint32_t a = T.edge[e][0], b = T.edge[e][1];
        const uint32_t x2 = 2 * x + (a & 1) + (b & 1);
        const uint32_t y2 = 2 * (y0 + r) + ((a >> 1) & 1) + ((b >> 1) & 1);
        const uint32_t z2 = 2 * (z - P.zFirst) + (a >> 2) + (b >> 2);
        P.indices[iBase + j] = latticeRank(P.latRecord, P.latStride, P.latRows, x2, y2, z2, topX, topY, numUnflagged);
    }
#else
    (void) numIdx;
    uint32_t *const sRank = sList + 2 * kVRows * kGenThreads;   // [kMaxCellVertices][kGenThreads]
    for (uint32_t ci = tid; ci < numCells; ci += kGenThreads)
    {
        const uint32_t cell = sCell[ci];
        const uint32_t x = cell & 0xFFFF, r = (cell >> 16) & 0xFF, code = cell >> 24;
        const uint32_t v0 = T.start[code][0], i0 = T.start[code][1];
        const uint32_t nV = T.count[code][0], nI = T.count[code][1];
#if MLSB_MC_EMIT == 2
        // unrolled and predicated, so that the record loads of all vertices are in flight together
#pragma unroll
        for (uint32_t v = 0; v < kMaxCellVertices; v++)
            if (v < nV)
#else
        for (uint32_t v = 0; v < nV; v++)
#endif
            {
                const uint32_t e = T.data[v0 + v];
                const uint32_t a = T.edge[e][0], b = T.edge[e][1];
                // computeKey(), :148-154: twice the cell plus the offsets of the edge's two corners
                const uint32_t x2 = 2 * x + (a & 1) + (b & 1);
                const uint32_t y2 = 2 * (y0 + r) + ((a >> 1) & 1) + ((b >> 1) & 1);
                const uint32_t z2 = 2 * (z - P.zFirst) + (a >> 2) + (b >> 2);
                sRank[v * kGenThreads + tid] = latticeRank(P.latRecord, P.latStride, P.latRows, x2, y2, z2, topX, topY, numUnflagged);
            }
        uint32_t *const out = P.indices + iBase + sIloc[ci];
        for (uint32_t k = 0; k < nI; k++) out[k] = sRank[(uint32_t) T.data[i0 + k] * kGenThreads + tid];
    }
#endif
}

// Exclusive scan of the tile totals, list of non-empty tiles, grand totals -> counters: the role of
// clogs::Scan(uint2) at src/marching.cpp:721.  Runs as ONE EXTRA CTA of the scanLattice launch (the lattice
// scan does not need the tile bases, the index pass after it does), so it costs neither a launch
// nor time of its own.
__device__ void scanTiles(uint64_t *tileTotals, uint32_t *tileList, uint32_t numTiles, uint32_t *counters,
                          uint32_t vertexSpace, uint32_t indexSpace)
{
    __shared__ uint64_t sScan[33];
    __shared__ uint32_t sScan32[33];
    const uint32_t threads = blockDim.x;
    const uint32_t per = (numTiles + threads - 1) / threads;
    const uint32_t first = min(threadIdx.x * per, numTiles), last = min(first + per, numTiles);
    uint64_t sum = 0, cells = 0;
    uint32_t nonEmpty = 0;
    for (uint32_t t = first; t < last; t++)
    {
        const uint64_t v = tileTotals[t];
        sum += (v & 0x1FFFFF) | (((v >> 21) & 0x1FFFFF) << 31);
        cells += v >> 42;
        nonEmpty += v != 0;
    }
    uint64_t total;
    uint32_t totalNonEmpty;
    uint64_t base = blockExclusiveScan(sum, sScan, &total);
    uint32_t listPos = blockExclusiveScan(nonEmpty, sScan32, &totalNonEmpty);
    for (uint32_t t = first; t < last; t++)
    {
        const uint64_t v = tileTotals[t];
        tileTotals[t] = base;
        base += (v & 0x1FFFFF) | (((v >> 21) & 0x1FFFFF) << 31);
        if (v != 0) tileList[listPos++] = t;
    }
    if (cells) atomicAdd(&counters[CTR_CELLS], (uint32_t) cells);
    if (threadIdx.x == 0)
    {
        counters[CTR_VERTS] = (uint32_t) (total & 0x7FFFFFFFu);
        counters[CTR_INDICES] = (uint32_t) (total >> 31);
        counters[CTR_TILES] = totalNonEmpty;
        if ((total & 0x7FFFFFFFu) > vertexSpace || (total >> 31) > indexSpace) counters[CTR_OVERFLOW] = 1;
    }
}

struct LatticeParams
{
    const uint32_t *mask;
    uint4 *record;                   // out, per non-empty word: (mask, unflagged before, flagged before, 0)
    uint32_t stride, rows, planes;   // words per row, rows (y2) per plane, planes (z2)
    uint32_t totalWords;
    uint32_t topX, topY;             // 2 * cellsX, 2 * cellsY
    uint32_t *counters;
    volatile uint64_t *status;       // per 2048-word chunk: (unflagged | flagged << 32) vertex counts
    // emitVertices
    const float *field;
    uint32_t pitch, zStride, zFirst;
    uint3 gridOffset;
    float *welded, *weldedFlagged;   // unflagged vertices by rank; flagged ones by their own rank
    uint64_t *weldedKeys, *weldedKeysFlagged;
    uint2 *compact, *compactFlagged; // optional: (t, lattice point) per vertex for the compact read-back
    uint64_t keyOffset;              // src/marching.cpp:594-599
    float scale, bx, by, bz;         // ScaleBiasFilter
    uint32_t vertexSpace, indexSpace;
    // the tile scan riding along (last CTA of the grid)
    uint64_t *tileTotals;
    uint32_t *tileList;
    uint32_t numTiles;
};

constexpr int kLatThreads = 256;
constexpr int kLatItems = 8;

// countUniqueVertices + scan + compactVertices + scaleBiasVertices (kernels/marching.cl:271-326,
// src/marching.cpp:721, kernels/scale_bias.cl:32-40) in two launches without any waiting between
// CTAs: latticeCountKernel leaves the (unflagged, flagged) vertex count of every 2048-word chunk of
// the lattice, and every CTA of latticeEmitKernel adds up the counts of the chunks in front of its own
// (at most ~2200 values from L2, 9 per thread) before it ranks and writes its vertices.  A decoupled
// look-back in one launch measured slower: its chain is ~2000 tiles of a few microseconds each, and a
// third of all warp stalls were spins on predecessors' status words.
// A vertex is interpolated from the lower corner A = lattice >> 1 of its edge along d = lattice & 1 —
// the corner order of the edge table — as interp() does (:130-138).  Flagged vertices follow ALL
// unflagged ones in the output, so they go to a buffer of their own and the readback concatenates.
__global__ void __launch_bounds__(kLatThreads) latticeCountKernel(LatticeParams P)
{
    __shared__ uint64_t sScan[33];
    if (blockIdx.x == gridDim.x - 1)
    {
        scanTiles(P.tileTotals, P.tileList, P.numTiles, P.counters, P.vertexSpace, P.indexSpace);
        return;
    }
    const uint32_t tile = blockIdx.x;
    const uint32_t base = tile * (kLatThreads * kLatItems) + threadIdx.x * kLatItems;
    uint32_t wcol = base % P.stride, row = base / P.stride;
    uint32_t y2 = row % P.rows, z2 = row / P.rows;
    uint32_t nU = 0, nF = 0;
    uint32_t m[kLatItems];
#pragma unroll
    for (int i = 0; i < kLatItems; i++) m[i] = base + i < P.totalWords ? P.mask[base + i] : 0u;
#pragma unroll
    for (int i = 0; i < kLatItems; i++)
    {
        if (m[i])
        {
            const uint32_t fm = latticeFlagMask(wcol, y2, z2, P.topX, P.topY);
            nU += __popc(m[i] & ~fm);
            nF += __popc(m[i] & fm);
        }
        if (++wcol == P.stride)
        {
            wcol = 0;
            if (++y2 == P.rows)
            {
                y2 = 0;
                z2++;
            }
        }
    }
    uint64_t total;
    blockExclusiveScan((uint64_t) nU | ((uint64_t) nF << 32), sScan, &total);
    if (threadIdx.x == 0) const_cast<uint64_t *>(P.status)[tile] = total;
}

#ifdef MLSB_MC_LATEMIT_MINBLOCKS
__global__ void __launch_bounds__(kLatThreads, MLSB_MC_LATEMIT_MINBLOCKS) latticeEmitKernel(LatticeParams P, uint32_t numChunks)
#else
__global__ void __launch_bounds__(kLatThreads) latticeEmitKernel(LatticeParams P, uint32_t numChunks)   // 44 registers; capped at 32 (spills) it measures the same
#endif
{
    __shared__ uint64_t sScan[33];
    __shared__ uint64_t sBefore;
    const uint32_t tile = blockIdx.x;
    const uint64_t *chunkCounts = const_cast<const uint64_t *>(P.status);
    const uint64_t own = chunkCounts[tile];
    const uint32_t lastPlane = (P.planes - 1) * P.rows * P.stride;   // its vertices are external, the ones before it are not
    const uint32_t tileBase = tile * (kLatThreads * kLatItems);
    const bool holdsLastPlane = tileBase <= lastPlane && lastPlane < tileBase + kLatThreads * kLatItems;
    const bool holdsLastWord = tileBase < P.totalWords && P.totalWords <= tileBase + kLatThreads * kLatItems;
    if (own == 0 && !holdsLastPlane && !holdsLastWord) return;   // nothing to rank, nothing to report: most chunks
    // ---- vertices in front of this chunk: everyone adds up the counts of the chunks before it ----
    {
        uint64_t part = 0;
        for (uint32_t t = threadIdx.x; t < tile; t += kLatThreads) part += chunkCounts[t];
        uint64_t all;
        blockExclusiveScan(part, sScan, &all);
        if (threadIdx.x == 0) sBefore = all;
        __syncthreads();
    }
    const uint32_t base = tileBase + threadIdx.x * kLatItems;
    const uint32_t row0 = base / P.stride, wcol0 = base % P.stride;
    const uint32_t y20 = row0 % P.rows, z20 = row0 / P.rows;
    uint32_t wcol = wcol0, y2 = y20, z2 = z20;
    auto advance = [&]() {
        if (++wcol == P.stride)
        {
            wcol = 0;
            if (++y2 == P.rows)
            {
                y2 = 0;
                z2++;
            }
        }
    };
    uint32_t m[kLatItems];
    uint32_t nU = 0, nF = 0, beforeLastPlane = 0;
#pragma unroll
    for (int i = 0; i < kLatItems; i++) m[i] = base + i < P.totalWords ? P.mask[base + i] : 0u;
#pragma unroll
    for (int i = 0; i < kLatItems; i++)
    {
        if (base + i == lastPlane) beforeLastPlane = nU;
        const uint32_t fm = latticeFlagMask(wcol, y2, z2, P.topX, P.topY);
        nU += __popc(m[i] & ~fm);
        nF += __popc(m[i] & fm);
        advance();
    }
    uint64_t total;
    const uint64_t excl = blockExclusiveScan((uint64_t) nU | ((uint64_t) nF << 32), sScan, &total);
    const uint64_t mine = sBefore + excl;
    uint32_t rankU = (uint32_t) mine, rankF = (uint32_t) (mine >> 32);
    // internal vertices = unflagged ones in front of the last plane
    if (base <= lastPlane && lastPlane < base + kLatItems) P.counters[CTR_FIRST_EXTERNAL] = rankU + beforeLastPlane;
    if (base + kLatItems >= P.totalWords && base < P.totalWords)
    {
        // the thread holding the last word: grand totals
        P.counters[CTR_UNFLAGGED] = rankU + nU;
        P.counters[CTR_WELDED] = rankU + nU + rankF + nF;
    }
    // ---- the tile's vertices, one per thread: the words are listed in shared memory with their
    // exclusive vertex count inside the tile, thread j finds the word that holds the tile's j-th
    // vertex and the bit inside it.  (Letting the owner of a word emit all its vertices leaves the
    // work — two dependent field loads per vertex — to a handful of threads.) ----
    __shared__ uint32_t sMask[kLatThreads * kLatItems], sCount[kLatThreads * kLatItems];
    __shared__ uint32_t sRankU[kLatThreads * kLatItems], sRankF[kLatThreads * kLatItems];
    const uint32_t tileVerts = (uint32_t) total + (uint32_t) (total >> 32);
    if (tileVerts == 0) return;   // uniform over the CTA
    {
        uint32_t c = (uint32_t) excl + (uint32_t) (excl >> 32);
        wcol = wcol0, y2 = y20, z2 = z20;
#pragma unroll
        for (int i = 0; i < kLatItems; i++, advance())
        {
            const uint32_t l = threadIdx.x * kLatItems + i;
            sMask[l] = m[i];
            sCount[l] = c;
            sRankU[l] = rankU;
            sRankF[l] = rankF;
            if (m[i])
            {
                P.record[base + i] = make_uint4(m[i], rankU, rankF, 0u);
                const uint32_t fm = latticeFlagMask(wcol, y2, z2, P.topX, P.topY);
                const uint32_t f = __popc(m[i] & fm), u = __popc(m[i]) - f;
                c += u + f;
                rankU += u;
                rankF += f;
            }
        }
    }
    __syncthreads();
    for (uint32_t j = threadIdx.x; j < tileVerts; j += kLatThreads)
    {
        // last word whose exclusive count is <= j (empty words share the count of their successor)
        uint32_t lo = 0, hi = kLatThreads * kLatItems;
        while (hi - lo > 1)
        {
            const uint32_t mid = (lo + hi) >> 1;
            if (sCount[mid] <= j) lo = mid; else hi = mid;
        }
        const uint32_t bits = sMask[lo];
        uint32_t rest = bits;
        for (uint32_t k = j - sCount[lo]; k > 0; k--) rest &= rest - 1;   // drop the k lowest set bits
        const uint32_t bit = __ffs(rest) - 1;
        const uint32_t word = tileBase + lo;
        const uint32_t row = word / P.stride, wc = word % P.stride;
        const uint32_t vy2 = row % P.rows, vz2 = row / P.rows;
        const uint32_t fm = latticeFlagMask(wc, vy2, vz2, P.topX, P.topY);
        const uint32_t below = (1u << bit) - 1;
        const bool flagged = (fm >> bit) & 1;
        const uint32_t rank = flagged ? sRankF[lo] + __popc(bits & fm & below) : sRankU[lo] + __popc(bits & ~fm & below);
        const uint32_t x2 = wc * 32 + bit, ax = x2 >> 1, dx = x2 & 1;
        const uint32_t ay = vy2 >> 1, dy = vy2 & 1, az = P.zFirst + (vz2 >> 1), dz = vz2 & 1;
        const float isoA = __ldg(P.field + ((size_t) az * P.zStride + ay) * P.pitch + ax);
        const float isoB = __ldg(P.field + ((size_t) (az + dz) * P.zStride + ay + dy) * P.pitch + ax + dx);
        // no contraction, explicit fma: kernels/marching.cl:133
        const float inv = __fdiv_rn(1.0f, __fsub_rn(isoA, isoB));
        const float t = __fmul_rn(isoA, inv);
        const float px = __fmaf_rn(t, (float) dx, (float) (ax + P.gridOffset.x));
        const float py = __fmaf_rn(t, (float) dy, (float) (ay + P.gridOffset.y));
        const float pz = __fmaf_rn(t, (float) dz, (float) (az + P.gridOffset.z));
        if (rank >= P.vertexSpace) continue;   // the block does not fit: the host splits it and starts over
        float *const out = (flagged ? P.weldedFlagged : P.welded) + 3 * (size_t) rank;
        out[0] = __fmaf_rn(px, P.scale, P.bx);
        out[1] = __fmaf_rn(py, P.scale, P.by);
        out[2] = __fmaf_rn(pz, P.scale, P.bz);
        // what the host needs to redo the four lines above bit for bit (unpackWireMesh): t and the lattice point
        if (P.compact) (flagged ? P.compactFlagged : P.compact)[rank] = make_uint2(__float_as_uint(t), x2 | (vy2 << 11) | (vz2 << 22));
        if (flagged || vz2 == P.planes - 1)
            (flagged ? P.weldedKeysFlagged : P.weldedKeys)[rank] =
                ((uint64_t) x2 | ((uint64_t) vy2 << kKeyAxisBits) | ((uint64_t) (vz2 + 2 * P.zFirst) << (2 * kKeyAxisBits))) + P.keyOffset;
    }
}

// enqueueReadMesh (src/mesh.cpp:62-102) without the host in the loop: the block's sizes live on the
// device, so a copy-engine read-back needs a round trip first (sizes to the host, wait, five
// cudaMemcpyAsync of the right lengths, wait again).  This kernel reads the sizes where they are and
// writes the HostKeyMesh blob — [keys of external vertices][vertices][triangles] — and the counters
// straight into PINNED HOST memory (zero-copy stores, 128 bytes per warp and instruction): one wait per
// block.  If the blob is too small it only reports the sizes; the host grows it and launches again.
// MLSB_READBACK=pack selects it; the copy engine stays the default.  Measured: one GPU, 12 workers — 364–369
// Msplats/s either way with 8 packing CTAs (331 with 64: they hold SM slots for the length of a PCIe transfer);
// eight GPUs on a 32-vCPU box — HBM-resident 1067 against 1345 Msplats/s, from host memory 1124 against 960.
constexpr int kPackThreads = 256, kPackBlocks = 8, kPackUnroll = 8;   // a few CTAs keep PCIe busy; each holds an SM slot for the whole transfer
__global__ void __launch_bounds__(kPackThreads) packMeshKernel(const uint32_t *__restrict__ counters, const float *__restrict__ welded,
                                                               const float *__restrict__ weldedFlagged,
                                                               const uint64_t *__restrict__ keys, const uint64_t *__restrict__ keysFlagged,
                                                               const uint32_t *__restrict__ indices, uint32_t *__restrict__ blob,
                                                               uint64_t capacityBytes, uint32_t *__restrict__ hostCounters,
                                                               uint32_t vertexSpace, uint32_t indexSpace)
{
    const uint32_t nV = counters[CTR_WELDED], nI = counters[CTR_INDICES], nInt = counters[CTR_FIRST_EXTERNAL], nU = counters[CTR_UNFLAGGED];
    const bool overflow = counters[CTR_OVERFLOW] || counters[CTR_VERTS] > vertexSpace || nI > indexSpace;
    const uint64_t nExt = nV - nInt, nTop = nU - nInt, nF = nV - nU;
    const uint64_t words = 2 * nExt + 3 * (uint64_t) nV + nI;
    const bool fits = !overflow && counters[CTR_VERTS] != 0 && 4 * words <= capacityBytes;
    if (blockIdx.x == 0 && threadIdx.x < CTR_COUNT)
        hostCounters[threadIdx.x] = threadIdx.x == CTR_PACKED ? (uint32_t) fits : counters[threadIdx.x];
    if (!fits) return;
    // the blob as 32-bit words: five source segments, in blob order
    const uint32_t *const src0 = reinterpret_cast<const uint32_t *>(keys + nInt);      // keys of the top plane (unflagged externals)
    const uint32_t *const src1 = reinterpret_cast<const uint32_t *>(keysFlagged);
    const uint32_t *const src2 = reinterpret_cast<const uint32_t *>(welded);
    const uint32_t *const src3 = reinterpret_cast<const uint32_t *>(weldedFlagged);
    const uint64_t e0 = 2 * nTop, e1 = e0 + 2 * nF, e2 = e1 + 3 * (uint64_t) nU, e3 = e2 + 3 * nF;
    const uint64_t stride = (uint64_t) gridDim.x * kPackThreads * kPackUnroll;
    for (uint64_t base = ((uint64_t) blockIdx.x * kPackUnroll) * kPackThreads + threadIdx.x; base < words; base += stride)
    {
        uint32_t v[kPackUnroll];
#pragma unroll
        for (int u = 0; u < kPackUnroll; u++)
        {
            const uint64_t i = base + (uint64_t) u * kPackThreads;
            if (i < words)
                v[u] = i < e0 ? src0[i] : i < e1 ? src1[i - e0] : i < e2 ? src2[i - e1] : i < e3 ? src3[i - e2] : indices[i - e3];
        }
#pragma unroll
        for (int u = 0; u < kPackUnroll; u++)
        {
            const uint64_t i = base + (uint64_t) u * kPackThreads;
            if (i < words) blob[i] = v[u];
        }
    }
}

}  // namespace

namespace
{

// MLSB_READBACK=compact.  The HostKeyMesh blob costs 72 bytes of PCIe per input splat on C2 (two vertices of 12
// bytes, four triangles of 12), and with eight GPUs copying at once the box delivers 15 GB/s per GPU
// (profiles/r2_probe_d2h_8gpu.txt): the read-back IS the step.  A block's vertices are determined by their lattice
// point and the interpolation parameter t (latticeEmitKernel), its indices are below 2^21 whenever it has at most
// 2^21 vertices: 8 bytes each.  This kernel gathers the wire blob — [external keys][vertices][triangles], one
// 64-bit word each — in DEVICE memory (no SM waits on PCIe, unlike packMeshKernel); the copy engine moves it and
// the worker thread expands it into the HostKeyMesh blob the sink expects (unpack.cpp), redoing latticeEmitKernel's
// four fused multiply-adds per vertex: bit-identical.  CTR_PACKED tells the host whether the block qualified.
constexpr uint32_t kWireIndexBits = 21;
__global__ void __launch_bounds__(256) compactMeshKernel(uint32_t *__restrict__ counters, const uint64_t *__restrict__ keys,
                                                         const uint64_t *__restrict__ keysFlagged, const uint2 *__restrict__ compact,
                                                         const uint2 *__restrict__ compactFlagged, const uint32_t *__restrict__ indices,
                                                         uint64_t *__restrict__ wire, uint64_t capacityWords, uint32_t vertexSpace,
                                                         uint32_t indexSpace)
{
    const uint32_t nV = counters[CTR_WELDED], nI = counters[CTR_INDICES], nInt = counters[CTR_FIRST_EXTERNAL], nU = counters[CTR_UNFLAGGED];
    const bool overflow = counters[CTR_OVERFLOW] || counters[CTR_VERTS] > vertexSpace || nI > indexSpace;
    const uint64_t nExt = nV - nInt, nTop = nU - nInt, nT = nI / 3;
    const uint64_t words = nExt + nV + nT;
    const bool fits = !overflow && counters[CTR_VERTS] != 0 && nV <= (1u << kWireIndexBits) && words <= capacityWords;
    if (!fits) return;   // CTR_PACKED stays 0 (the block's memset)
    if (blockIdx.x == 0 && threadIdx.x == 0) counters[CTR_PACKED] = 1;
    const uint64_t e0 = nTop, e1 = nExt, e2 = nExt + nU, e3 = nExt + nV;
    const uint64_t stride = (uint64_t) gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < words; i += stride)
    {
        uint64_t v;
        if (i < e0) v = keys[nInt + i];
        else if (i < e1) v = keysFlagged[i - e0];
        else if (i < e3)
        {
            const uint2 c = i < e2 ? compact[i - e1] : compactFlagged[i - e2];
            v = (uint64_t) c.x | ((uint64_t) c.y << 32);
        }
        else
        {
            const uint32_t *t = indices + 3 * (i - e3);
            v = (uint64_t) t[0] | ((uint64_t) t[1] << kWireIndexBits) | ((uint64_t) t[2] << (2 * kWireIndexBits));
        }
        wire[i] = v;
    }
}

}  // namespace

int readbackModeFromEnv()
{
    static const int mode = [] {
        const char *e = getenv("MLSB_READBACK");
        if (e == nullptr) return 0;
        return strcmp(e, "pack") == 0 ? 1 : strcmp(e, "compact") == 0 ? 2 : 0;
    }();
    return mode;
}

int marchingTablesHost(uint8_t *count, uint16_t *start, uint8_t *data, uint32_t *key)
{
    buildTables();
    if (count) memcpy(count, gTables.t.count, sizeof(gTables.t.count));
    if (start) memcpy(start, gTables.t.start, sizeof(gTables.t.start));
    if (data) memcpy(data, gTables.t.data, sizeof(gTables.t.data));
    if (key) memcpy(key, gTables.key, sizeof(gTables.key));
    return MLSB_OK;
}

namespace
{

// Every buffer marchingAlloc makes, in bytes: the ONE place sizes are computed, so that
// mlsb_worker_resource_usage reports what a worker really takes.
struct MarchSizes
{
    uint32_t fieldDim, maxTiles, latStatusWords;
    uint64_t vertexSpace, indexSpace;
    size_t latticeWords;
    size_t field, blockActive, latMask, latRecord, indices, welded, weldedKeys, counters, tileTotals, tileList, tables;
    size_t hostBlobWorst;
    size_t compactVerts, wireBlob;   // 0 unless MLSB_READBACK=compact
};

MarchSizes marchSizes(const mlsb_worker_config &c)
{
    MarchSizes z{};
    z.fieldDim = roundUp(c.maxCells + 1, (uint32_t) kMlsWgs);
    const uint64_t meshCells = c.meshBytes / kMaxCellBytes;
    z.vertexSpace = meshCells * kMaxCellVertices;    // src/marching.cpp:389-392
    z.indexSpace = meshCells * kMaxCellIndices;
    z.field = (size_t) z.fieldDim * z.fieldDim * z.fieldDim * sizeof(float);
    const size_t nb = roundUp(z.fieldDim / kMlsWgs, 4u);
    z.blockActive = nb * nb * nb + 16;
    // lattice of edge midpoints for the largest block: (2 cells + 1)^2 rows of 2 words per 32 cells (+1 carry word)
    z.latticeWords = (size_t) (2 * c.maxCells + 1) * (2 * c.maxCells + 1) * (2 * divUp(c.maxCells, 32u) + 1);
    z.latMask = z.latticeWords * sizeof(uint32_t);
    z.latRecord = z.latticeWords * sizeof(uint4);
    z.indices = (z.indexSpace + 1) * sizeof(uint32_t);
    z.welded = (z.vertexSpace + 1) * 3 * sizeof(float);        // twice: unflagged and flagged vertices
    z.weldedKeys = (z.vertexSpace + 1) * sizeof(uint64_t);     // twice
    z.latStatusWords = (uint32_t) divUp(z.latticeWords, (size_t) (kLatThreads * kLatItems)) + 2;
    z.counters = CTR_COUNT * sizeof(uint32_t) + (size_t) z.latStatusWords * sizeof(uint64_t);
    z.maxTiles = c.maxCells * divUp(c.maxCells, (uint32_t) kGenRows);
    z.tileTotals = (size_t) z.maxTiles * sizeof(uint64_t);
    z.tileList = (size_t) z.maxTiles * sizeof(uint32_t);
    z.tables = sizeof(Tables);
    mlsb_mesh_sizes worst{z.vertexSpace, z.indexSpace / 3, 0};
    z.hostBlobWorst = mlsb_mesh_host_bytes(&worst) + 64;
    if (readbackModeFromEnv() == 2)
    {
        z.compactVerts = (z.vertexSpace + 1) * sizeof(uint2);      // twice: unflagged and flagged
        z.wireBlob = (2 * z.vertexSpace + z.indexSpace / 3 + 1) * sizeof(uint64_t);   // keys + vertices + triangles
    }
    return z;
}

// The read-back blob grows on demand: the reference's sizing rule gives 124 MB of pinned memory per
// worker at 256 MB of mesh memory, the largest C2 block needs 12.
int ensureHostBlob(mlsb_worker *w, size_t bytes)
{
    if (bytes <= w->hostBlobBytes) return MLSB_OK;
    size_t want = std::max(bytes + bytes / 2, (size_t) 4 << 20);
    if (w->hostBlob) MLSB_CUDA(cudaFreeHost(w->hostBlob));
    w->hostBlob = nullptr;
    w->hostBlobBytes = 0;
    MLSB_CUDA(cudaMallocHost(&w->hostBlob, want));
    w->hostBlobBytes = want;
    return MLSB_OK;
}

}  // namespace

uint64_t marchingDeviceBytes(const mlsb_worker_config &c, uint64_t *maxAllocation, uint64_t *pinnedHost)
{
    const MarchSizes z = marchSizes(c);
    const size_t all[] = {z.field, z.blockActive, z.latMask, z.latRecord, z.indices, z.welded, z.welded, z.weldedKeys, z.weldedKeys,
                          z.counters, z.tileTotals, z.tileList, z.tables, z.compactVerts, z.compactVerts, z.wireBlob};
    uint64_t total = 0, largest = 0;
    for (size_t b : all)
    {
        total += b;
        largest = std::max<uint64_t>(largest, b);
    }
    if (maxAllocation) *maxAllocation = largest;
    // the blob's upper bound: it grows on demand (compact read-back: plus the landing buffer of the wire blob)
    if (pinnedHost) *pinnedHost = z.hostBlobWorst + z.wireBlob + CTR_COUNT * sizeof(uint32_t);
    return total;
}

int marchingAlloc(mlsb_worker *w)
{
    const mlsb_worker_config &c = w->cfg;
    buildTables();
    const MarchSizes z = marchSizes(c);
    w->fieldDim = z.fieldDim;
    w->vertexSpace = z.vertexSpace;
    w->indexSpace = z.indexSpace;
    w->latticeWords = z.latticeWords;
    w->latStatusWords = z.latStatusWords;
    w->maxTiles = z.maxTiles;
    MLSB_REQUIRE(w->latticeWords < (1ull << 32) - kLatThreads * kLatItems, MLSB_ERR_LENGTH, "lattice too large");
    MLSB_CUDA(cudaMalloc(&w->field, z.field));
    MLSB_CUDA(cudaMalloc(&w->blockActive, z.blockActive));
    MLSB_CUDA(cudaMemset(w->blockActive, 0, z.blockActive));
    MLSB_CUDA(cudaMalloc(&w->latMask, z.latMask));
    MLSB_CUDA(cudaMalloc(&w->latRecord, z.latRecord));
    MLSB_CUDA(cudaMalloc(&w->indices, z.indices));
    MLSB_CUDA(cudaMalloc(&w->welded, z.welded));
    MLSB_CUDA(cudaMalloc(&w->weldedKeys, z.weldedKeys));
    MLSB_CUDA(cudaMalloc(&w->weldedFlagged, z.welded));
    MLSB_CUDA(cudaMalloc(&w->weldedKeysFlagged, z.weldedKeys));
    // [counters | per-chunk counts of the lattice scan]
    MLSB_CUDA(cudaMalloc(&w->marchCounters, z.counters));
    w->latStatus = reinterpret_cast<uint64_t *>(w->marchCounters + CTR_COUNT);
    MLSB_CUDA(cudaMalloc(&w->tileTotals, z.tileTotals));
    MLSB_CUDA(cudaMalloc(&w->tileList, z.tileList));
    MLSB_CUDA(cudaMalloc(&w->tables, z.tables));
    if (z.wireBlob)
    {
        MLSB_CUDA(cudaMalloc(&w->compactVerts, z.compactVerts));
        MLSB_CUDA(cudaMalloc(&w->compactVertsFlagged, z.compactVerts));
        MLSB_CUDA(cudaMalloc(&w->wireBlob, z.wireBlob));
        w->wireWords = z.wireBlob / sizeof(uint64_t);
        w->compactReadback = true;
    }
    MLSB_CUDA(cudaMemcpy(w->tables, &gTables.t, sizeof(Tables), cudaMemcpyHostToDevice));
    MLSB_TRY(ensureHostBlob(w, (size_t) 8 << 20));
    MLSB_CUDA(cudaMallocHost(&w->hostCounters, CTR_COUNT * sizeof(uint32_t)));
    MLSB_CUDA(cudaFuncSetAttribute(generateCellsKernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (2 * kGenRows * 2 + (int) kMaxCellVertices) * kGenThreads * (int) sizeof(uint32_t)));
    return MLSB_OK;
}

void marchingFree(mlsb_worker *w)
{
    cudaFree(w->field);
    cudaFree(w->blockActive);
    cudaFree(w->latMask);
    cudaFree(w->latRecord);
    cudaFree(w->indices);
    cudaFree(w->welded);
    cudaFree(w->weldedKeys);
    cudaFree(w->weldedFlagged);
    cudaFree(w->weldedKeysFlagged);
    cudaFree(w->marchCounters);
    cudaFree(w->tileTotals);
    cudaFree(w->tileList);
    cudaFree(w->tables);
    cudaFree(w->compactVerts);
    cudaFree(w->compactVertsFlagged);
    cudaFree(w->wireBlob);
    cudaFreeHost(w->hostWire);
    cudaFreeHost(w->hostBlob);
    cudaFreeHost(w->hostCounters);
}

namespace
{

}  // namespace

// MeshFilterChain::operator() (src/mesh_filter.cpp:44-64) without the output functor: `mesh` becomes the last
// filter's output.  Filters enqueue on the worker's stream, so they are ordered by it.
int runMeshFilters(mlsb_worker *w, mlsb_device_mesh &mesh)
{
    for (const std::pair<mlsb_mesh_filter, void *> &f : w->meshFilters)
    {
        mlsb_device_mesh out = mesh;
        const int rc = f.first(f.second, &mesh, &out, (void *) w->stream);
        if (rc != 0)
        {
            waitStream(w);
            setError("a mesh filter returned %d", rc);
            return MLSB_ERR_STATE;
        }
        MLSB_REQUIRE(out.sizes.numInternalVertices <= out.sizes.numVertices, MLSB_ERR_STATE,
                     "a mesh filter produced more internal vertices than vertices");
        mesh = out;
    }
    return MLSB_OK;
}

namespace
{

// One shipped-out block: the cells between corner slices zFirst..zLast (Marching::addSlices +
// shipOut with zTop = zFirst, zMax = zLast).  Returns MLSB_ERR_OVERFLOW if the block does not
// fit in the worker's mesh memory (the caller then splits the z range, as
// src/marching.cpp:651-704 does with its per-slice histogram).
int shipBlock(mlsb_worker *w, bool useActive, uint32_t pitch, uint32_t zStride, const uint32_t size[3], uint32_t zFirst,
              uint32_t zLast, const uint32_t keyOffset[3], const MarchIo &io)
{
    cudaStream_t s = w->stream;
    const uint32_t cellsX = size[0] - 1, cellsY = size[1] - 1;
    const uint64_t numCells64 = (uint64_t) cellsX * cellsY * (zLast - zFirst);
    if (numCells64 == 0) return MLSB_OK;
    MLSB_REQUIRE(numCells64 < (1ull << 32) - kGenThreads, MLSB_ERR_LENGTH, "too many cells");
    const uint32_t tiles = (zLast - zFirst) * divUp(cellsY, (uint32_t) kGenRows);
    const uint32_t chunks = divUp(cellsX, (uint32_t) kGenThreads);
    MLSB_REQUIRE(chunks <= 2, MLSB_ERR_LENGTH, "bin too wide");
    const size_t genSmem = (2 * (size_t) kGenRows * chunks + kMaxCellVertices) * kGenThreads * sizeof(uint32_t);
    const uint32_t latStride = 2 * divUp(cellsX, 32u) + 1, latRows = 2 * cellsY + 1, latPlanes = 2 * (zLast - zFirst) + 1;
    const size_t latWords = (size_t) latPlanes * latRows * latStride;
    const uint32_t latTiles = (uint32_t) divUp(latWords, (size_t) (kLatThreads * kLatItems));
    MLSB_REQUIRE(latWords <= w->latticeWords && latTiles <= w->latStatusWords && tiles <= w->maxTiles, MLSB_ERR_LENGTH,
                 "scan scratch too small");

    GenParams G;
    G.field = w->field;
    G.pitch = pitch;
    G.zStride = zStride;
    G.cellsX = cellsX;
    G.cellsY = cellsY;
    G.zFirst = zFirst;
    G.indices = w->indices;
    G.counters = w->marchCounters;
    G.latMask = w->latMask;
    G.latRecord = w->latRecord;
    G.latStride = latStride;
    G.latRows = latRows;
    G.tileTotals = w->tileTotals;
    G.blockActive = useActive ? w->blockActive : nullptr;
    G.activeDim = roundUp(w->fieldDim / kMlsWgs, 4u);
    G.tileList = w->tileList;
    G.tables = reinterpret_cast<const Tables *>(w->tables);

    LatticeParams L;
    L.mask = w->latMask;
    L.record = w->latRecord;
    L.stride = latStride;
    L.rows = latRows;
    L.planes = latPlanes;
    L.totalWords = (uint32_t) latWords;
    L.topX = 2 * cellsX;
    L.topY = 2 * cellsY;
    L.counters = w->marchCounters;
    L.status = w->latStatus;
    L.field = w->field;
    L.pitch = pitch;
    L.zStride = zStride;
    L.zFirst = zFirst;
    L.gridOffset = make_uint3(keyOffset[0], keyOffset[1], keyOffset[2]);
    L.welded = w->welded;
    L.weldedKeys = w->weldedKeys;
    L.weldedFlagged = w->weldedFlagged;
    L.weldedKeysFlagged = w->weldedKeysFlagged;
    L.compact = w->compactVerts;          // null unless MLSB_READBACK=compact
    L.compactFlagged = w->compactVertsFlagged;
    L.keyOffset = ((uint64_t) keyOffset[2] << (2 * kKeyAxisBits + 1)) | ((uint64_t) keyOffset[1] << (kKeyAxisBits + 1))
                  | ((uint64_t) keyOffset[0] << 1);
    L.scale = w->scale;
    L.bx = w->bias[0];
    L.by = w->bias[1];
    L.bz = w->bias[2];
    L.vertexSpace = (uint32_t) w->vertexSpace;
    L.indexSpace = (uint32_t) w->indexSpace;
    L.tileTotals = w->tileTotals;
    L.tileList = w->tileList;
    L.numTiles = tiles;

    {
        ScopedStat st(w, "kernel.marching.zero.time", 0);
        // (writing the mask also leaves it in L2, where the classification's red.or find it: against a mask
        // that cleans itself and is never memset the classify kernel ran 50 % longer on DRAM-resident lines)
        MLSB_CUDA(cudaMemsetAsync(w->marchCounters, 0, CTR_COUNT * sizeof(uint32_t), s));
        MLSB_CUDA(cudaMemsetAsync(w->latMask, 0, latWords * sizeof(uint32_t), s));
    }
    {
        ScopedStat st(w, "kernel.marching.genOccupied.time");
        if (chunks == 1) generateCellsKernel<false, 1><<<tiles, kGenThreads, 0, s>>>(G);
        else generateCellsKernel<false, 2><<<tiles, kGenThreads, 0, s>>>(G);
        w->launches++;
    }
    {
        // + 1: the CTA that scans the tile totals (kernel.marching.scanElements in the reference's vocabulary)
        ScopedStat st(w, "kernel.marching.countUniqueVertices.time");
        latticeCountKernel<<<latTiles + 1, kLatThreads, 0, s>>>(L);
        w->launches++;
    }
    {
        ScopedStat st(w, "kernel.marching.compactVertices.time");
        latticeEmitKernel<<<latTiles, kLatThreads, 0, s>>>(L, latTiles);
        w->launches++;
    }
    {
        ScopedStat st(w, "kernel.marching.generateElements.time");
        if (chunks == 1) generateCellsKernel<true, 1><<<tiles, kGenThreads, genSmem, s>>>(G);
        else generateCellsKernel<true, 2><<<tiles, kGenThreads, genSmem, s>>>(G);
        w->launches++;
    }
    MLSB_CUDA(cudaGetLastError());

    // ---- sizes and mesh to the host: the pack kernel writes both into pinned memory (one wait); the device-mesh
    // variant only needs the sizes (src/marching.cpp:587-617) ----
    auto pack = [&]() {
        ScopedStat st(w, "kernel.marching.readback.time");
        packMeshKernel<<<kPackBlocks, kPackThreads, 0, s>>>(w->marchCounters, w->welded, w->weldedFlagged, w->weldedKeys,
                                                            w->weldedKeysFlagged, w->indices, static_cast<uint32_t *>(w->hostBlob),
                                                            w->hostBlobBytes, w->hostCounters, (uint32_t) w->vertexSpace,
                                                            (uint32_t) w->indexSpace);
        w->launches++;
    };
    const bool packed = !io.deviceSink && w->meshFilters.empty() && w->packReadback;
    const bool compact = !io.deviceSink && w->meshFilters.empty() && w->compactReadback;
    if (packed)
        pack();
    else
    {
        if (compact)
        {
            // the wire blob is gathered while the sizes travel: the host learns from CTR_PACKED whether to fetch it
            ScopedStat st(w, "kernel.marching.readback.time");
            compactMeshKernel<<<296, 256, 0, s>>>(w->marchCounters, w->weldedKeys, w->weldedKeysFlagged, w->compactVerts,
                                                  w->compactVertsFlagged, w->indices, w->wireBlob, w->wireWords,
                                                  (uint32_t) w->vertexSpace, (uint32_t) w->indexSpace);
            w->launches++;
        }
        ScopedStat st(w, "kernel.marching.readback.time", 0);
        MLSB_CUDA(cudaMemcpyAsync(w->hostCounters, w->marchCounters, CTR_COUNT * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    }
    MLSB_CUDA(cudaGetLastError());
    MLSB_TRY(waitStream(w));
    const uint32_t *hc = w->hostCounters;
    w->marchWork[0] += (uint64_t) size[0] * size[1] * (zLast - zFirst + 1);
    if (hc[CTR_OVERFLOW] || hc[CTR_VERTS] > w->vertexSpace || hc[CTR_INDICES] > w->indexSpace) return MLSB_ERR_OVERFLOW;
    if (hc[CTR_VERTS] == 0) return MLSB_OK;

    w->shipouts++;
    w->marchWork[1] += hc[CTR_CELLS];
    w->marchWork[2] += hc[CTR_VERTS];
    w->marchWork[3] += hc[CTR_INDICES];
    w->marchWork[4] += hc[CTR_WELDED];
    mlsb_mesh_sizes sizes{hc[CTR_WELDED], hc[CTR_INDICES] / 3, hc[CTR_FIRST_EXTERNAL]};
    const uint64_t numExternal = sizes.numVertices - sizes.numInternalVertices;
    // unflagged vertices (internal, then the top plane) and flagged ones live in two buffers
    const uint64_t numUnflagged = hc[CTR_UNFLAGGED], numFlagged = sizes.numVertices - numUnflagged;
    const uint64_t numTop = numUnflagged - sizes.numInternalVertices;
    if (io.deviceSink || !w->meshFilters.empty())
    {
        // DeviceKeyMesh (src/mesh.h:101-123): one vertex and one key array, internal vertices first.  The flagged
        // vertices (and their keys) move in behind the unflagged ones, device to device.
        if (numFlagged)
        {
            MLSB_CUDA(cudaMemcpyAsync(w->welded + 3 * numUnflagged, w->weldedFlagged, numFlagged * 3 * sizeof(float),
                                      cudaMemcpyDeviceToDevice, s));
            MLSB_CUDA(cudaMemcpyAsync(w->weldedKeys + numUnflagged, w->weldedKeysFlagged, numFlagged * sizeof(uint64_t),
                                      cudaMemcpyDeviceToDevice, s));
        }
        mlsb_device_mesh mesh{sizes, w->welded, w->weldedKeys, w->indices};
        // MeshFilterChain::operator() (src/mesh_filter.cpp:44-64): each filter reads the previous one's output;
        // the scale-bias step of the reference's chain has already happened in the vertex pass
        MLSB_TRY(runMeshFilters(w, mesh));
        if (io.deviceSink)
        {
            const int rc = io.deviceSink(&mesh, (void *) s, io.deviceUser);
            MLSB_TRY(waitStream(w));
            if (rc != 0)
            {
                setError("the device mesh sink returned %d", rc);
                return MLSB_ERR_STATE;
            }
            return MLSB_OK;
        }
        MLSB_TRY(ensureHostBlob(w, (size_t) mlsb_mesh_host_bytes(&mesh.sizes)));
        MLSB_TRY(marchingReadDeviceMesh(w, &mesh, w->hostBlob));
        if (io.sink) io.sink(&mesh.sizes, w->hostBlob, io.user);
        return MLSB_OK;
    }
    // HostKeyMesh layout, src/mesh.cpp:51-60
    MLSB_TRY(ensureHostBlob(w, (size_t) mlsb_mesh_host_bytes(&sizes)));   // keeps the pointer if it is large enough
    if (packed)
    {
        // already in the blob unless it was too small
        if (!hc[CTR_PACKED])
        {
            pack();
            MLSB_CUDA(cudaGetLastError());
            MLSB_TRY(waitStream(w));
            MLSB_REQUIRE(w->hostCounters[CTR_PACKED], MLSB_ERR_STATE, "the mesh did not fit the grown read-back blob");
        }
    }
    else if (compact && hc[CTR_PACKED])
    {
        // one copy of 8 bytes per key, vertex and triangle; the worker thread expands it (unpack.cpp)
        const size_t wireBytes = (size_t) (numExternal + sizes.numVertices + sizes.numTriangles) * sizeof(uint64_t);
        if (wireBytes > w->hostWireBytes)
        {
            if (w->hostWire) MLSB_CUDA(cudaFreeHost(w->hostWire));
            w->hostWire = nullptr;
            w->hostWireBytes = 0;
            const size_t want = std::max(wireBytes + wireBytes / 2, (size_t) 4 << 20);
            MLSB_CUDA(cudaMallocHost(&w->hostWire, want));
            w->hostWireBytes = want;
        }
        MLSB_CUDA(cudaMemcpyAsync(w->hostWire, w->wireBlob, wireBytes, cudaMemcpyDeviceToHost, s));
        MLSB_TRY(waitStream(w));
        addSample(w, "readback.bytes", (double) wireBytes);
        WireParams wp;
        wp.numVertices = sizes.numVertices;
        wp.numTriangles = sizes.numTriangles;
        wp.numInternalVertices = sizes.numInternalVertices;
        wp.zFirst = zFirst;
        for (int a = 0; a < 3; a++) wp.gridOffset[a] = keyOffset[a], wp.bias[a] = w->bias[a];
        wp.scale = w->scale;
        unpackWireMesh(static_cast<const uint64_t *>(w->hostWire), wp, w->hostBlob);
    }
    else
    {
        // enqueueReadMesh, src/mesh.cpp:62-102, on the copy engine: no SM is involved
        uint8_t *hKeys = static_cast<uint8_t *>(w->hostBlob);
        uint8_t *hVerts = hKeys + numExternal * sizeof(uint64_t);
        uint8_t *hTris = hVerts + sizes.numVertices * 3 * sizeof(float);
        MLSB_CUDA(cudaMemcpyAsync(hTris, w->indices, sizes.numTriangles * 3 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        if (numTop)
            MLSB_CUDA(cudaMemcpyAsync(hKeys, w->weldedKeys + sizes.numInternalVertices, numTop * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
        if (numFlagged)
        {
            MLSB_CUDA(cudaMemcpyAsync(hKeys + numTop * sizeof(uint64_t), w->weldedKeysFlagged, numFlagged * sizeof(uint64_t),
                                      cudaMemcpyDeviceToHost, s));
            MLSB_CUDA(cudaMemcpyAsync(hVerts + numUnflagged * 3 * sizeof(float), w->weldedFlagged, numFlagged * 3 * sizeof(float),
                                      cudaMemcpyDeviceToHost, s));
        }
        if (numUnflagged)
            MLSB_CUDA(cudaMemcpyAsync(hVerts, w->welded, numUnflagged * 3 * sizeof(float), cudaMemcpyDeviceToHost, s));
        MLSB_TRY(waitStream(w));
    }
    // what crossed PCIe for this block (statistic readback.bytes): the blob, unless the compact wire format was used
    if (!(compact && hc[CTR_PACKED])) addSample(w, "readback.bytes", (double) mlsb_mesh_host_bytes(&sizes));
    const uint8_t *blob = static_cast<const uint8_t *>(w->hostBlob);
    if (io.sink) io.sink(&sizes, blob, io.user);
    return MLSB_OK;
}

int shipRange(mlsb_worker *w, bool useActive, uint32_t pitch, uint32_t zStride, const uint32_t size[3], uint32_t zFirst,
              uint32_t zLast, const uint32_t keyOffset[3], const MarchIo &io)
{
    int st = shipBlock(w, useActive, pitch, zStride, size, zFirst, zLast, keyOffset, io);
    if (st != MLSB_ERR_OVERFLOW) return st;
    addSample(w, "marching.overflow", 1.0);
    if (zLast - zFirst <= 1)
    {
        setError("a single slice of cells does not fit in meshBytes");
        return MLSB_ERR_OVERFLOW;
    }
    // Too big for the mesh memory: split the slice range and ship the halves separately; the
    // seam is welded by the consumer through the external vertex keys.
    uint32_t mid = zFirst + (zLast - zFirst) / 2;
    MLSB_TRY(shipRange(w, useActive, pitch, zStride, size, zFirst, mid, keyOffset, io));
    return shipRange(w, useActive, pitch, zStride, size, mid, zLast, keyOffset, io);
}

}  // namespace

int marchingGenerate(mlsb_worker *w, const MarchIo &io, const uint32_t size[3], const uint32_t keyOffset[3])
{
    // src/marching.cpp:767-769
    for (int i = 0; i < 3; i++)
        MLSB_REQUIRE(size[i] >= 1 && size[i] <= w->cfg.maxCells + 1, MLSB_ERR_LENGTH, "size exceeds the worker's maximum");
    for (int i = 0; i < 3; i++)
        MLSB_REQUIRE((uint64_t) keyOffset[i] + size[i] <= (1u << (kKeyAxisBits - 1)) - 1, MLSB_ERR_LENGTH,
                     "global coordinates exceed MAX_GLOBAL_DIMENSION");
    const uint32_t pitch = roundUp(size[0], (uint32_t) kMlsWgs);
    const uint32_t zStride = roundUp(size[1], (uint32_t) kMlsWgs);
    const mlsb_swathe sw{size[0], size[1], zStride, 0, 0, size[2] - 1};
    if (io.useMls)
        MLSB_TRY(mlsCorners(w, w->field, pitch, sw, true));
    else if (io.generate)
    {
        // Marching::Generator::enqueue (src/marching.h:229-252): one swathe, the whole bin
        const int rc = io.generate(io.generateUser, w->field, pitch, &sw, (void *) w->stream);
        if (rc != 0)
        {
            setError("the field generator returned %d", rc);
            return MLSB_ERR_STATE;
        }
    }
    else
    {
        MLSB_REQUIRE(io.hostField != nullptr, MLSB_ERR_INVALID_ARGUMENT, "null field");
        for (uint32_t z = 0; z < size[2]; z++)
            MLSB_CUDA(cudaMemcpy2DAsync(w->field + (size_t) z * zStride * pitch, pitch * sizeof(float),
                                        io.hostField + (size_t) z * size[1] * size[0], size[0] * sizeof(float),
                                        size[0] * sizeof(float), size[1], cudaMemcpyHostToDevice, w->stream));
    }
    if (size[0] < 2 || size[1] < 2 || size[2] < 2)
    {
        MLSB_TRY(waitStream(w));
        return MLSB_OK;
    }
    w->shipouts = 0;
    const int status = shipRange(w, io.useMls, pitch, zStride, size, 0, size[2] - 1, keyOffset, io);
    if (status == MLSB_OK) addSample(w, "marching.shipouts", (double) w->shipouts);
    return status;
}

// enqueueReadMesh, src/mesh.cpp:62-102
int marchingReadDeviceMesh(mlsb_worker *w, const mlsb_device_mesh *mesh, void *hostBlob)
{
    const mlsb_mesh_sizes &sz = mesh->sizes;
    const uint64_t numExternal = sz.numVertices - sz.numInternalVertices;
    uint8_t *hKeys = static_cast<uint8_t *>(hostBlob);
    uint8_t *hVerts = hKeys + numExternal * sizeof(uint64_t);
    uint8_t *hTris = hVerts + sz.numVertices * 3 * sizeof(float);
    cudaStream_t s = w->stream;
    if (numExternal)
        MLSB_CUDA(cudaMemcpyAsync(hKeys, mesh->vertexKeys + sz.numInternalVertices, numExternal * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    if (sz.numVertices)
        MLSB_CUDA(cudaMemcpyAsync(hVerts, mesh->vertices, sz.numVertices * 3 * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (sz.numTriangles)
        MLSB_CUDA(cudaMemcpyAsync(hTris, mesh->triangles, sz.numTriangles * 3 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    return waitStream(w);
}

}  // namespace mlsb
