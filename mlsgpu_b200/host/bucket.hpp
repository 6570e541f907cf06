// Bounding grid, world -> grid transform and spatial binning in front of the device path.
//
// Grid: SplatSet::FastBlobSet::getBoundingGrid (src/splat_set_impl.h:771-810): reference point 0,
// spacing = --fit-grid, low extent rounded down to a multiple of the blob bucket size.
// Transform: Grid::worldToVertex + radius rescale (src/grid.cpp:99-106, src/bucket_loader.cpp:76-85).
// Bins: the same contract as Bucket::bucket (src/bucket_impl.h:439-561) — at most maxCells cells
// per side, at most maxSplats splats, every splat whose ball reaches the bin is included — but
// found by cutting the grid into near-cubic blocks and halving any block that is still too full,
// instead of the reference's histogram-driven recursion over blobs.
#ifndef MLSB200_BUCKET_HPP
#define MLSB200_BUCKET_HPP

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <omp.h>
#include <stdexcept>
#include <vector>

#include "../../include/mlsb200.h"

namespace mlsb
{

struct Grid
{
    float spacing = 1.0f;
    std::int64_t lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};   // vertex-index extents, reference point = origin
    std::int64_t numCells(int a) const { return hi[a] - lo[a]; }
    void getVertex0(float out[3]) const   // Grid::getVertex(0,0,0), src/grid.cpp:92-97
    {
        for (int a = 0; a < 3; a++) out[a] = 0.0f + spacing * (float) lo[a];
    }
};

struct Bin
{
    std::int32_t offset[3];
    std::uint32_t size[3];              // corners per axis
    std::uint32_t chunk[3] = {0, 0, 0}; // output chunk the bin belongs to (--split; ChunkId::coords, src/chunk_id.h)
    std::vector<std::uint32_t> ids;     // ascending indices into the grid-space splat array
};

inline Grid boundingGrid(const std::vector<mlsb_splat> &splats, float spacing, std::int64_t bucketSize)
{
    if (splats.empty()) throw std::runtime_error("no splats");
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
#pragma omp parallel
    {
        double tmn[3] = {1e300, 1e300, 1e300}, tmx[3] = {-1e300, -1e300, -1e300};
#pragma omp for schedule(static) nowait
        for (std::int64_t i = 0; i < (std::int64_t) splats.size(); i++)
        {
            const mlsb_splat &s = splats[(std::size_t) i];
            for (int a = 0; a < 3; a++)
            {
                tmn[a] = std::min(tmn[a], (double) s.position[a] - s.radius);
                tmx[a] = std::max(tmx[a], (double) s.position[a] + s.radius);
            }
        }
#pragma omp critical(mlsb_bounding_grid)
        for (int a = 0; a < 3; a++)
        {
            mn[a] = std::min(mn[a], tmn[a]);
            mx[a] = std::max(mx[a], tmx[a]);
        }
    }
    Grid g;
    g.spacing = spacing;
    for (int a = 0; a < 3; a++)
    {
        std::int64_t lo = (std::int64_t) std::floor(mn[a] / spacing);
        lo = (lo >= 0 ? lo / bucketSize : -((-lo + bucketSize - 1) / bucketSize)) * bucketSize;
        g.lo[a] = lo;
        g.hi[a] = (std::int64_t) std::ceil(mx[a] / spacing);
    }
    return g;
}

inline void worldToGrid(std::vector<mlsb_splat> &splats, const Grid &g)
{
    const float inv = 1.0f / g.spacing;
#pragma omp parallel for schedule(static)
    for (std::int64_t i = 0; i < (std::int64_t) splats.size(); i++)
    {
        mlsb_splat &s = splats[(std::size_t) i];
        for (int a = 0; a < 3; a++) s.position[a] = (s.position[a] - 0.0f) * inv - (float) g.lo[a];
        s.radius *= inv;
    }
}

namespace detail
{
inline void split(const std::vector<mlsb_splat> &splats, const std::int64_t lo[3], const std::int64_t hi[3],
                  const std::vector<std::uint32_t> &candidates, std::int64_t maxCells, std::size_t maxSplats,
                  const std::uint32_t chunk[3], std::vector<Bin> &bins)
{
    std::vector<std::uint32_t> ids;
    for (std::uint32_t id : candidates)
    {
        const mlsb_splat &s = splats[id];
        bool in = true;
        for (int a = 0; a < 3; a++)
            in = in && s.position[a] + s.radius >= (float) lo[a] && s.position[a] - s.radius <= (float) hi[a];
        if (in) ids.push_back(id);
    }
    if (ids.empty()) return;
    std::int64_t ext[3] = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
    const int axis = (int) (std::max_element(ext, ext + 3) - ext);
    if (ext[axis] <= maxCells && ids.size() <= maxSplats)
    {
        Bin b;
        for (int a = 0; a < 3; a++)
        {
            b.offset[a] = (std::int32_t) lo[a];
            b.size[a] = (std::uint32_t) (ext[a] + 1);
            b.chunk[a] = chunk[a];
        }
        b.ids.swap(ids);
        bins.push_back(std::move(b));
        return;
    }
    if (ext[axis] <= 1) throw std::runtime_error("too many splats over a single cell (DensityError upstream)");
    std::int64_t hi0[3] = {hi[0], hi[1], hi[2]}, lo1[3] = {lo[0], lo[1], lo[2]};
    hi0[axis] = lo1[axis] = lo[axis] + ext[axis] / 2;
    split(splats, lo, hi0, ids, maxCells, maxSplats, chunk, bins);
    split(splats, lo1, hi, ids, maxCells, maxSplats, chunk, bins);
}
}  // namespace detail

/// The first cut of the grid into blocks of at most maxCells cells per side (and, with --split, never
/// across an output chunk), shared by the in-core binning below and the streaming front-end
/// (stream.hpp), which must agree on it block for block.
struct BlockLayout
{
    std::int64_t n[3], counts[3];
    std::vector<float> edges[3];             // block b of axis a spans cells [edges[a][b], edges[a][b + 1]]
    std::vector<std::uint32_t> chunkOf[3];   // chunk coordinate of block b

    BlockLayout(const Grid &g, std::int64_t maxCells, std::int64_t chunkCells)
    {
        for (int a = 0; a < 3; a++)
        {
            n[a] = std::max<std::int64_t>(g.numCells(a), 1);
            const std::int64_t chunkLen = chunkCells > 0 ? chunkCells : n[a];
            for (std::int64_t c0 = 0, chunk = 0; c0 < n[a]; c0 += chunkLen, chunk++)
            {
                // a chunk is cut into equal blocks of at most maxCells cells
                const std::int64_t len = std::min(chunkLen, n[a] - c0), parts = (len + maxCells - 1) / maxCells;
                for (std::int64_t b = 0; b < parts; b++)
                {
                    edges[a].push_back((float) (c0 + b * len / parts));
                    chunkOf[a].push_back((std::uint32_t) chunk);
                }
            }
            edges[a].push_back((float) n[a]);
            counts[a] = (std::int64_t) chunkOf[a].size();
        }
    }

    std::size_t numBlocks() const { return (std::size_t) (counts[0] * counts[1] * counts[2]); }
    std::size_t index(std::int64_t bx, std::int64_t by, std::int64_t bz) const { return (std::size_t) ((bz * counts[1] + by) * counts[0] + bx); }

    /// The blocks a splat's ball reaches along axis a (closed-interval test); false if none.
    bool range(const mlsb_splat &s, int a, std::int64_t &first, std::int64_t &last) const
    {
        const float l = s.position[a] - s.radius, h = s.position[a] + s.radius;
        const std::vector<float> &e = edges[a];
        // blocks b with h >= e[b] and l <= e[b + 1]
        first = std::lower_bound(e.begin() + 1, e.end(), l) - (e.begin() + 1);      // first b with e[b + 1] >= l
        last = (std::upper_bound(e.begin(), e.end() - 1, h) - e.begin()) - 1;      // last b with e[b] <= h
        return first <= last && first < counts[a] && last >= 0;
    }
};

/// splats: grid coordinates (after worldToGrid).  Bins come out in z-major block order.
/// chunkCells > 0 (--split): the grid is first cut into output chunks of that many cells per side,
/// counted from the grid's low corner, and no bin crosses a chunk boundary (Bucket::bucket's
/// chunkCells, src/bucket.h); every bin records its chunk.
inline std::vector<Bin> makeBins(const std::vector<mlsb_splat> &splats, const Grid &g, std::int64_t maxCells, std::size_t maxSplats,
                                 std::int64_t chunkCells = 0)
{
    const BlockLayout layout(g, maxCells, chunkCells);
    const std::int64_t *counts = layout.counts;
    const std::vector<float> *edges = layout.edges;
    const std::vector<std::uint32_t> *chunkOf = layout.chunkOf;
    const std::size_t numBlocks = layout.numBlocks();
    const std::int64_t numSplats = (std::int64_t) splats.size();
    auto range = [&](const mlsb_splat &s, int a, std::int64_t &first, std::int64_t &last) { return layout.range(s, a, first, last); };

    // Two passes over the splats with the same static schedule: count per (thread, block), then
    // fill exact-size lists.  Each thread owns a contiguous range of splats and the threads'
    // segments of a block are laid out in thread order, so the ids of a block come out ascending.
    std::vector<std::vector<std::uint32_t>> blocks(numBlocks);
    std::vector<std::vector<std::size_t>> cursor;
#pragma omp parallel
    {
        const int numThreads = omp_get_num_threads(), me = omp_get_thread_num();
#pragma omp single
        cursor.assign((std::size_t) numThreads, std::vector<std::size_t>(numBlocks, 0));
        const std::int64_t begin = numSplats * me / numThreads, end = numSplats * (me + 1) / numThreads;
        std::vector<std::size_t> &mine = cursor[(std::size_t) me];
        for (int pass = 0; pass < 2; pass++)
        {
            for (std::int64_t id = begin; id < end; id++)
            {
                const mlsb_splat &s = splats[(std::size_t) id];
                std::int64_t first[3], last[3];
                if (!(range(s, 0, first[0], last[0]) && range(s, 1, first[1], last[1]) && range(s, 2, first[2], last[2])))
                    continue;
                for (std::int64_t bz = first[2]; bz <= last[2]; bz++)
                    for (std::int64_t by = first[1]; by <= last[1]; by++)
                        for (std::int64_t bx = first[0]; bx <= last[0]; bx++)
                        {
                            const std::size_t b = (std::size_t) ((bz * counts[1] + by) * counts[0] + bx);
                            if (pass == 0)
                                mine[b]++;
                            else
                                blocks[b][mine[b]++] = (std::uint32_t) id;
                        }
            }
            if (pass == 0)
            {
#pragma omp barrier
#pragma omp for schedule(static)
                for (std::int64_t b = 0; b < (std::int64_t) numBlocks; b++)
                {
                    std::size_t total = 0;
                    for (int t = 0; t < numThreads; t++)
                    {
                        const std::size_t c = cursor[(std::size_t) t][(std::size_t) b];
                        cursor[(std::size_t) t][(std::size_t) b] = total;
                        total += c;
                    }
                    blocks[(std::size_t) b].resize(total);
                }   // implicit barrier
            }
        }
    }
    // blocks are independent: a block that fits becomes a bin as it is, the rest are halved
    std::vector<std::vector<Bin>> perBlock(numBlocks);
#pragma omp parallel for schedule(dynamic, 1)
    for (std::int64_t bi = 0; bi < (std::int64_t) numBlocks; bi++)
    {
        std::vector<std::uint32_t> &ids = blocks[(std::size_t) bi];
        if (ids.empty()) continue;
        const std::int64_t bx = bi % counts[0], by = (bi / counts[0]) % counts[1], bz = bi / (counts[0] * counts[1]);
        const std::int64_t lo[3] = {(std::int64_t) edges[0][bx], (std::int64_t) edges[1][by], (std::int64_t) edges[2][bz]};
        const std::int64_t hi[3] = {(std::int64_t) edges[0][bx + 1], (std::int64_t) edges[1][by + 1], (std::int64_t) edges[2][bz + 1]};
        const std::uint32_t chunk[3] = {chunkOf[0][(std::size_t) bx], chunkOf[1][(std::size_t) by], chunkOf[2][(std::size_t) bz]};
        if (ids.size() <= maxSplats)   // the block's extent is within maxCells by construction
        {
            Bin b;
            for (int a = 0; a < 3; a++)
            {
                b.offset[a] = (std::int32_t) lo[a];
                b.size[a] = (std::uint32_t) (hi[a] - lo[a] + 1);
                b.chunk[a] = chunk[a];
            }
            b.ids.swap(ids);
            perBlock[(std::size_t) bi].push_back(std::move(b));
        }
        else
            detail::split(splats, lo, hi, ids, maxCells, maxSplats, chunk, perBlock[(std::size_t) bi]);
    }
    std::vector<Bin> bins;
    for (std::vector<Bin> &v : perBlock)
        for (Bin &b : v) bins.push_back(std::move(b));
    return bins;
}

}  // namespace mlsb

#endif
