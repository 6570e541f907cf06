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
    std::vector<std::uint32_t> ids;     // ascending indices into the grid-space splat array
};

inline Grid boundingGrid(const std::vector<mlsb_splat> &splats, float spacing, std::int64_t bucketSize)
{
    if (splats.empty()) throw std::runtime_error("no splats");
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
    for (const mlsb_splat &s : splats)
        for (int a = 0; a < 3; a++)
        {
            mn[a] = std::min(mn[a], (double) s.position[a] - s.radius);
            mx[a] = std::max(mx[a], (double) s.position[a] + s.radius);
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
    for (mlsb_splat &s : splats)
    {
        for (int a = 0; a < 3; a++) s.position[a] = (s.position[a] - 0.0f) * inv - (float) g.lo[a];
        s.radius *= inv;
    }
}

namespace detail
{
inline void split(const std::vector<mlsb_splat> &splats, const std::int64_t lo[3], const std::int64_t hi[3],
                  const std::vector<std::uint32_t> &candidates, std::int64_t maxCells, std::size_t maxSplats,
                  std::vector<Bin> &bins)
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
        }
        b.ids.swap(ids);
        bins.push_back(std::move(b));
        return;
    }
    if (ext[axis] <= 1) throw std::runtime_error("too many splats over a single cell (DensityError upstream)");
    std::int64_t hi0[3] = {hi[0], hi[1], hi[2]}, lo1[3] = {lo[0], lo[1], lo[2]};
    hi0[axis] = lo1[axis] = lo[axis] + ext[axis] / 2;
    split(splats, lo, hi0, ids, maxCells, maxSplats, bins);
    split(splats, lo1, hi, ids, maxCells, maxSplats, bins);
}
}  // namespace detail

/// splats: grid coordinates (after worldToGrid).  Bins come out in z-major block order.
inline std::vector<Bin> makeBins(const std::vector<mlsb_splat> &splats, const Grid &g, std::int64_t maxCells, std::size_t maxSplats)
{
    std::vector<Bin> bins;
    std::int64_t n[3], counts[3];
    for (int a = 0; a < 3; a++)
    {
        n[a] = std::max<std::int64_t>(g.numCells(a), 1);
        counts[a] = (n[a] + maxCells - 1) / maxCells;
    }
    // candidate lists per block from each splat's bounding box in block units; every thread takes
    // a contiguous range of splats, so concatenating the per-thread lists keeps ids ascending
    const std::size_t numBlocks = (std::size_t) (counts[0] * counts[1] * counts[2]);
    std::vector<std::vector<std::uint32_t>> blocks(numBlocks);
    auto edge = [&](int a, std::int64_t b) { return b * n[a] / counts[a]; };
#pragma omp parallel
    {
        std::vector<std::vector<std::uint32_t>> mine(numBlocks);
#pragma omp for schedule(static)
        for (std::int64_t id = 0; id < (std::int64_t) splats.size(); id++)
        {
            const mlsb_splat &s = splats[(std::size_t) id];
            std::int64_t first[3], last[3];
            for (int a = 0; a < 3; a++)
            {
                const double l = (double) s.position[a] - s.radius, h = (double) s.position[a] + s.radius;
                std::int64_t f = (std::int64_t) std::floor(l * counts[a] / n[a]) - 1, t = (std::int64_t) std::floor(h * counts[a] / n[a]) + 1;
                first[a] = std::min(std::max<std::int64_t>(f, 0), counts[a] - 1);
                last[a] = std::min(std::max<std::int64_t>(t, 0), counts[a] - 1);
            }
            for (std::int64_t bz = first[2]; bz <= last[2]; bz++)
                for (std::int64_t by = first[1]; by <= last[1]; by++)
                    for (std::int64_t bx = first[0]; bx <= last[0]; bx++)
                        mine[(std::size_t) ((bz * counts[1] + by) * counts[0] + bx)].push_back((std::uint32_t) id);
        }
#pragma omp for ordered schedule(static, 1)
        for (int t = 0; t < omp_get_num_threads(); t++)
        {
#pragma omp ordered
            for (std::size_t b = 0; b < numBlocks; b++) blocks[b].insert(blocks[b].end(), mine[b].begin(), mine[b].end());
        }
    }
    // blocks are independent: split them in parallel, then flatten in z-major block order
    std::vector<std::vector<Bin>> perBlock(numBlocks);
#pragma omp parallel for schedule(dynamic, 1)
    for (std::int64_t bi = 0; bi < (std::int64_t) numBlocks; bi++)
    {
        const std::int64_t bx = bi % counts[0], by = (bi / counts[0]) % counts[1], bz = bi / (counts[0] * counts[1]);
        const std::int64_t lo[3] = {edge(0, bx), edge(1, by), edge(2, bz)};
        const std::int64_t hi[3] = {edge(0, bx + 1), edge(1, by + 1), edge(2, bz + 1)};
        detail::split(splats, lo, hi, blocks[(std::size_t) bi], maxCells, maxSplats, perBlock[(std::size_t) bi]);
    }
    for (std::vector<Bin> &v : perBlock)
        for (Bin &b : v) bins.push_back(std::move(b));
    return bins;
}

}  // namespace mlsb

#endif
