// Out-of-core front-end: bins from inputs that do not fit --mem-load-splats, without ever holding
// more than one batch of splats in host memory.  The reference's scheme (blobs + recursive
// histogram partition, src/splat_set_impl.h:814-854, src/bucket_impl.h:439-561,
// src/bucket_collector.cpp:48-69, src/bucket_loader.cpp:56-128) relies on spatially coherent input
// (doc/mlsgpu-user-manual.xml:211-216): a first pass records, per run of consecutive splats, which
// coarse buckets it touches, and later passes read only the runs a group of bins needs.  The same idea
// here, on the block layout the in-core binning uses:
//
//   pass 1   every input chunk (64 Ki consecutive vertices): bounding box, finite count  -> the grid
//   pass 2   exact number of splat instances per block, block bounding box per chunk     -> batches:
//            runs of consecutive blocks whose instances fit --mem-load-splats
//   batch    read the chunks that can reach the batch's blocks, keep the splats that do (in stream
//            order), bin them in core (makeBins) and hand over the bins of the batch's blocks
//
// A bin produced this way holds the same splats in the same order as the in-core path gives it, so
// the meshes are identical (tests/test_host_tools.py checks with meshdiff).  The host mesher is still
// in core: the OUTPUT is held in memory until it is written.
#ifndef MLSB200_STREAM_HPP
#define MLSB200_STREAM_HPP

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <vector>

#include "bucket.hpp"
#include "ply_io.hpp"

namespace mlsb
{

struct InputChunk
{
    std::size_t file, first, count;     // vertices [first, first + count) of readers[file]
    std::size_t finite = 0;             // splats that survive the non-finite filter
    double mn[3] = {0, 0, 0}, mx[3] = {0, 0, 0};   // world bounds of position -+ radius over them
    std::int64_t blockFirst[3] = {0, 0, 0}, blockLast[3] = {-1, -1, -1};   // blocks they reach (pass 2); empty if blockFirst[0] > blockLast[0]
    std::vector<std::uint32_t> blocks;          // exactly which ones (sorted indices): a handful for a coherent chunk
};

class SplatStream
{
public:
    static constexpr std::size_t kChunkVertices = 1u << 16;

    explicit SplatStream(const std::vector<std::unique_ptr<PlyReader>> &readers) : readers(readers)
    {
        for (std::size_t f = 0; f < readers.size(); f++)
            for (std::size_t first = 0; first < readers[f]->size(); first += kChunkVertices)
            {
                InputChunk c;
                c.file = f;
                c.first = first;
                c.count = std::min(kChunkVertices, readers[f]->size() - first);
                chunks.push_back(c);
            }
    }

    /// Pass 1: bounds and finite counts per chunk; returns the number of finite splats.
    std::size_t scan()
    {
        std::size_t total = 0;
#pragma omp parallel reduction(+ : total)
        {
            std::vector<mlsb_splat> buf(kChunkVertices);
            std::vector<char> scratch;
#pragma omp for schedule(dynamic, 4)
            for (std::int64_t ci = 0; ci < (std::int64_t) chunks.size(); ci++)
            {
                InputChunk &c = chunks[(std::size_t) ci];
                readers[c.file]->readRange(c.first, c.count, buf.data(), scratch);
                for (int a = 0; a < 3; a++) c.mn[a] = 1e300, c.mx[a] = -1e300;
                for (std::size_t i = 0; i < c.count; i++)
                {
                    const mlsb_splat &s = buf[i];
                    if (!PlyReader::isFinite(s)) continue;
                    c.finite++;
                    for (int a = 0; a < 3; a++)
                    {
                        c.mn[a] = std::min(c.mn[a], (double) s.position[a] - s.radius);
                        c.mx[a] = std::max(c.mx[a], (double) s.position[a] + s.radius);
                    }
                }
                total += c.finite;
            }
        }
        return total;
    }

    /// The bounding grid of everything scanned (boundingGrid()'s rule).
    Grid grid(float spacing, std::int64_t bucketSize) const
    {
        double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
        for (const InputChunk &c : chunks)
            if (c.finite)
                for (int a = 0; a < 3; a++)
                {
                    mn[a] = std::min(mn[a], c.mn[a]);
                    mx[a] = std::max(mx[a], c.mx[a]);
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

    /// Pass 2: splat instances per block and the blocks each chunk reaches.
    std::vector<std::uint64_t> countBlocks(const Grid &g, const BlockLayout &layout)
    {
        std::vector<std::uint64_t> counts(layout.numBlocks(), 0);
#pragma omp parallel
        {
            std::vector<mlsb_splat> buf(kChunkVertices);
            std::vector<char> scratch;
            std::vector<std::pair<std::size_t, std::uint64_t>> local;   // (block, instances) of the chunk in hand
#pragma omp for schedule(dynamic, 4)
            for (std::int64_t ci = 0; ci < (std::int64_t) chunks.size(); ci++)
            {
                InputChunk &c = chunks[(std::size_t) ci];
                for (int a = 0; a < 3; a++) c.blockFirst[a] = INT64_MAX, c.blockLast[a] = INT64_MIN;
                if (c.finite == 0) continue;
                readers[c.file]->readRange(c.first, c.count, buf.data(), scratch);
                // instances per block of this chunk first (consecutive splats mostly reach the same block),
                // then one atomic add per block the chunk reaches
                c.blocks.clear();
                local.clear();
                std::size_t lastIndex = SIZE_MAX, lastSlot = 0;
                for (std::size_t i = 0; i < c.count; i++)
                {
                    mlsb_splat s = buf[i];
                    if (!PlyReader::isFinite(s)) continue;
                    toGrid(s, g);
                    std::int64_t first[3], last[3];
                    if (!(layout.range(s, 0, first[0], last[0]) && layout.range(s, 1, first[1], last[1]) && layout.range(s, 2, first[2], last[2])))
                        continue;
                    for (int a = 0; a < 3; a++)
                    {
                        c.blockFirst[a] = std::min(c.blockFirst[a], first[a]);
                        c.blockLast[a] = std::max(c.blockLast[a], last[a]);
                    }
                    for (std::int64_t bz = first[2]; bz <= last[2]; bz++)
                        for (std::int64_t by = first[1]; by <= last[1]; by++)
                            for (std::int64_t bx = first[0]; bx <= last[0]; bx++)
                            {
                                const std::size_t index = layout.index(bx, by, bz);
                                if (index != lastIndex)
                                {
                                    lastIndex = index;
                                    for (lastSlot = 0; lastSlot < local.size() && local[lastSlot].first != index; lastSlot++) {}
                                    if (lastSlot == local.size()) local.push_back(std::make_pair(index, (std::uint64_t) 0));
                                }
                                local[lastSlot].second++;
                            }
                }
                std::sort(local.begin(), local.end());
                for (const std::pair<std::size_t, std::uint64_t> &e : local)
                {
                    c.blocks.push_back((std::uint32_t) e.first);
                    std::uint64_t &slot = counts[e.first];
#pragma omp atomic
                    slot += e.second;
                }
            }
        }
        return counts;
    }

    /// Runs [begin, end) of consecutive blocks whose instances fit maxLoadSplats (a single block may exceed it).
    static std::vector<std::pair<std::size_t, std::size_t>> batches(const std::vector<std::uint64_t> &counts, std::uint64_t maxLoadSplats)
    {
        std::vector<std::pair<std::size_t, std::size_t>> out;
        std::size_t begin = 0;
        std::uint64_t sum = 0;
        for (std::size_t b = 0; b < counts.size(); b++)
        {
            if (b > begin && sum + counts[b] > maxLoadSplats)
            {
                if (sum > 0) out.push_back(std::make_pair(begin, b));
                begin = b;
                sum = 0;
            }
            sum += counts[b];
        }
        if (sum > 0) out.push_back(std::make_pair(begin, counts.size()));
        return out;
    }

    /// The splats (grid coordinates, stream order) that reach a block of [begin, end).
    void load(const Grid &g, const BlockLayout &layout, std::size_t begin, std::size_t end, std::vector<mlsb_splat> &out,
              std::size_t *chunksRead = nullptr)
    {
        std::vector<std::size_t> wanted;
        for (std::size_t ci = 0; ci < chunks.size(); ci++)
        {
            const InputChunk &c = chunks[ci];
            if (c.finite == 0 || c.blocks.empty()) continue;
            // the chunk is read again only if one of the blocks it reaches belongs to the batch
            const std::vector<std::uint32_t>::const_iterator it = std::lower_bound(c.blocks.begin(), c.blocks.end(), (std::uint32_t) begin);
            if (it != c.blocks.end() && *it < end) wanted.push_back(ci);
        }
        std::vector<std::vector<mlsb_splat>> parts(wanted.size());
#pragma omp parallel
        {
            std::vector<mlsb_splat> buf(kChunkVertices);
            std::vector<char> scratch;
#pragma omp for schedule(dynamic, 2)
            for (std::int64_t wi = 0; wi < (std::int64_t) wanted.size(); wi++)
            {
                const InputChunk &c = chunks[wanted[(std::size_t) wi]];
                readers[c.file]->readRange(c.first, c.count, buf.data(), scratch);
                std::vector<mlsb_splat> &part = parts[(std::size_t) wi];
                for (std::size_t i = 0; i < c.count; i++)
                {
                    mlsb_splat s = buf[i];
                    if (!PlyReader::isFinite(s)) continue;
                    toGrid(s, g);
                    std::int64_t first[3], last[3];
                    if (!(layout.range(s, 0, first[0], last[0]) && layout.range(s, 1, first[1], last[1]) && layout.range(s, 2, first[2], last[2])))
                        continue;
                    bool reaches = false;
                    for (std::int64_t bz = first[2]; bz <= last[2] && !reaches; bz++)
                        for (std::int64_t by = first[1]; by <= last[1] && !reaches; by++)
                        {
                            // a row of blocks is a run of consecutive indices
                            const std::size_t lo = layout.index(first[0], by, bz), hi = layout.index(last[0], by, bz);
                            reaches = hi >= begin && lo < end;
                        }
                    if (reaches) part.push_back(s);
                }
            }
        }
        std::size_t total = 0;
        for (const std::vector<mlsb_splat> &p : parts) total += p.size();
        out.clear();
        out.reserve(total);
        for (const std::vector<mlsb_splat> &p : parts) out.insert(out.end(), p.begin(), p.end());
        if (chunksRead) *chunksRead = wanted.size();
    }

    std::size_t numChunks() const { return chunks.size(); }

    /// Grid::worldToVertex + radius rescale for one splat, the float operations of worldToGrid().
    static void toGrid(mlsb_splat &s, const Grid &g)
    {
        const float inv = 1.0f / g.spacing;
        for (int a = 0; a < 3; a++) s.position[a] = (s.position[a] - 0.0f) * inv - (float) g.lo[a];
        s.radius *= inv;
    }

private:
    const std::vector<std::unique_ptr<PlyReader>> &readers;
    std::vector<InputChunk> chunks;
};

/// Index of the block a bin lies in (bins never straddle blocks).
inline std::size_t blockOfBin(const BlockLayout &layout, const Bin &bin)
{
    std::int64_t b[3];
    for (int a = 0; a < 3; a++)
    {
        const std::vector<float> &e = layout.edges[a];
        b[a] = (std::upper_bound(e.begin(), e.end() - 1, (float) bin.offset[a]) - e.begin()) - 1;
    }
    return layout.index(b[0], b[1], b[2]);
}

}  // namespace mlsb

#endif
