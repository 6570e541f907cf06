// stream_test — the streaming front-end (stream.hpp) against the in-core binning (bucket.hpp) on the
// same PLY inputs: every bin must come out with the same box, the same chunk and the same splats in
// the same order, whatever the load limit.
//
//   stream_test FIT_GRID MAX_CELLS MAX_BIN_SPLATS MAX_LOAD_SPLATS CHUNK_CELLS file.ply [file.ply ...]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <tuple>

#include "stream.hpp"

int main(int argc, char **argv)
{
    try
    {
        if (argc < 7) throw std::runtime_error("usage: stream_test FIT_GRID MAX_CELLS MAX_BIN_SPLATS MAX_LOAD_SPLATS CHUNK_CELLS file.ply...");
        const float fitGrid = (float) std::atof(argv[1]);
        const std::int64_t maxCells = std::atoll(argv[2]);
        const std::size_t maxBinSplats = (std::size_t) std::atoll(argv[3]);
        const std::uint64_t maxLoad = (std::uint64_t) std::atoll(argv[4]);
        const std::int64_t chunkCells = std::atoll(argv[5]);
        std::vector<std::unique_ptr<mlsb::PlyReader>> readers;
        for (int i = 6; i < argc; i++) readers.emplace_back(new mlsb::PlyReader(argv[i], 4.0f, 1e30f));
        const std::int64_t bucketSize = std::min<std::int64_t>(63, maxCells);

        // in core
        std::vector<mlsb_splat> all;
        for (auto &r : readers) r->read(0, r->size(), all);
        const mlsb::Grid grid = mlsb::boundingGrid(all, fitGrid, bucketSize);
        mlsb::worldToGrid(all, grid);
        const std::vector<mlsb::Bin> want = mlsb::makeBins(all, grid, maxCells, maxBinSplats, chunkCells);
        typedef std::tuple<int, int, int, unsigned, unsigned, unsigned> Key;
        std::map<Key, const mlsb::Bin *> index;
        for (const mlsb::Bin &b : want) index[Key(b.offset[0], b.offset[1], b.offset[2], b.size[0], b.size[1], b.size[2])] = &b;

        // streamed
        mlsb::SplatStream stream(readers);
        const std::size_t n = stream.scan();
        if (n != all.size()) throw std::runtime_error("finite splat count differs");
        const mlsb::Grid g2 = stream.grid(fitGrid, bucketSize);
        for (int a = 0; a < 3; a++)
            if (g2.lo[a] != grid.lo[a] || g2.hi[a] != grid.hi[a]) throw std::runtime_error("grid differs");
        const mlsb::BlockLayout layout(g2, maxCells, chunkCells);
        const std::vector<std::uint64_t> counts = stream.countBlocks(g2, layout);
        const auto batches = mlsb::SplatStream::batches(counts, maxLoad);
        std::size_t found = 0, chunksRead = 0, largestLoad = 0;
        std::vector<mlsb_splat> local;
        for (const auto &batch : batches)
        {
            std::size_t read = 0;
            stream.load(g2, layout, batch.first, batch.second, local, &read);
            chunksRead += read;
            largestLoad = std::max(largestLoad, local.size());
            std::uint64_t expect = 0;
            for (std::size_t b = batch.first; b < batch.second; b++) expect += counts[b];
            if (batch.second - batch.first > 1 && expect > maxLoad) throw std::runtime_error("batch exceeds the load limit");
            for (const mlsb::Bin &b : mlsb::makeBins(local, g2, maxCells, maxBinSplats, chunkCells))
            {
                const std::size_t block = mlsb::blockOfBin(layout, b);
                if (block < batch.first || block >= batch.second) continue;
                auto it = index.find(Key(b.offset[0], b.offset[1], b.offset[2], b.size[0], b.size[1], b.size[2]));
                if (it == index.end()) throw std::runtime_error("streamed bin has no in-core counterpart");
                const mlsb::Bin &w = *it->second;
                if (std::memcmp(w.chunk, b.chunk, sizeof(b.chunk)) != 0) throw std::runtime_error("chunk differs");
                if (w.ids.size() != b.ids.size()) throw std::runtime_error("bin population differs");
                for (std::size_t i = 0; i < b.ids.size(); i++)
                    if (std::memcmp(&all[w.ids[i]], &local[b.ids[i]], sizeof(mlsb_splat)) != 0) throw std::runtime_error("bin contents or order differ");
                index.erase(it);
                found++;
            }
        }
        if (!index.empty()) throw std::runtime_error("bins missing from the streamed run");
        std::printf("ok: %zu splats, %zu bins, %zu batches, %zu of %zu chunks read in all, largest load %zu splats\n", n, found,
                    batches.size(), chunksRead, stream.numChunks(), largestLoad);
        return 0;
    }
    catch (const std::exception &e)
    {
        std::fprintf(stderr, "stream_test: %s\n", e.what());
        return 1;
    }
}
