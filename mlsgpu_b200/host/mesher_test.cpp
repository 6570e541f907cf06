// Test harness for the host mesher and PLY writer (driven by tests/test_host_tools.py).
//   mesher_test blocks.bin prune out.ply [threads [checkpoint]]
// With a checkpoint file name the mesher's state goes through Mesher::save and Mesher::load (into a
// fresh object) before it is finished: --checkpoint / --resume.
// With threads > 1 the blocks go through Mesher::prepare / commit / complete from that many threads,
// the way the device workers' sink drives it in mlsgpu.cpp.
// If blocks.bin continues after the last block with uint32 x 3 chunk coordinates per block, the
// mesher runs in --split mode and out.ply is the base name of the chunk files
// (base_XXXX_YYYY_ZZZZ.ply, ChunkNamer, src/mesher.cpp:88-96); one line per written file follows
// the summary line.
// blocks.bin: uint64 numBlocks, then per block uint64 nv, nt, ni followed by the HostKeyMesh blob
// ([keys of external vertices][float3 vertices][uint32x3 triangles]).
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <mutex>
#include <thread>
#include <vector>

#include "mesher.hpp"
#include "mlsb200.hpp"
#include "ply_io.hpp"

int main(int argc, char **argv)
{
    if (argc < 4 || argc > 6) return 2;
    const int threads = argc >= 5 ? std::atoi(argv[4]) : 1;
    std::FILE *f = std::fopen(argv[1], "rb");
    if (!f) return 2;
    std::uint64_t numBlocks = 0;
    if (std::fread(&numBlocks, 8, 1, f) != 1) return 2;
    mlsb::Mesher mesher;
    std::vector<mlsb_mesh_sizes> sizes(numBlocks);
    std::vector<std::vector<std::uint64_t>> blobs(numBlocks);
    for (std::uint64_t b = 0; b < numBlocks; b++)
    {
        std::uint64_t h[3];
        if (std::fread(h, 8, 3, f) != 3) return 2;
        sizes[b] = mlsb_mesh_sizes{h[0], h[1], h[2]};
        blobs[b].resize((mlsb::MeshSizes(sizes[b]).getHostBytes() + 7) / 8 + 1);
        const std::size_t bytes = 8 * (h[0] - h[2]) + 12 * h[0] + 12 * h[1];
        if (bytes && std::fread(blobs[b].data(), 1, bytes, f) != bytes) return 2;
    }
    std::vector<mlsb::Mesher::chunk_type> chunkIds(numBlocks, mlsb::Mesher::chunk_type{{0, 0, 0}});
    const bool split = numBlocks > 0 && std::fread(chunkIds.data(), 12, numBlocks, f) == numBlocks;
    std::fclose(f);
    if (split) mesher = mlsb::Mesher(true);
    // MLSB_MESHER_TMP=dir: the out-of-core mesher (blocks spilled to a temporary file there, output streamed)
    const char *tmpDir = std::getenv("MLSB_MESHER_TMP");
    // MLSB_MESHER_KEEP=bytes: blocks stay in memory up to that many bytes, the rest is spilled
    const char *keep = std::getenv("MLSB_MESHER_KEEP");
    if (tmpDir && argc != 6) mesher.enableSpill(tmpDir, keep ? std::strtoull(keep, nullptr, 10) : 0);
    if (threads <= 1)
        for (std::uint64_t b = 0; b < numBlocks; b++)
        {
            mlsb::Mesher::Block &block = mesher.commit(mlsb::Mesher::prepare(sizes[b], blobs[b].data()), chunkIds[b]);
            mlsb::Mesher::complete(block);
            mesher.spill(block);
        }
    else
    {
        std::mutex mutex;
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; t++)
            pool.emplace_back([&, t] {
                for (std::uint64_t b = (std::uint64_t) t; b < numBlocks; b += (std::uint64_t) threads)
                {
                    std::unique_ptr<mlsb::Mesher::Block> prepared = mlsb::Mesher::prepare(sizes[b], blobs[b].data());
                    mlsb::Mesher::Block *block;
                    {
                        std::lock_guard<std::mutex> lock(mutex);
                        block = &mesher.commit(std::move(prepared), chunkIds[b]);
                    }
                    mlsb::Mesher::complete(*block);
                    mesher.spill(*block);
                }
            });
        for (std::thread &t : pool) t.join();
    }
    if (argc == 6)
    {
        {
            std::ofstream out(argv[5], std::ios::binary);
            mesher.save(out);
        }
        const std::size_t nv = mesher.numVertices(), nt = mesher.numTriangles();
        mesher = mlsb::Mesher();
        std::ifstream in(argv[5], std::ios::binary);
        mesher.load(in);
        if (mesher.numVertices() != nv || mesher.numTriangles() != nt) return 3;
    }
    mlsb::PlyWriter writer;
    writer.addComment("mesher_test");
    if (mesher.spilling())
    {
        const std::size_t nv = mesher.numVertices(), nt = mesher.numTriangles();
        std::vector<std::string> names;
        const std::vector<mlsb::Mesher::Result> all = mesher.stream(std::atof(argv[2]),
            [&](const mlsb::Mesher::chunk_type &chunk, std::size_t numV, std::size_t numT) {
                if (split && numT == 0) return std::unique_ptr<mlsb::PlyStreamWriter>();
                names.push_back(split ? mlsb::chunkName(argv[3], chunk) : std::string(argv[3]));
                return writer.openStream(names.back(), numV, numT);
            });
        if (!split && names.empty()) writer.write(argv[3], std::vector<mlsb::PlyWriter::vertex_type>(), std::vector<mlsb::PlyWriter::triangle_type>());
        std::cout << nv << " " << nt << " " << all[0].totalComponents << " " << all[0].keptComponents << " " << all[0].thresholdVertices << "\n";
        if (split)
        {
            std::size_t k = 0;
            for (const mlsb::Mesher::Result &r : all)
                if (r.numTriangles != 0) std::cout << names[k++] << " " << r.numVertices << " " << r.numTriangles << "\n";
        }
        return 0;
    }
    if (!split)
    {
        mlsb::Mesher::Result r = mesher.finish(std::atof(argv[2]));
        writer.write(argv[3], r.vertexChunks, r.triangleChunks);
        std::cout << mesher.numVertices() << " " << mesher.numTriangles() << " " << r.totalComponents << " " << r.keptComponents
                  << " " << r.thresholdVertices << "\n";
        return 0;
    }
    std::vector<mlsb::Mesher::Result> all = mesher.finishChunks(std::atof(argv[2]));
    std::cout << mesher.numVertices() << " " << mesher.numTriangles() << " " << (all.empty() ? 0 : all[0].totalComponents) << " "
              << (all.empty() ? 0 : all[0].keptComponents) << " " << (all.empty() ? 0 : all[0].thresholdVertices) << "\n";
    for (const mlsb::Mesher::Result &r : all)
    {
        if (r.numTriangles == 0) continue;   // only files with at least one triangle are written
        const std::string name = mlsb::chunkName(argv[3], r.chunk);
        writer.write(name, r.vertexChunks, r.triangleChunks);
        std::cout << name << " " << r.numVertices << " " << r.numTriangles << "\n";
    }
    return 0;
}
