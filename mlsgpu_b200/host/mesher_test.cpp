// Test harness for the host mesher and PLY writer (driven by tests/test_host_tools.py).
//   mesher_test blocks.bin prune out.ply
// blocks.bin: uint64 numBlocks, then per block uint64 nv, nt, ni followed by the HostKeyMesh blob
// ([keys of external vertices][float3 vertices][uint32x3 triangles]).
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <vector>

#include "mesher.hpp"
#include "mlsb200.hpp"
#include "ply_io.hpp"

int main(int argc, char **argv)
{
    if (argc != 4) return 2;
    std::FILE *f = std::fopen(argv[1], "rb");
    if (!f) return 2;
    std::uint64_t numBlocks = 0;
    if (std::fread(&numBlocks, 8, 1, f) != 1) return 2;
    mlsb::Mesher mesher;
    for (std::uint64_t b = 0; b < numBlocks; b++)
    {
        std::uint64_t h[3];
        if (std::fread(h, 8, 3, f) != 3) return 2;
        mlsb_mesh_sizes sizes{h[0], h[1], h[2]};
        std::vector<std::uint64_t> blob((mlsb::MeshSizes(sizes).getHostBytes() + 7) / 8 + 1);
        const std::size_t bytes = 8 * (h[0] - h[2]) + 12 * h[0] + 12 * h[1];
        if (bytes && std::fread(blob.data(), 1, bytes, f) != bytes) return 2;
        mesher.add(sizes, blob.data());
    }
    std::fclose(f);
    mlsb::Mesher::Result r = mesher.finish(std::atof(argv[2]));
    mlsb::PlyWriter writer;
    writer.addComment("mesher_test");
    writer.write(argv[3], r.vertices, r.triangles);
    std::cout << mesher.numVertices() << " " << mesher.numTriangles() << " " << r.totalComponents << " " << r.keptComponents
              << " " << r.thresholdVertices << "\n";
    return 0;
}
