// One bin through the C++ mirror classes, written the way the reference's device worker is
// (DeviceWorkerGroupBase::Worker::operator(), src/workers.cpp:232-286).
//
//   mirror_demo splats.bin sizeX sizeY sizeZ offX offY offZ [maxCells]
//
// splats.bin: raw array of 32-byte splats in grid coordinates.  Prints
// "<vertices> <triangles> <internal> <fnv1a of the mesh blobs>" for the test to compare with the
// Python binding's result.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "mlsb200.hpp"

int main(int argc, char **argv)
{
    if (argc < 8)
    {
        std::fprintf(stderr, "usage: %s splats.bin sx sy sz ox oy oz [maxCells]\n", argv[0]);
        return 2;
    }
    std::FILE *f = std::fopen(argv[1], "rb");
    if (!f) { std::perror(argv[1]); return 2; }
    std::fseek(f, 0, SEEK_END);
    const long bytes = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    std::vector<mlsb::Splat> splats(bytes / sizeof(mlsb::Splat));
    if (std::fread(splats.data(), sizeof(mlsb::Splat), splats.size(), f) != splats.size()) return 2;
    std::fclose(f);

    std::uint32_t size[3], keyOffset[3], expanded[3];
    std::int32_t offset[3];
    for (int i = 0; i < 3; i++)
    {
        size[i] = std::atoi(argv[2 + i]);
        offset[i] = std::atoi(argv[5 + i]);
        keyOffset[i] = offset[i];
        expanded[i] = (size[i] + 7) / 8 * 8;   // roundUp(size, MlsFunctor::wgs), workers.cpp:255-257
    }
    const unsigned maxCells = argc > 8 ? std::atoi(argv[8]) : 255;
    const unsigned levels = 6, subsampling = 3;

    try
    {
        mlsb::Worker worker(0, levels, subsampling, splats.size(), maxCells, 64u << 20, mlsb::MLS_SHAPE_SPHERE, 1.0f);
        mlsb::SplatTreeCL tree(worker);
        mlsb::MlsFunctor input(worker);
        mlsb::Marching marching(worker);
        mlsb::ScaleBiasFilter scaleBias(worker);
        scaleBias.setScaleBias(1.0f, 0.0f, 0.0f, 0.0f);

        std::uint64_t verts = 0, tris = 0, internal = 0, hash = 1469598103934665603ull;
        mlsb::Marching::OutputFunctor output = [&](const mlsb::HostKeyMesh &mesh) {
            verts += mesh.numVertices();
            tris += mesh.numTriangles();
            internal += mesh.numInternalVertices();
            const unsigned char *p = reinterpret_cast<const unsigned char *>(mesh.vertexKeys);
            for (std::size_t i = 0; i < mesh.getHostBytes(); i++) hash = (hash ^ p[i]) * 1099511628211ull;
        };

        worker.uploadSplats(splats.data(), 0, splats.size());
        tree.enqueueBuild(0, splats.size(), expanded, offset, subsampling);
        input.set(offset, tree, subsampling);
        marching.generate(input, output, size, keyOffset);
        tree.clearSplats();
        std::printf("%llu %llu %llu %016llx\n", (unsigned long long) verts, (unsigned long long) tris,
                    (unsigned long long) internal, (unsigned long long) hash);
    }
    catch (const std::exception &e)
    {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
