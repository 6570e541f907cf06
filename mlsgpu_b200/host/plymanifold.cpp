// plymanifold — checks that a PLY mesh is an oriented manifold with boundary and prints its
// vertex, triangle, component and boundary counts: the reference's acceptance tool
// (extras/plymanifold.cpp:135-205), same output and exit codes (0 manifold, 1 not / unreadable,
// 2 usage).
#include <cmath>
#include <cstdio>
#include <iostream>

#include "manifold.hpp"
#include "ply_mesh.hpp"

int main(int argc, const char **argv)
{
    if (argc != 2)
    {
        std::cerr << "Usage: plymanifold file.ply\n";
        return 2;
    }
    const char *filename = argv[1];
    try
    {
        const mlsb::PlyMesh mesh = mlsb::readPlyMesh(filename);
        for (std::size_t i = 0; i < mesh.vertices.size(); i++)
            for (int j = 0; j < 3; j++)
                if (!std::isfinite(mesh.vertices[i][j]))
                {
                    std::cout << "Vertex " << i << " has non-finite value\n";
                    return 1;
                }
        mlsb::manifold::Metadata metadata;
        const std::string reason =
            mlsb::manifold::isManifold(mesh.vertices.size(), mesh.triangles.data(), mesh.triangles.size(), &metadata);
        if (!reason.empty())
        {
            std::printf("Mesh is not manifold: %s\n", reason.c_str());
            return 1;
        }
        // the report lines of extras/plymanifold.cpp:183-187
        std::printf("Mesh is manifold.\nVertices: %zu\nTriangles: %zu\nComponents: %zu\nBoundaries: %zu\n",
                    metadata.numVertices, metadata.numTriangles, metadata.numComponents, metadata.numBoundaries);
    }
    catch (mlsb::PlyFormatError &e)
    {
        std::cerr << "Failed to read " << filename << ": " << e.what() << "\n";
        return 1;
    }
    catch (std::exception &e)
    {
        std::cerr << e.what() << "\n";
        return 1;
    }
    return 0;
}
