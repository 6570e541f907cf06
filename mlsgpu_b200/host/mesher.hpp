// Host mesher: welds the bins' meshes at their seams and prunes small components.
//
// An in-core restatement of what OOCMesher does with the blocks it receives (src/mesher.cpp):
// vertices that are external to a block carry a 64-bit key and are unified across blocks through a
// key -> vertex map (updateClumpKeyMap, :280-306); connected components are tracked with a
// union-find over triangle edges (computeLocalComponents / updateGlobalClumps, :220-278); at write
// time components with fewer than floor(fit-prune x total vertices) vertices are dropped and the
// rest are compacted (getStatistics :491-513, write :763-852).  The reference streams blocks
// through temporary files to bound memory; this version keeps the mesh in host memory.
#ifndef MLSB200_MESHER_HPP
#define MLSB200_MESHER_HPP

#include <array>
#include <cstdint>
#include <numeric>
#include <unordered_map>
#include <vector>

#include "../../include/mlsb200.h"

namespace mlsb
{

class Mesher
{
public:
    typedef std::array<float, 3> vertex_type;
    typedef std::array<std::uint32_t, 3> triangle_type;

    /// One shipped-out block in the HostKeyMesh layout (src/mesh.cpp:51-60).  Not thread-safe:
    /// the worker group delivers blocks serialised, like the reference's single mesher thread.
    void add(const mlsb_mesh_sizes &sizes, const void *blob)
    {
        const std::size_t nv = sizes.numVertices, nt = sizes.numTriangles, ni = sizes.numInternalVertices;
        const std::uint64_t *keys = static_cast<const std::uint64_t *>(blob);
        const vertex_type *verts = reinterpret_cast<const vertex_type *>(keys + (nv - ni));
        const triangle_type *tris = reinterpret_cast<const triangle_type *>(verts + nv);
        remap.resize(nv);
        for (std::size_t i = 0; i < ni; i++) remap[i] = newVertex(verts[i]);
        for (std::size_t i = ni; i < nv; i++)
        {
            auto added = keyMap.insert(std::make_pair(keys[i - ni], (std::uint32_t) vertices.size()));
            if (added.second) newVertex(verts[i]);
            remap[i] = added.first->second;
        }
        triangles.reserve(triangles.size() + nt);
        for (std::size_t i = 0; i < nt; i++)
        {
            triangle_type t = {remap[tris[i][0]], remap[tris[i][1]], remap[tris[i][2]]};
            triangles.push_back(t);
            merge(t[0], t[1]);
            merge(t[1], t[2]);
        }
    }

    std::size_t numVertices() const { return vertices.size(); }
    std::size_t numTriangles() const { return triangles.size(); }

    struct Result
    {
        std::vector<vertex_type> vertices;
        std::vector<triangle_type> triangles;
        std::size_t totalComponents = 0, keptComponents = 0;
        std::uint64_t thresholdVertices = 0;
    };

    /// Prune and compact.  pruneThreshold = --fit-prune (fraction of all welded vertices).
    Result finish(double pruneThreshold)
    {
        Result r;
        const std::size_t n = vertices.size();
        std::vector<std::uint32_t> componentSize(n, 0);
        for (std::size_t i = 0; i < n; i++) componentSize[find((std::uint32_t) i)]++;
        r.thresholdVertices = (std::uint64_t) (n * pruneThreshold);
        std::vector<std::uint32_t> newIndex(n, UINT32_MAX);
        for (std::size_t i = 0; i < n; i++)
        {
            const std::uint32_t root = find((std::uint32_t) i);
            if (root == i)
            {
                r.totalComponents++;
                if (componentSize[root] >= r.thresholdVertices) r.keptComponents++;
            }
            if (componentSize[root] >= r.thresholdVertices)
            {
                newIndex[i] = (std::uint32_t) r.vertices.size();
                r.vertices.push_back(vertices[i]);
            }
        }
        for (const triangle_type &t : triangles)
            if (newIndex[t[0]] != UINT32_MAX)
            {
                triangle_type o = {newIndex[t[0]], newIndex[t[1]], newIndex[t[2]]};
                r.triangles.push_back(o);
            }
        return r;
    }

private:
    std::uint32_t newVertex(const vertex_type &v)
    {
        vertices.push_back(v);
        parent.push_back((std::uint32_t) parent.size());
        return (std::uint32_t) (vertices.size() - 1);
    }
    std::uint32_t find(std::uint32_t x)
    {
        while (parent[x] != x)
        {
            parent[x] = parent[parent[x]];
            x = parent[x];
        }
        return x;
    }
    void merge(std::uint32_t a, std::uint32_t b)
    {
        a = find(a);
        b = find(b);
        if (a != b) parent[std::max(a, b)] = std::min(a, b);
    }

    std::vector<vertex_type> vertices;
    std::vector<triangle_type> triangles;
    std::vector<std::uint32_t> parent, remap;
    std::unordered_map<std::uint64_t, std::uint32_t> keyMap;
};

}  // namespace mlsb

#endif
