// Host mesher: welds the bins' meshes at their seams and prunes small components.
//
// An in-core restatement of what OOCMesher does with the blocks it receives (src/mesher.cpp):
//  * per block, a LOCAL union-find over the triangle edges groups its vertices into "clumps"
//    (computeLocalComponents / updateGlobalClumps, :220-278) — the block's nodes fit in cache, so
//    the 2 merges per triangle never touch the global arrays;
//  * vertices that are external to a block carry a 64-bit key and are unified across blocks through
//    a key -> vertex map; unifying two vertices also unifies their clumps and un-counts the shared
//    vertex (updateClumpKeyMap, :280-306);
//  * at write time components (root clumps) with fewer than floor(fit-prune x total vertices)
//    vertices are dropped and the rest are compacted (getStatistics :491-513, write :763-852).
// The reference streams blocks through temporary files to bound memory; this version keeps the
// mesh in host memory.
#ifndef MLSB200_MESHER_HPP
#define MLSB200_MESHER_HPP

#include <algorithm>
#include <array>
#include <cstdint>
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

        // ---- local components ----
        local.resize(nv);
        for (std::size_t i = 0; i < nv; i++) local[i] = (std::int32_t) i;
        for (std::size_t i = 0; i < nt; i++)
        {
            // two edges are enough: the third is redundant (src/mesher.cpp:230-234)
            mergeLocal(tris[i][0], tris[i][1]);
            mergeLocal(tris[i][1], tris[i][2]);
        }
        // ---- one clump per local component ----
        clumpOfLocal.resize(nv);
        for (std::size_t i = 0; i < nv; i++)
            if (local[i] == (std::int32_t) i)
            {
                clumpOfLocal[i] = (std::uint32_t) clumps.size();
                clumps.push_back(Clump{(std::uint32_t) clumps.size(), 0});
            }
        for (std::size_t i = 0; i < nv; i++)
        {
            const std::uint32_t c = clumpOfLocal[findLocal((std::int32_t) i)];
            clumpOfLocal[i] = c;
            clumps[c].vertices++;
        }
        // ---- global vertices: internal ones are new, external ones are looked up by key ----
        remap.resize(nv);
        // (no exact-size reserve here: it would defeat the vectors' geometric growth and recopy
        // the whole mesh for every block)
        for (std::size_t i = 0; i < ni; i++)
        {
            remap[i] = (std::uint32_t) vertices.size();
            vertices.push_back(verts[i]);
            vertexClump.push_back(clumpOfLocal[i]);
        }
        for (std::size_t i = ni; i < nv; i++)
        {
            auto added = keyMap.insert(std::make_pair(keys[i - ni], (std::uint32_t) vertices.size()));
            if (added.second)
            {
                vertices.push_back(verts[i]);
                vertexClump.push_back(clumpOfLocal[i]);
            }
            else
            {
                // the same vertex seen from two blocks: unify their clumps, count it once
                const std::uint32_t root = mergeClumps(clumpOfLocal[i], vertexClump[added.first->second]);
                clumps[root].vertices--;
            }
            remap[i] = added.first->second;
        }
        const std::size_t firstTriangle = triangles.size();
        triangles.resize(firstTriangle + nt);
        for (std::size_t i = 0; i < nt; i++)
            triangles[firstTriangle + i] = {remap[tris[i][0]], remap[tris[i][1]], remap[tris[i][2]]};
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
        std::uint64_t total = 0;
        for (std::size_t c = 0; c < clumps.size(); c++)
            if (clumps[c].parent == c) total += clumps[c].vertices;
        r.thresholdVertices = (std::uint64_t) (total * pruneThreshold);
        std::vector<char> keep(clumps.size(), 0);
        for (std::size_t c = 0; c < clumps.size(); c++)
            if (clumps[c].parent == c)
            {
                r.totalComponents++;
                keep[c] = clumps[c].vertices >= r.thresholdVertices;
                r.keptComponents += keep[c];
            }
        if (r.keptComponents == r.totalComponents)
        {
            // nothing to prune: hand the mesh over as it is
            r.vertices.swap(vertices);
            r.triangles.swap(triangles);
            return r;
        }
        std::vector<std::uint32_t> newIndex(n, UINT32_MAX);
        for (std::size_t i = 0; i < n; i++)
            if (keep[findClump(vertexClump[i])])
            {
                newIndex[i] = (std::uint32_t) r.vertices.size();
                r.vertices.push_back(vertices[i]);
            }
        r.triangles.reserve(triangles.size());
        for (const triangle_type &t : triangles)
            if (newIndex[t[0]] != UINT32_MAX)
            {
                triangle_type o = {newIndex[t[0]], newIndex[t[1]], newIndex[t[2]]};
                r.triangles.push_back(o);
            }
        return r;
    }

private:
    struct Clump
    {
        std::uint32_t parent;
        std::uint64_t vertices;
    };

    std::int32_t findLocal(std::int32_t x)
    {
        while (local[x] != x)
        {
            local[x] = local[local[x]];
            x = local[x];
        }
        return x;
    }
    void mergeLocal(std::uint32_t a, std::uint32_t b)
    {
        const std::int32_t ra = findLocal((std::int32_t) a), rb = findLocal((std::int32_t) b);
        if (ra != rb) local[std::max(ra, rb)] = std::min(ra, rb);
    }
    std::uint32_t findClump(std::uint32_t x)
    {
        while (clumps[x].parent != x)
        {
            clumps[x].parent = clumps[clumps[x].parent].parent;
            x = clumps[x].parent;
        }
        return x;
    }
    std::uint32_t mergeClumps(std::uint32_t a, std::uint32_t b)
    {
        a = findClump(a);
        b = findClump(b);
        if (a == b) return a;
        const std::uint32_t lo = std::min(a, b), hi = std::max(a, b);
        clumps[hi].parent = lo;
        clumps[lo].vertices += clumps[hi].vertices;
        return lo;
    }

    std::vector<vertex_type> vertices;
    std::vector<triangle_type> triangles;
    std::vector<std::uint32_t> vertexClump;   // clump of every global vertex (as first seen)
    std::vector<Clump> clumps;
    std::vector<std::int32_t> local;
    std::vector<std::uint32_t> clumpOfLocal, remap;
    std::unordered_map<std::uint64_t, std::uint32_t> keyMap;
};

}  // namespace mlsb

#endif
