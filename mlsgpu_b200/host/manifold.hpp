// Oriented-manifold-with-boundary check for triangle meshes: the acceptance test the reference
// applies to its output (Manifold::isManifold, test/manifold.h:80-232; extras/plymanifold.cpp).
//
// Same definition and the same diagnostics: indices in range, no repeated vertex in a triangle,
// no isolated vertex, no directed edge used twice, and around every vertex the opposite edges
// form either one ring or a set of open runs (so a vertex may sit on several boundary loops, which
// are then counted as one).  Components are those of the vertex/edge graph.
//
// The reference keeps a vector of (map, set) per vertex; meshes here have tens of millions of
// triangles, so this version builds one CSR table of opposite edges (counting sort by vertex) and
// walks each vertex's handful of arrows in place.
#ifndef MLSB200_MANIFOLD_HPP
#define MLSB200_MANIFOLD_HPP

#include <algorithm>
#include <array>
#include <cstddef>
#include <cstdint>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

namespace mlsb
{
namespace manifold
{

struct Metadata   // test/manifold.h:52-78
{
    std::size_t numVertices = 0;
    std::size_t numTriangles = 0;
    std::size_t numComponents = 0;   // of the vertex/edge graph
    std::size_t numBoundaries = 0;   // boundary loops; loops meeting at a vertex count once
};

namespace detail
{
// union-find with sizes; roots hold -size, the rest their parent
class Forest
{
public:
    explicit Forest(std::size_t n) : node(n, -1) {}
    std::size_t find(std::size_t x)
    {
        std::size_t root = x;
        while (node[root] >= 0) root = (std::size_t) node[root];
        while (node[x] >= 0)
        {
            const std::size_t next = (std::size_t) node[x];
            node[x] = (std::int64_t) root;
            x = next;
        }
        return root;
    }
    void merge(std::size_t a, std::size_t b)
    {
        a = find(a);
        b = find(b);
        if (a == b) return;
        if (node[a] > node[b]) std::swap(a, b);   // a is the larger tree
        node[a] += node[b];
        node[b] = (std::int64_t) a;
    }
    bool isRoot(std::size_t x) const { return node[x] < 0; }
    std::size_t size(std::size_t root) const { return (std::size_t) -node[root]; }
    void reset() { std::fill(node.begin(), node.end(), -1); }

private:
    std::vector<std::int64_t> node;
};
}  // namespace detail

/// Returns "" if the mesh is manifold, else the reason (the reference's wording).  On failure
/// *data is zeroed.  Triangle is any type indexable with [0..2] giving an unsigned index of at most 32 bits.
template <typename Triangle>
std::string isManifold(std::size_t numVertices, const Triangle *triangles, std::size_t numTriangles, Metadata *data = nullptr)
{
    typedef std::uint32_t index_type;   // PLY faces carry at most 32-bit indices
    std::ostringstream reason;
    if (data) *data = Metadata();
    Metadata out;
    out.numVertices = numVertices;
    detail::Forest forest(numVertices);

    // ---- per-triangle checks, components, and the degree of every vertex ----
    std::vector<std::size_t> start(numVertices + 1, 0);
    for (std::size_t t = 0; t < numTriangles; t++)
    {
        const index_type v[3] = {triangles[t][0], triangles[t][1], triangles[t][2]};
        for (int j = 0; j < 3; j++)
        {
            if (v[j] >= numVertices)
            {
                reason << "Triangle " << t << " contains out-of-range index " << v[j] << "\n";
                return reason.str();
            }
            if (v[j] == v[(j + 1) % 3])
            {
                reason << "Triangle " << t << " contains vertex " << v[j] << " twice\n";
                return reason.str();
            }
        }
        for (int j = 0; j < 3; j++) start[v[j] + 1]++;
        forest.merge(v[0], v[1]);
        forest.merge(v[1], v[2]);
    }
    out.numTriangles = numTriangles;
    for (std::size_t i = 0; i < numVertices; i++) out.numComponents += forest.isRoot(i);
    forest.reset();   // reused for the boundary loops

    // ---- opposite edges of every vertex, in triangle order (CSR) ----
    for (std::size_t i = 0; i < numVertices; i++) start[i + 1] += start[i];
    std::vector<std::pair<index_type, index_type>> arrows(3 * numTriangles);
    {
        std::vector<std::size_t> cursor(start.begin(), start.end() - 1);
        for (std::size_t t = 0; t < numTriangles; t++)
            for (int j = 0; j < 3; j++)
                arrows[cursor[triangles[t][j]]++] = std::make_pair((index_type) triangles[t][(j + 1) % 3], (index_type) triangles[t][(j + 2) % 3]);
    }

    // ---- the neighbourhood of every vertex is a ring or a set of runs ----
    std::vector<std::pair<index_type, std::size_t>> order;   // scratch: (value, position)
    std::vector<std::pair<index_type, index_type>> sorted;
    for (std::size_t i = 0; i < numVertices; i++)
    {
        const std::pair<index_type, index_type> *neigh = arrows.data() + start[i];
        const std::size_t degree = start[i + 1] - start[i];
        if (degree == 0)
        {
            reason << "Vertex " << i << " is isolated\n";
            return reason.str();
        }
        // The reference scans the arrows in order and stops at the first one whose tail (.first)
        // or head (.second) was already seen, tails checked before heads.  Find that position.
        std::size_t badTail = degree, badHead = degree;
        for (int side = 0; side < 2; side++)
        {
            order.resize(degree);
            for (std::size_t j = 0; j < degree; j++) order[j] = std::make_pair(side ? neigh[j].second : neigh[j].first, j);
            std::sort(order.begin(), order.end());
            std::size_t &bad = side ? badHead : badTail;
            for (std::size_t j = 1; j < degree; j++)
                if (order[j].first == order[j - 1].first) bad = std::min(bad, order[j].second);
        }
        if (badTail < degree && badTail <= badHead)
        {
            reason << "Edge " << i << " - " << neigh[badTail].first << " occurs twice with same winding\n";
            return reason.str();
        }
        if (badHead < degree)
        {
            reason << "Edge " << neigh[badHead].second << " - " << i << " occurs twice with same winding\n";
            return reason.str();
        }
        // In- and out-degree are now at most one: runs and rings.  sorted = arrows by tail.
        sorted.assign(neigh, neigh + degree);
        std::sort(sorted.begin(), sorted.end());
        auto next = [&](index_type x, index_type &y) {
            auto pos = std::lower_bound(sorted.begin(), sorted.end(), std::make_pair(x, (index_type) 0));
            if (pos == sorted.end() || pos->first != x) return false;
            y = pos->second;
            return true;
        };
        std::size_t len = 0;
        for (std::size_t j = 0; j < degree; j++)
        {
            // is this tail the head of another arrow?  (order holds the heads, sorted)
            const index_type first = neigh[j].first;
            auto pos = std::lower_bound(order.begin(), order.end(), std::make_pair(first, (std::size_t) 0));
            if (pos != order.end() && pos->first == first) continue;
            index_type cur = first, y;
            while (next(cur, y))
            {
                cur = y;
                len++;
            }
            forest.merge(first, i);   // both ends of the run and the vertex lie on one boundary
            forest.merge(cur, i);
        }
        if (len != 0 && len != degree)
        {
            reason << "Vertex " << i << " is both in the interior and on the boundary\n";
            return reason.str();
        }
        if (len == 0)
        {
            const index_type first = neigh[0].first;
            index_type cur = first;
            do
            {
                next(cur, cur);
                len++;
            } while (cur != first);
            if (len != degree)
            {
                reason << "Vertex " << i << " tunnels between interior regions\n";
                return reason.str();
            }
        }
    }
    for (std::size_t i = 0; i < numVertices; i++)
        if (forest.isRoot(i) && forest.size(i) >= 3) out.numBoundaries++;
    if (data) *data = out;
    return "";
}

}  // namespace manifold
}  // namespace mlsb

#endif
