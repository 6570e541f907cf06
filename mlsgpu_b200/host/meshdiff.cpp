// meshdiff — compares two PLY meshes up to vertex numbering and triangle order/rotation.
//
// The reference's output order varies from run to run (doc/mlsgpu-user-manual.xml:497-501), so
// two runs — or the reference and this library — can only be compared after canonicalising:
// vertices are identified by position (exactly, or within --tolerance on a lattice of that
// pitch), triangles by their three vertices rotated so that the smallest comes first (winding is
// kept).  Prints what is common and what is on one side only; exit 0 if the meshes are the same,
// 1 if they differ, 2 on usage or read errors.
//
//   meshdiff [--tolerance EPS] [--show N] a.ply b.ply
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <string>
#include <vector>

#include "ply_mesh.hpp"

namespace
{

typedef std::array<std::int64_t, 3> Cell;        // position on the comparison lattice
typedef std::array<std::uint32_t, 3> Triangle;

// Exact mode keys a vertex by the bit patterns of its coordinates (-0 folded into +0); tolerance
// mode by the nearest lattice point, which identifies positions that differ by much less than EPS
// (positions within EPS of a lattice boundary can still land apart: they are reported, not hidden).
Cell cellOf(const std::array<float, 3> &v, double tolerance)
{
    Cell c;
    for (int a = 0; a < 3; a++)
    {
        if (tolerance > 0.0)
            c[a] = (std::int64_t) std::llround((double) v[a] / tolerance);
        else
        {
            float f = v[a] == 0.0f ? 0.0f : v[a];
            std::int32_t bits;
            std::memcpy(&bits, &f, 4);
            c[a] = bits >= 0 ? bits : (std::int64_t) INT32_MIN - bits;   // monotone in the float order
        }
    }
    return c;
}

struct Canonical
{
    std::vector<Cell> cells;                // sorted, unique
    std::vector<std::array<Cell, 3>> triangles;   // sorted
    std::size_t duplicateVertices = 0;
};

Canonical canonicalise(const mlsb::PlyMesh &mesh, double tolerance)
{
    Canonical c;
    std::vector<Cell> of(mesh.vertices.size());
    for (std::size_t i = 0; i < mesh.vertices.size(); i++) of[i] = cellOf(mesh.vertices[i], tolerance);
    c.cells = of;
    std::sort(c.cells.begin(), c.cells.end());
    const std::size_t before = c.cells.size();
    c.cells.erase(std::unique(c.cells.begin(), c.cells.end()), c.cells.end());
    c.duplicateVertices = before - c.cells.size();
    c.triangles.reserve(mesh.triangles.size());
    for (const Triangle &t : mesh.triangles)
    {
        std::array<Cell, 3> k;
        for (int j = 0; j < 3; j++)
        {
            if (t[j] >= of.size()) throw std::runtime_error("triangle index out of range");
            k[j] = of[t[j]];
        }
        std::rotate(k.begin(), std::min_element(k.begin(), k.end()), k.end());
        c.triangles.push_back(k);
    }
    std::sort(c.triangles.begin(), c.triangles.end());
    return c;
}

template <typename T>
void split(const std::vector<T> &a, const std::vector<T> &b, std::vector<T> &onlyA, std::vector<T> &onlyB)
{
    std::set_difference(a.begin(), a.end(), b.begin(), b.end(), std::back_inserter(onlyA));
    std::set_difference(b.begin(), b.end(), a.begin(), a.end(), std::back_inserter(onlyB));
}

void show(const char *label, const std::vector<Cell> &cells, double tolerance, std::size_t limit)
{
    for (std::size_t i = 0; i < cells.size() && i < limit; i++)
    {
        std::cout << "  " << label << " vertex";
        for (int a = 0; a < 3; a++)
        {
            if (tolerance > 0.0)
                std::cout << ' ' << cells[i][a] * tolerance;
            else
            {
                const std::int64_t m = cells[i][a];
                const std::int32_t bits = (std::int32_t) (m >= 0 ? m : (std::int64_t) INT32_MIN - m);
                float f;
                std::memcpy(&f, &bits, 4);
                std::cout << ' ' << f;
            }
        }
        std::cout << '\n';
    }
}

}  // namespace

int main(int argc, char **argv)
{
    double tolerance = 0.0;
    std::size_t limit = 10;
    std::vector<std::string> files;
    for (int i = 1; i < argc; i++)
    {
        const std::string a = argv[i];
        if (a == "--tolerance" && i + 1 < argc) tolerance = std::atof(argv[++i]);
        else if (a == "--show" && i + 1 < argc) limit = (std::size_t) std::atol(argv[++i]);
        else if (a.rfind("--", 0) == 0) files.clear(), i = argc;
        else files.push_back(a);
    }
    if (files.size() != 2 || tolerance < 0.0)
    {
        std::cerr << "Usage: meshdiff [--tolerance EPS] [--show N] a.ply b.ply\n";
        return 2;
    }
    try
    {
        const mlsb::PlyMesh meshA = mlsb::readPlyMesh(files[0]), meshB = mlsb::readPlyMesh(files[1]);
        const Canonical a = canonicalise(meshA, tolerance), b = canonicalise(meshB, tolerance);
        std::vector<Cell> vertsA, vertsB;
        std::vector<std::array<Cell, 3>> trisA, trisB;
        split(a.cells, b.cells, vertsA, vertsB);
        split(a.triangles, b.triangles, trisA, trisB);
        std::cout << "A: " << meshA.vertices.size() << " vertices, " << meshA.triangles.size() << " triangles\n"
                  << "B: " << meshB.vertices.size() << " vertices, " << meshB.triangles.size() << " triangles\n"
                  << "coincident vertices within a mesh: A " << a.duplicateVertices << ", B " << b.duplicateVertices << "\n"
                  << "vertices: " << a.cells.size() - vertsA.size() << " common, " << vertsA.size() << " only in A, "
                  << vertsB.size() << " only in B\n"
                  << "triangles: " << a.triangles.size() - trisA.size() << " common, " << trisA.size() << " only in A, "
                  << trisB.size() << " only in B\n";
        show("A-only", vertsA, tolerance, limit);
        show("B-only", vertsB, tolerance, limit);
        const bool same = vertsA.empty() && vertsB.empty() && trisA.empty() && trisB.empty() &&
                          meshA.vertices.size() == meshB.vertices.size() && meshA.triangles.size() == meshB.triangles.size();
        std::cout << (same ? "Meshes are identical up to ordering.\n" : "Meshes differ.\n");
        return same ? 0 : 1;
    }
    catch (std::exception &e)
    {
        std::cerr << "meshdiff: " << e.what() << "\n";
        return 2;
    }
}
