// Test harness for the binning front-end (driven by tests/test_host_tools.py).
//   bucket_test numSplats seed fitGrid maxCells maxSplats chunkCells
// Scatters splats over a bumpy shell, bins them and checks the contract of Bucket::bucket
// (src/bucket_impl.h:439-561) by brute force: a bin holds EXACTLY the splats whose ball touches it
// (ascending ids), has at most maxCells cells per side and at most maxSplats splats, lies inside one
// chunk, and every cell of the grid is covered by exactly one bin unless no splat touches it.
#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <random>

#include "bucket.hpp"

int main(int argc, char **argv)
{
    if (argc != 7) return 2;
    const std::size_t n = (std::size_t) std::atol(argv[1]);
    const unsigned seed = (unsigned) std::atol(argv[2]);
    const float fitGrid = (float) std::atof(argv[3]);
    const std::int64_t maxCells = std::atol(argv[4]);
    const std::size_t maxSplats = (std::size_t) std::atol(argv[5]);
    const std::int64_t chunkCells = std::atol(argv[6]);

    std::mt19937 rng(seed);
    std::uniform_real_distribution<float> angle(0.0f, 6.2831853f), radius(0.004f, 0.03f), jitter(-0.3f, 0.3f);
    std::vector<mlsb_splat> splats(n);
    for (mlsb_splat &s : splats)
    {
        const float u = angle(rng), v = angle(rng), r = 0.4f + 0.05f * std::sin(5 * u);
        s.position[0] = (1.0f + r * std::cos(v)) * std::cos(u) + jitter(rng) * 0.01f;
        s.position[1] = (1.0f + r * std::cos(v)) * std::sin(u) - 3.0f;       // negative coordinates too
        s.position[2] = r * std::sin(v) + 10.0f;
        s.radius = radius(rng);
    }
    const mlsb::Grid grid = mlsb::boundingGrid(splats, fitGrid, 63);
    mlsb::worldToGrid(splats, grid);
    const std::vector<mlsb::Bin> bins = mlsb::makeBins(splats, grid, maxCells, maxSplats, chunkCells);

    std::map<std::array<std::uint32_t, 3>, int> chunks;
    std::size_t instances = 0;
    std::vector<int> covered(n, 0);
    for (const mlsb::Bin &b : bins)
    {
        chunks[{{b.chunk[0], b.chunk[1], b.chunk[2]}}]++;
        instances += b.ids.size();
        if (b.ids.empty() || b.ids.size() > maxSplats) return std::printf("FAIL: bin with %zu splats\n", b.ids.size()), 1;
        std::size_t at = 0;
        for (std::size_t i = 0; i < n; i++)
        {
            bool in = true;
            for (int a = 0; a < 3; a++)
                in = in && splats[i].position[a] + splats[i].radius >= (float) b.offset[a]
                     && splats[i].position[a] - splats[i].radius <= (float) (b.offset[a] + (std::int64_t) b.size[a] - 1);
            if (!in) continue;
            if (at >= b.ids.size() || b.ids[at] != i) return std::printf("FAIL: membership of splat %zu\n", i), 1;
            at++;
            covered[i] = 1;
        }
        if (at != b.ids.size()) return std::printf("FAIL: bin lists a splat that does not touch it\n"), 1;
        for (int a = 0; a < 3; a++)
        {
            const std::int64_t lo = b.offset[a], hi = b.offset[a] + (std::int64_t) b.size[a] - 1;   // cells lo..hi-1
            if (hi - lo > maxCells || hi <= lo || lo < 0 || hi > grid.numCells(a)) return std::printf("FAIL: extent\n"), 1;
            if (chunkCells > 0 && (lo / chunkCells != (hi - 1) / chunkCells || (std::uint32_t) (lo / chunkCells) != b.chunk[a]))
                return std::printf("FAIL: bin crosses a chunk\n"), 1;
            if (chunkCells == 0 && b.chunk[a] != 0) return std::printf("FAIL: chunk without --split\n"), 1;
        }
    }
    // bins do not overlap in cells: total cell volume <= grid volume, and every splat is in some bin
    double volume = 0.0;
    for (const mlsb::Bin &b : bins) volume += (double) (b.size[0] - 1) * (b.size[1] - 1) * (b.size[2] - 1);
    if (volume > (double) grid.numCells(0) * grid.numCells(1) * grid.numCells(2) + 0.5) return std::printf("FAIL: bins overlap\n"), 1;
    for (std::size_t i = 0; i < n; i++)
        if (!covered[i]) return std::printf("FAIL: splat %zu is in no bin\n", i), 1;
    std::printf("OK %zu bins %zu chunks %zu instances\n", bins.size(), chunks.size(), instances);
    return 0;
}
