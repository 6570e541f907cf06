// gen_synth — the synthetic inputs of SURVEY.md 8(d) that are too large to generate in numpy:
//
//   C3  multi-scan terrain: K overlapping "scans" (rectangular patches with 10 % overlap) of a
//       height field z = sum of octaves of value noise over a 10 x 10 domain, independent noise of
//       0.25 x spacing per scan, normals from the analytic gradient, radius 1.5 x scan spacing,
//       one PLY per scan, seeds 3, 4, ...; --fit-grid = the mean sample spacing
//   C5  near-planar terrain (|z| <= 0.01 x extent) with 64 rectangular holes, a ragged outer rim and
//       200 isolated 500-point islands off to the side (to be pruned), seed 11
//
// Output: little-endian binary PLY, float32 x y z nx ny nz radius (src/fast_ply.cpp:185-188), points
// in Morton order of their position within a scan (the manual's coherence advice,
// doc/mlsgpu-user-manual.xml:211-216).  Counter-based randomness (a 64-bit mix of seed, scan and
// point index), so any point can be generated independently and the files do not depend on the
// number of threads.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace
{

inline std::uint64_t mix(std::uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

struct Rng
{
    std::uint64_t key, ctr = 0;
    Rng(std::uint64_t seed, std::uint64_t stream, std::uint64_t index) : key(mix(mix(seed) ^ mix(stream * 0x100000001B3ull + 7) ^ index)) {}
    double uniform() { return (double) (mix(key + (ctr++)) >> 11) * (1.0 / 9007199254740992.0); }
    double normal()
    {
        const double u = std::max(uniform(), 1e-300), v = uniform();
        return std::sqrt(-2.0 * std::log(u)) * std::cos(6.283185307179586 * v);
    }
};

// value noise on an integer lattice with smooth (quintic) interpolation, and its gradient
double lattice(std::uint64_t seed, std::int64_t ix, std::int64_t iy)
{
    return (double) (mix(seed ^ mix((std::uint64_t) ix * 0x9E3779B1ull + (std::uint64_t) iy * 0x85EBCA77ull)) >> 11) * (2.0 / 9007199254740992.0) - 1.0;
}

void valueNoise(std::uint64_t seed, double x, double y, double &v, double &dx, double &dy)
{
    const double fx = std::floor(x), fy = std::floor(y);
    const std::int64_t ix = (std::int64_t) fx, iy = (std::int64_t) fy;
    const double tx = x - fx, ty = y - fy;
    const double sx = tx * tx * tx * (tx * (tx * 6 - 15) + 10), sy = ty * ty * ty * (ty * (ty * 6 - 15) + 10);
    const double dsx = 30 * tx * tx * (tx * (tx - 2) + 1), dsy = 30 * ty * ty * (ty * (ty - 2) + 1);
    const double a = lattice(seed, ix, iy), b = lattice(seed, ix + 1, iy), c = lattice(seed, ix, iy + 1), d = lattice(seed, ix + 1, iy + 1);
    v = a + (b - a) * sx + (c - a) * sy + (a - b - c + d) * sx * sy;
    dx = ((b - a) + (a - b - c + d) * sy) * dsx;
    dy = ((c - a) + (a - b - c + d) * sx) * dsy;
}

// the terrain: `octaves` octaves starting at wavelength 2.5, amplitude halving as frequency doubles
void terrain(std::uint64_t seed, int octaves, double amplitude, double x, double y, double &z, double &zx, double &zy)
{
    z = zx = zy = 0.0;
    double freq = 0.4, amp = amplitude;
    for (int o = 0; o < octaves; o++, freq *= 2.0, amp *= 0.5)
    {
        double v, dx, dy;
        valueNoise(seed + 1000 * o, x * freq, y * freq, v, dx, dy);
        z += amp * v;
        zx += amp * dx * freq;
        zy += amp * dy * freq;
    }
}

struct Vertex
{
    float x, y, z, nx, ny, nz, radius;
};
static_assert(sizeof(Vertex) == 28, "PLY vertex is 7 floats");

std::uint64_t morton2(std::uint32_t x, std::uint32_t y)
{
    auto spread = [](std::uint64_t v) {
        v = (v | (v << 16)) & 0x0000FFFF0000FFFFull;
        v = (v | (v << 8)) & 0x00FF00FF00FF00FFull;
        v = (v | (v << 4)) & 0x0F0F0F0F0F0F0F0Full;
        v = (v | (v << 2)) & 0x3333333333333333ull;
        v = (v | (v << 1)) & 0x5555555555555555ull;
        return v;
    };
    return spread(x) | (spread(y) << 1);
}

void writePly(const std::string &path, std::vector<Vertex> &v, double x0, double y0, double extent)
{
    // Morton order of the quantised (x, y)
    std::vector<std::pair<std::uint64_t, std::uint32_t>> order(v.size());
#pragma omp parallel for schedule(static)
    for (std::int64_t i = 0; i < (std::int64_t) v.size(); i++)
    {
        const double qx = std::min(std::max((v[i].x - x0) / extent, 0.0), 1.0), qy = std::min(std::max((v[i].y - y0) / extent, 0.0), 1.0);
        order[i] = std::make_pair(morton2((std::uint32_t) (qx * 65535.0), (std::uint32_t) (qy * 65535.0)), (std::uint32_t) i);
    }
    std::sort(order.begin(), order.end());
    std::vector<Vertex> sorted(v.size());
#pragma omp parallel for schedule(static)
    for (std::int64_t i = 0; i < (std::int64_t) v.size(); i++) sorted[i] = v[order[i].second];
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot create " + path);
    std::fprintf(f, "ply\nformat binary_little_endian 1.0\ncomment mlsb200 gen_synth\nelement vertex %zu\n"
                    "property float32 x\nproperty float32 y\nproperty float32 z\nproperty float32 nx\nproperty float32 ny\n"
                    "property float32 nz\nproperty float32 radius\nend_header\n", sorted.size());
    if (!sorted.empty() && std::fwrite(sorted.data(), sizeof(Vertex), sorted.size(), f) != sorted.size())
        throw std::runtime_error("short write to " + path);
    std::fclose(f);
}

void fillTerrainPoint(Vertex &out, std::uint64_t surfaceSeed, int octaves, double amplitude, double x, double y, double noise,
                      double radius, Rng &rng)
{
    double z, zx, zy;
    terrain(surfaceSeed, octaves, amplitude, x, y, z, zx, zy);
    const double inv = 1.0 / std::sqrt(zx * zx + zy * zy + 1.0);
    const double nx = -zx * inv, ny = -zy * inv, nz = inv;
    const double d = noise * rng.normal();   // scanner noise along the normal
    out.x = (float) (x + d * nx);
    out.y = (float) (y + d * ny);
    out.z = (float) (z + d * nz);
    out.nx = (float) nx;
    out.ny = (float) ny;
    out.nz = (float) nz;
    out.radius = (float) radius;
}

}  // namespace

int main(int argc, char **argv)
{
    try
    {
        std::string config = "C3", prefix = "synth";
        std::uint64_t points = 1000000;
        int scans = 8;
        std::uint64_t seed = 0;
        for (int i = 1; i < argc; i++)
        {
            const std::string a = argv[i];
            auto value = [&]() -> std::string {
                if (i + 1 >= argc) throw std::invalid_argument("option " + a + " needs a value");
                return argv[++i];
            };
            if (a == "--config") config = value();
            else if (a == "--points") points = std::strtoull(value().c_str(), nullptr, 10);
            else if (a == "--scans") scans = std::atoi(value().c_str());
            else if (a == "--out") prefix = value();
            else if (a == "--seed") seed = std::strtoull(value().c_str(), nullptr, 10);
            else throw std::invalid_argument("unknown option " + a + " (--config C3|C5 --points N [--scans K] --out PREFIX [--seed S])");
        }
        if (config == "C3")
        {
            // K scans in a cols x rows arrangement over the 10 x 10 domain, each patch 10 % larger than its cell
            if (seed == 0) seed = 3;
            const double domain = 10.0;
            int cols = 1;
            while (cols * cols < scans) cols++;
            while (scans % cols != 0) cols++;
            const int rows = scans / cols;
            const std::uint64_t per = points / (std::uint64_t) scans;
            const double w = domain / cols * 1.1, h = domain / rows * 1.1;
            const double spacing = std::sqrt(w * h / (double) per);
            for (int s = 0; s < scans; s++)
            {
                const double cx = (s % cols + 0.5) * domain / cols, cy = (s / cols + 0.5) * domain / rows;
                const double x0 = cx - 0.5 * w, y0 = cy - 0.5 * h;
                std::vector<Vertex> v(per);
#pragma omp parallel for schedule(static)
                for (std::int64_t i = 0; i < (std::int64_t) per; i++)
                {
                    Rng rng(seed + (std::uint64_t) s, 1, (std::uint64_t) i);
                    const double x = x0 + w * rng.uniform(), y = y0 + h * rng.uniform();
                    fillTerrainPoint(v[i], 3 /* the surface is common to all scans */, 5, 0.35, x, y, 0.25 * spacing, 1.5 * spacing, rng);
                }
                char name[512];
                std::snprintf(name, sizeof(name), "%s_%02d.ply", prefix.c_str(), s);
                writePly(name, v, x0, y0, std::max(w, h));
            }
            // overlapping scans: the mean sample spacing of the union is a little below one scan's
            std::printf("config C3 points %llu scans %d spacing %.9g fit-grid %.9g\n", (unsigned long long) (per * scans), scans, spacing, spacing);
        }
        else if (config == "C5")
        {
            if (seed == 0) seed = 11;
            const double extent = 10.0;
            const std::uint64_t islandPoints = 200 * 500;
            const std::uint64_t sheet = points > islandPoints ? points - islandPoints : points;
            // 64 rectangular holes on a jittered 8 x 8 arrangement
            struct Hole { double x0, y0, x1, y1; };
            std::vector<Hole> holes;
            for (int k = 0; k < 64; k++)
            {
                Rng r(seed, 2, (std::uint64_t) k);
                const double cx = (k % 8 + 0.5 + 0.3 * (r.uniform() - 0.5)) * extent / 8, cy = (k / 8 + 0.5 + 0.3 * (r.uniform() - 0.5)) * extent / 8;
                const double hw = 0.12 + 0.18 * r.uniform(), hh = 0.12 + 0.18 * r.uniform();
                holes.push_back(Hole{cx - hw, cy - hh, cx + hw, cy + hh});
            }
            // accepted fraction of the square: measured on a sample, to hit the requested count
            auto accept = [&](double x, double y) {
                // ragged rim: the sheet ends 0.15 +- 0.1 inside the square, with a ripple along the edge
                const double rim = 0.15 + 0.1 * std::sin(7.0 * x) * std::sin(5.0 * y) + 0.05 * std::sin(23.0 * (x + y));
                if (x < rim || y < rim || x > extent - rim || y > extent - rim) return false;
                for (const Hole &h : holes)
                    if (x > h.x0 && x < h.x1 && y > h.y0 && y < h.y1) return false;
                return true;
            };
            std::uint64_t hits = 0;
            const std::uint64_t probe = 2000000;
            for (std::uint64_t i = 0; i < probe; i++)
            {
                Rng r(seed, 3, i);
                hits += accept(extent * r.uniform(), extent * r.uniform());
            }
            const double fraction = (double) hits / (double) probe;
            const std::uint64_t candidates = (std::uint64_t) ((double) sheet / fraction);
            const double spacing = std::sqrt(extent * extent * fraction / (double) sheet);
            std::vector<Vertex> v(candidates);
            std::vector<unsigned char> keep(candidates);
#pragma omp parallel for schedule(static)
            for (std::int64_t i = 0; i < (std::int64_t) candidates; i++)
            {
                Rng rng(seed, 1, (std::uint64_t) i);
                const double x = extent * rng.uniform(), y = extent * rng.uniform();
                keep[i] = accept(x, y);
                if (keep[i]) fillTerrainPoint(v[i], seed, 3, 0.01 * extent / 1.75, x, y, 0.25 * spacing, 1.5 * spacing, rng);
            }
            std::vector<Vertex> out;
            out.reserve(sheet + islandPoints + 1024);
            for (std::uint64_t i = 0; i < candidates; i++)
                if (keep[i]) out.push_back(v[i]);
            v.clear();
            v.shrink_to_fit();
            // 200 islands of 500 points: discs of radius 12 spacings on a 20 x 10 arrangement beside the sheet (x > extent + 0.5)
            for (int k = 0; k < 200; k++)
            {
                const double cx = extent + 0.5 + (k % 10 + 0.5) * 0.25, cy = (k / 10 + 0.5) * extent / 20;
                for (int j = 0; j < 500; j++)
                {
                    Rng rng(seed, 4, (std::uint64_t) k * 500 + j);
                    const double rr = 12.0 * spacing * std::sqrt(rng.uniform()), t = 6.283185307179586 * rng.uniform();
                    Vertex p;
                    p.x = (float) (cx + rr * std::cos(t));
                    p.y = (float) (cy + rr * std::sin(t));
                    p.z = (float) (0.25 * spacing * rng.normal());
                    p.nx = 0.0f;
                    p.ny = 0.0f;
                    p.nz = 1.0f;
                    p.radius = (float) (1.5 * spacing);
                    out.push_back(p);
                }
            }
            writePly(prefix + "_00.ply", out, 0.0, 0.0, extent + 3.5);
            std::printf("config C5 points %zu islands 200 holes 64 spacing %.9g fit-grid %.9g\n", out.size(), spacing, spacing);
        }
        else
            throw std::invalid_argument("unknown config " + config);
        return 0;
    }
    catch (const std::exception &e)
    {
        std::fprintf(stderr, "gen_synth: %s\n", e.what());
        return 1;
    }
}
