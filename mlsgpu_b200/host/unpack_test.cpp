// unpackWireMesh (csrc/unpack.cpp) against a scalar restatement of latticeEmitKernel's vertex arithmetic, on
// random wire blobs; also reports the expansion rate of one thread.  "unpack_test: ok" or the failed check.
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../csrc/wire.h"

#define CHECK(cond) \
    do { if (!(cond)) { std::fprintf(stderr, "unpack_test: %s:%d: %s\n", __FILE__, __LINE__, #cond); std::exit(1); } } while (0)

int main()
{
    std::mt19937_64 rng(7);
    for (int round = 0; round < 4; round++)
    {
        const uint64_t nV = round == 0 ? 0 : round == 1 ? 1 : 250007 + round, nT = round == 0 ? 0 : 2 * nV + 3;
        const uint64_t nInt = nV - nV / 17, nExt = nV - nInt;
        mlsb::WireParams p;
        p.numVertices = nV;
        p.numTriangles = nT;
        p.numInternalVertices = nInt;
        p.zFirst = 8 * (uint32_t) round;
        p.gridOffset[0] = 1000003 % 1048000;
        p.gridOffset[1] = 77;
        p.gridOffset[2] = 524287;
        p.scale = 0.0021f;
        p.bias[0] = -1.25f;
        p.bias[1] = 3.0e-3f;
        p.bias[2] = 17.5f;
        std::vector<uint64_t> wire(nExt + nV + nT);
        std::vector<uint64_t> keys(nExt);
        std::vector<float> want(3 * nV);
        std::vector<uint32_t> tris(3 * nT);
        for (uint64_t i = 0; i < nExt; i++) wire[i] = keys[i] = rng();
        for (uint64_t i = 0; i < nV; i++)
        {
            const uint32_t x2 = rng() % 1023, y2 = rng() % 1023, z2 = rng() % 1023;
            // t in [0, 1] mostly, with the odd subnormal, negative zero and huge value thrown in
            float t = (float) ((rng() >> 11) * (1.0 / 9007199254740992.0));
            if (i % 1001 == 5) t = 1e-41f;
            if (i % 1001 == 6) t = -0.0f;
            if (i % 1001 == 7) t = 3e38f;
            uint32_t tb;
            std::memcpy(&tb, &t, 4);
            wire[nExt + i] = (uint64_t) tb | ((uint64_t) (x2 | (y2 << 11) | (z2 << 22)) << 32);
            const uint32_t c[3] = {x2, y2, z2};
            for (int a = 0; a < 3; a++)
            {
                const uint32_t base = (c[a] >> 1) + p.gridOffset[a] + (a == 2 ? p.zFirst : 0);
                volatile float pa = fmaf(t, (float) (c[a] & 1), (float) base);   // one rounding each, no vector unit
                volatile float out = fmaf(pa, p.scale, p.bias[a]);
                want[3 * i + a] = out;
            }
        }
        for (uint64_t i = 0; i < nT; i++)
        {
            uint64_t w = 0;
            for (int k = 0; k < 3; k++)
            {
                tris[3 * i + k] = k == 0 && i == 0 ? (1u << 21) - 1 : (uint32_t) (rng() % (nV ? nV : 1));
                w |= (uint64_t) tris[3 * i + k] << (21 * k);
            }
            wire[nExt + nV + i] = w;
        }
        std::vector<uint8_t> blob(8 * nExt + 12 * nV + 12 * nT + 8, 0xAB);
        const auto t0 = std::chrono::steady_clock::now();
        mlsb::unpackWireMesh(wire.data(), p, blob.data());
        const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        CHECK(std::memcmp(blob.data(), keys.data(), 8 * nExt) == 0);
        CHECK(std::memcmp(blob.data() + 8 * nExt, want.data(), 12 * nV) == 0);   // bit for bit, NaN and -0 included
        CHECK(std::memcmp(blob.data() + 8 * nExt + 12 * nV, tris.data(), 12 * nT) == 0);
        for (int k = 0; k < 8; k++) CHECK(blob[8 * nExt + 12 * nV + 12 * nT + k] == 0xAB);   // nothing written past the end
        if (round == 3)
            std::printf("unpack_test: %.2f ns per vertex+2 triangles (%.1f GB/s of blob written)\n", sec * 1e9 / (double) nV,
                        (double) (8 * nExt + 12 * nV + 12 * nT) / sec * 1e-9);
    }
    std::puts("unpack_test: ok");
    return 0;
}
