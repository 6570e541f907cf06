// Host half of the compact read-back (MLSB_READBACK=compact, see compactMeshKernel in marching.cu): expands the
// wire blob of one shipped block — [external keys][vertices][triangles], one 64-bit word each — into the
// HostKeyMesh blob the sink expects (src/mesh.cpp:51-60: [uint64 keys][float[3] vertices][uint32[3] triangles]).
//
//   vertex word    low 32 bits: the interpolation parameter t (float bits)
//                  high 32 bits: the edge's point on the half-cell lattice, x2 | y2 << 11 | z2 << 22
//   triangle word  three 21-bit vertex indices
//
// A vertex is rebuilt with exactly the operations of latticeEmitKernel (kernels/marching.cl:130-138 and
// kernels/scale_bias.cl:32-40): p = fma(t, d, a + gridOffset), out = fma(p, scale, bias), each a single IEEE
// fused multiply-add in round-to-nearest — std::fmaf is one on every platform, so the result is bit-identical to
// the device's whichever version below runs.  The loops are written for the vectoriser; the AVX2 + FMA version is
// picked at run time on CPUs that have both, the portable one calls libm.
#include <cmath>
#include <cstdint>
#include <cstring>

#include "wire.h"

namespace mlsb
{

namespace
{

// The two loops, as a macro so that they can be compiled twice: for the baseline ISA (fmaf is a libm call there)
// and for AVX2 + FMA (the vectoriser turns them into 8-wide code with vfmadd).
#define MLSB_UNPACK_BODY(SUFFIX, ATTR)                                                                                     \
    ATTR void unpackVertices##SUFFIX(const uint64_t *__restrict__ wire, uint64_t n, uint32_t zFirst, const uint32_t gridOffset[3], \
                                     float scale, const float bias[3], float *__restrict__ out)                           \
    {                                                                                                                      \
        const uint32_t gx = gridOffset[0], gy = gridOffset[1], gz = gridOffset[2] + zFirst;                               \
        const float bx = bias[0], by = bias[1], bz = bias[2];                                                             \
        for (uint64_t i = 0; i < n; i++)                                                                                   \
        {                                                                                                                  \
            const uint64_t w = wire[i];                                                                                    \
            const uint32_t tb = (uint32_t) w, l = (uint32_t) (w >> 32);                                                    \
            float t;                                                                                                       \
            std::memcpy(&t, &tb, sizeof(t));                                                                               \
            const uint32_t x2 = l & 2047u, y2 = (l >> 11) & 2047u, z2 = l >> 22;                                           \
            /* unsigned sums, then one conversion: as (float) (ax + P.gridOffset.x) on the device */                      \
            const float ax = (float) ((x2 >> 1) + gx), ay = (float) ((y2 >> 1) + gy), az = (float) ((z2 >> 1) + gz);      \
            const float px = std::fmaf(t, (float) (x2 & 1u), ax);                                                         \
            const float py = std::fmaf(t, (float) (y2 & 1u), ay);                                                         \
            const float pz = std::fmaf(t, (float) (z2 & 1u), az);                                                         \
            out[3 * i + 0] = std::fmaf(px, scale, bx);                                                                    \
            out[3 * i + 1] = std::fmaf(py, scale, by);                                                                    \
            out[3 * i + 2] = std::fmaf(pz, scale, bz);                                                                    \
        }                                                                                                                  \
    }                                                                                                                      \
    ATTR void unpackTriangles##SUFFIX(const uint64_t *__restrict__ wire, uint64_t n, uint32_t *__restrict__ out)          \
    {                                                                                                                      \
        for (uint64_t i = 0; i < n; i++)                                                                                   \
        {                                                                                                                  \
            const uint64_t w = wire[i];                                                                                    \
            out[3 * i + 0] = (uint32_t) (w & 0x1FFFFFu);                                                                   \
            out[3 * i + 1] = (uint32_t) ((w >> 21) & 0x1FFFFFu);                                                           \
            out[3 * i + 2] = (uint32_t) ((w >> 42) & 0x1FFFFFu);                                                           \
        }                                                                                                                  \
    }

MLSB_UNPACK_BODY(Base, )
#if defined(__x86_64__) && defined(__GNUC__)
#define MLSB_UNPACK_HAVE_AVX2 1
MLSB_UNPACK_BODY(Avx2, __attribute__((target("avx2,fma"))))
#endif

}  // namespace

void unpackWireMesh(const uint64_t *wire, const WireParams &p, void *hostBlob)
{
    const uint64_t numExternal = p.numVertices - p.numInternalVertices;
    uint8_t *keys = static_cast<uint8_t *>(hostBlob);
    float *vertices = reinterpret_cast<float *>(keys + numExternal * sizeof(uint64_t));
    uint32_t *triangles = reinterpret_cast<uint32_t *>(vertices + 3 * p.numVertices);
    if (numExternal) std::memcpy(keys, wire, numExternal * sizeof(uint64_t));
#ifdef MLSB_UNPACK_HAVE_AVX2
    static const bool fast = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma");
    if (fast)
    {
        unpackVerticesAvx2(wire + numExternal, p.numVertices, p.zFirst, p.gridOffset, p.scale, p.bias, vertices);
        unpackTrianglesAvx2(wire + numExternal + p.numVertices, p.numTriangles, triangles);
        return;
    }
#endif
    unpackVerticesBase(wire + numExternal, p.numVertices, p.zFirst, p.gridOffset, p.scale, p.bias, vertices);
    unpackTrianglesBase(wire + numExternal + p.numVertices, p.numTriangles, triangles);
}

}  // namespace mlsb
