/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see orc_common.h).
 *
 * Host-side input conditioning that sits immediately before the device path.
 */
#include "orc_common.h"

/* FastPly::Reader::decode, src/fast_ply.cpp:334-350: radius clamp, smoothing, quality.
 * The quality expression is evaluated in double (1.0 / float) and rounded once. */
ORC_EXPORT void orc_condition_splat(OrcSplat *s, float smooth, float maxRadius)
{
    float r = s->radius < maxRadius ? s->radius : maxRadius;
    r *= smooth;
    s->radius = r;
    s->quality = (float) (1.0 / (double) (r * r));
}

/* Grid::worldToVertex (src/grid.cpp:99-106) + radius rescale, as applied by
 * BucketLoader::operator(), src/bucket_loader.cpp:76-85.  lo = fullGrid extent.first. */
ORC_EXPORT void orc_world_to_grid(OrcSplat *splats, uint32_t n, const float reference[3],
                                  float spacing, const int lo[3])
{
    float invSpacing = 1.0f / spacing;
    for (uint32_t i = 0; i < n; i++)
    {
        for (int a = 0; a < 3; a++)
            splats[i].position[a] = (splats[i].position[a] - reference[a]) * invSpacing - (float) lo[a];
        splats[i].radius *= invSpacing;
    }
}
