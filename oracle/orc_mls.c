/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see orc_common.h).
 *
 * Restatement of kernels/mls.cl (processCorners and helpers) and the launch geometry
 * of MlsFunctor::enqueue, src/mls.cpp:101-144.
 *
 * Contraction choices (mls.cl leaves FP_CONTRACT at the OpenCL default, ON, so the
 * reference's own rounding depends on its OpenCL compiler; north_star tolerance for the
 * field is 1e-5 relative).  The accumulation is taken in the fully fused form that an
 * OpenCL compiler is permitted to produce:
 *      sumW   += w
 *      sumWp   = fma(w, p, sumWp)            (kernels/mls.cl:131,134)
 *      sumWn   = fma(w, n, sumWn)            (kernels/mls.cl:132,135)
 *      sumWpp  = fma(w, pp, sumWpp)          (kernels/mls.cl:136)
 *      sumWpn  = fma(w, dot3(n, p), sumWpn)  (kernels/mls.cl:137; (w n).p = w (n.p))
 * Building with -DORC_MLS_AS_WRITTEN gives the literal unfused statement order instead
 * (used by tests to show that both forms agree to well within the tolerance).
 * Everything else is evaluated exactly as written, one rounding per operator.
 */
#include "orc_common.h"
#include <math.h>
#include <float.h>
#include <limits.h>
#include <string.h>

#define RADIUS_CUTOFF 0.99f   /* kernels/mls.cl:36 */
#define HITS_CUTOFF 4         /* kernels/mls.cl:37 */
#define WGS 8                 /* src/mls.cpp:53 */

typedef struct
{
    float sumWpp, sumWpn;
    float sumWp[3], sumWn[3];
    float sumW;
    uint32_t hits;
} SphereFit;   /* kernels/mls.cl:69-77; PlaneFit (:79-86) is the same minus sumWpn */

typedef struct
{
    float a, b[3], c, qDen, b2;
} Sphere;      /* kernels/mls.cl:88-95 */

/* kernels/mls.cl:175-189 */
ORC_EXPORT void orc_decode(uint32_t code, int out[3])
{
    int x = 0, y = 0, z = 0;
    uint32_t scale = 1;
    while (code >= scale)
    {
        x += code & scale;
        y += (code >> 1) & scale;
        z += (code >> 2) & scale;
        code >>= 2;
        scale <<= 1;
    }
    out[0] = x; out[1] = y; out[2] = z;
}

static void fit_init(SphereFit *sf)
{
    memset(sf, 0, sizeof(*sf));
}

/* kernels/mls.cl:127-148 (sphereFitAdd / planeFitAdd) */
static inline void fit_add(SphereFit *sf, float w, const float p[3], float pp, const float n[3])
{
#ifdef ORC_MLS_AS_WRITTEN
    float wp[3] = { w * p[0], w * p[1], w * p[2] };
    float wn[3] = { w * n[0], w * n[1], w * n[2] };
    sf->sumW += w;
    for (int i = 0; i < 3; i++) sf->sumWp[i] += wp[i];
    for (int i = 0; i < 3; i++) sf->sumWn[i] += wn[i];
    sf->sumWpp += w * pp;
    sf->sumWpn += orc_dot3(wn, p);
#else
    sf->sumW += w;
    for (int i = 0; i < 3; i++) sf->sumWp[i] = __builtin_fmaf(w, p[i], sf->sumWp[i]);
    for (int i = 0; i < 3; i++) sf->sumWn[i] = __builtin_fmaf(w, n[i], sf->sumWn[i]);
    sf->sumWpp = __builtin_fmaf(w, pp, sf->sumWpp);
    sf->sumWpn = __builtin_fmaf(w, orc_dot3(n, p), sf->sumWpn);
#endif
    sf->hits++;
}

/* kernels/mls.cl:210-229 */
static void fit_sphere(const SphereFit *sf, Sphere *out)
{
    float invSumW = 1.0f / sf->sumW;
    float m[3] = { sf->sumWp[0] * invSumW, sf->sumWp[1] * invSumW, sf->sumWp[2] * invSumW };
    float qNum = sf->sumWpn - orc_dot3(m, sf->sumWn);
    float qDen = sf->sumWpp - orc_dot3(m, sf->sumWp);
    float q = qNum / qDen;
    if (fabsf(qDen) < (4 * FLT_EPSILON) * (float) sf->hits * fabsf(sf->sumWpp) || !orc_isfinite(q))
        q = 0.0f;
    float a = 0.5f * q;
    float b[3];
    for (int i = 0; i < 3; i++)
        b[i] = (sf->sumWn[i] - q * sf->sumWp[i]) * invSumW;
    out->a = a;
    memcpy(out->b, b, sizeof(b));
    out->c = (-a * sf->sumWpp - orc_dot3(b, sf->sumWp)) * invSumW;
    out->qDen = qDen;
    out->b2 = orc_dot3(b, b);
}

/* kernels/mls.cl:237-248 */
ORC_EXPORT float orc_solve_quadratic(float a, float b, float c)
{
    float bdet = b + sqrtf(b * b - 4.0f * a * c);
    float x = -2.0f * c / bdet;
    if (!orc_isfinite(x))
        x = bdet / (-2.0f * a);
    return orc_isfinite(x) ? x : orc_nan();
}

/* kernels/mls.cl:254-258; params = {b.x, b.y, b.z, a, c} as in testProjectDistOriginSphere (:446-450).
 * OpenCL length() taken as sqrtf(dot3(b, b)). */
ORC_EXPORT float orc_project_dist_origin_sphere(const float params[5])
{
    float b[3] = { params[0], params[1], params[2] };
    return -orc_solve_quadratic(params[3], sqrtf(orc_dot3(b, b)), params[4]);
}

/* testFitSphere shim, kernels/mls.cl:452-469: out = {b.x, b.y, b.z, a, c} */
ORC_EXPORT void orc_fit_sphere(const OrcSplat *in, uint32_t nsplats, float out[5])
{
    SphereFit sf;
    fit_init(&sf);
    for (uint32_t i = 0; i < nsplats; i++)
        fit_add(&sf, in[i].quality, in[i].position, orc_dot3(in[i].position, in[i].position), in[i].normal);
    Sphere s;
    fit_sphere(&sf, &s);
    out[0] = s.b[0]; out[1] = s.b[1]; out[2] = s.b[2]; out[3] = s.a; out[4] = s.c;
}

/* src/mls.cpp:137-144 */
ORC_EXPORT float orc_boundary_factor(float limit)
{
    const float pi = 3.14159265358979323846f;
    const float boundaryScale = (sqrtf(6.0f) * 512) / (693 * pi);
    const float gamma = boundaryScale * limit;
    return 1.0f - gamma * gamma;
}

/* Evaluate one corner: body of processCorners for a single work-item,
 * kernels/mls.cl:322-424.  pos = start[code] (>= 0).  shape: 0 sphere, 1 plane. */
static float corner_value(const OrcSplat *splats, const int32_t *commands, int32_t pos,
                          const float coord[3], float boundaryFactor, int shape,
                          uint64_t *pairTests, uint64_t *hitCount)
{
    SphereFit fit;
    fit_init(&fit);
    uint64_t tests = 0;
    /* The kernel stages <=256 ids per batch and never lets a batch span two ranges
     * (:336-358); per work-item the visit order is simply the chain order. */
    while (pos >= 0)
    {
        int32_t end = commands[pos++];
        for (; pos < end; pos++)
        {
            int32_t splatId = commands[pos];
            const OrcSplat *s = &splats[splatId];
            float p[3] = { s->position[0] - coord[0], s->position[1] - coord[1], s->position[2] - coord[2] };
            float pp = orc_dot3(p, p);
            float d = pp * s->radius;        /* .w holds 1/r^2 */
            tests++;
            if (d < RADIUS_CUTOFF)
            {
                float w = 1.0f - d;
                w *= w;
                w *= w;
                w *= s->quality;
                fit_add(&fit, w, p, pp, s->normal);
            }
        }
        pos = commands[end];
    }
    *pairTests += tests;
    *hitCount += fit.hits;

    float f = orc_nan();
    if (fit.hits >= HITS_CUTOFF)
    {
        if (shape == 0)
        {
            Sphere sphere;
            fit_sphere(&fit, &sphere);
            /* projectOriginSphere, :263-267 */
            float l = orc_solve_quadratic(sphere.a * sphere.b2, sphere.b2, sphere.c);
            float a[3] = { l * sphere.b[0], l * sphere.b[1], l * sphere.b[2] };
            float aa = orc_dot3(a, a);
            if (aa < 3.0f)
            {
                float rhs = (fit.sumWpp - 2 * orc_dot3(fit.sumWp, a) + fit.sumW * aa);
                if (sphere.qDen > boundaryFactor * rhs)
                    f = -orc_dot3(sphere.b, a) * (1.0f / sqrtf(sphere.b2));  /* half_rsqrt */
            }
        }
        else
        {
            /* fitPlane :198-203; normalize(v) taken as v * (1/sqrt(dot3(v,v))) */
            float mean[3] = { fit.sumWp[0] / fit.sumW, fit.sumWp[1] / fit.sumW, fit.sumWp[2] / fit.sumW };
            float rl = 1.0f / sqrtf(orc_dot3(fit.sumWn, fit.sumWn));
            float normal[3] = { fit.sumWn[0] * rl, fit.sumWn[1] * rl, fit.sumWn[2] * rl };
            float dist = -orc_dot3(normal, mean);
            /* projectOriginPlane :277-280 */
            float a[3] = { normal[0] * -dist, normal[1] * -dist, normal[2] * -dist };
            float aa = orc_dot3(a, a);
            if (aa < 3.0f)
            {
                float qDen = fit.sumWpp - orc_dot3(mean, fit.sumWp);
                float rhs = (fit.sumWpp - 2 * orc_dot3(fit.sumWp, a) + fit.sumW * aa);
                if (qDen > boundaryFactor * rhs)
                    f = dist;
            }
        }
    }
    return f;
}

/*
 * MlsFunctor::enqueue + processCorners.
 * image: float array with row pitch `pitch` (>= roundUp(width, 8)); corner (x, y, z) is stored at
 *        image[(y + z*zStride + zBias) * pitch + x]  (Marching::ImageParams, src/marching.h:167-180).
 * Work-groups cover roundUp(width,8) x roundUp(height,8) x the z range
 * [zFirst, zFirst + 8*ceil((zLast-zFirst+1)/8)), exactly as src/mls.cpp:119-134.
 * stats (optional, 2 entries): [0] += pair tests T, [1] += hits H (SURVEY 8d work model).
 */
ORC_EXPORT void orc_mls_process_corners(float *image, uint32_t pitch,
                                        const OrcSplat *splats, const int32_t *commands,
                                        const int32_t *start, uint32_t startShift,
                                        const int offset[3],
                                        uint32_t width, uint32_t height,
                                        uint32_t zStride, int32_t zBias,
                                        uint32_t zFirst, uint32_t zLast,
                                        float boundaryFactor, int shape, uint64_t *stats)
{
    uint32_t bx = (width + WGS - 1) / WGS, by = (height + WGS - 1) / WGS;
    uint32_t bz = (zLast - zFirst + 1 + WGS - 1) / WGS;
    uint64_t totalTests = 0, totalHits = 0;
    long nblocks = (long) bx * by * bz;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : totalTests, totalHits)
    for (long blk = 0; blk < nblocks; blk++)
    {
        uint32_t gx = (uint32_t) (blk % bx), gy = (uint32_t) ((blk / bx) % by), gz = (uint32_t) (blk / ((long) bx * by));
        int wid[3] = { (int) (gx * WGS), (int) (gy * WGS), (int) (gz * WGS + zFirst) };
        uint32_t code = orc_make_code(wid[0], wid[1], wid[2]) >> startShift;
        int32_t pos = start[code];
        for (uint32_t lid = 0; lid < WGS * WGS * WGS; lid++)
        {
            int l3[3];
            orc_decode(lid, l3);
            float f = orc_nan();
            if (pos >= 0)
            {
                float coord[3] = { (float) (wid[0] + l3[0] + offset[0]),
                                   (float) (wid[1] + l3[1] + offset[1]),
                                   (float) (wid[2] + l3[2] + offset[2]) };
                f = corner_value(splats, commands, pos, coord, boundaryFactor, shape,
                                 &totalTests, &totalHits);
            }
            int ox = wid[0] + l3[0];
            long oy = (long) (wid[1] + l3[1]) + (long) (wid[2] + l3[2]) * zStride + zBias;
            image[oy * pitch + ox] = f;
        }
    }
    if (stats)
    {
        stats[0] += totalTests;
        stats[1] += totalHits;
    }
}
