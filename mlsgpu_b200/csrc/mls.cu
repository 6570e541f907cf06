// Per-corner MLS field: MlsFunctor::enqueue (src/mls.cpp:101-135) and processCorners
// (kernels/mls.cl:299-433), sphere and plane fits.
//
// One CTA per HALF of the reference's 8x8x8 block of corners (8 x 8 x 4, see MLSB_MLS_CTA_Z): 128 threads that each
// own TWO corners sharing (y, z) (MLSB_MLS_SLOTS = 2):  dot3(p,p) = fma(px,px, fma(py,py, pz*pz))
// (kernels/mls.cl:105-108), so the inner fma(py,py,pz*pz) — and py, pz themselves — are computed
// once per splat per thread and reused by both corners, with bit-identical results.  Each warp owns
// a 4x4x4 brick of the block, walks the block's command chain on its own, and pair-tests only the
// splats whose ball can reach its brick.
//
// Rounding contract (shared with oracle/orc_mls.c): every operation is written with an
// explicit round-to-nearest intrinsic, fused only where the oracle fuses; the translation
// unit is also compiled with -fmad=false.  The field is bit-identical to the oracle's.
#include "common.cuh"
#include <cfloat>
#include <climits>
#include <cstdlib>
#include <cstring>

namespace mlsb
{

namespace
{

#ifndef MLSB_MLS_SLOTS
#define MLSB_MLS_SLOTS 2
#endif
#ifndef MLSB_MLS_VARIANT
#define MLSB_MLS_VARIANT 3
#endif
// CTAs per SM the register allocation aims for: 4 (64 registers, no spills) measures the same as 3
// (72 registers) when the kernel has the GPU to itself and 3 % better when it shares the GPU with
// the other workers' kernels; 5 (47 registers, spills) is 25 % slower
#ifndef MLSB_MLS_MINBLOCKS
#define MLSB_MLS_MINBLOCKS 4
#endif
constexpr int kSlots = MLSB_MLS_SLOTS;                      // corners per thread (sharing y and z)
// Corners per CTA along y and z (x is always 8): the warps of a CTA walk the same chain but finish at different
// times (a brick far from the surface keeps nothing), and a CTA holds its registers until its slowest warp is done —
// smaller CTAs give the slots back sooner.  Measured on C2 (ms per 70-bin step alone / bench Msplats/s):
// 8 x 8 x 8 = one octree block per CTA (8 warps) 19.40 / 365.8; 8 x 8 x 4 (4 warps, 8 CTAs per SM) 18.33 / 377.5;
// 8 x 4 x 4 (2 warps) 20.07 / 352.9 — twice the CTAs to launch for the empty blocks, and each re-reads the chain head.
#ifndef MLSB_MLS_CTA_Y
#define MLSB_MLS_CTA_Y 8
#endif
#ifndef MLSB_MLS_CTA_Z
#define MLSB_MLS_CTA_Z 4
#endif
constexpr int kCtaY = MLSB_MLS_CTA_Y, kCtaZ = MLSB_MLS_CTA_Z;
static_assert((kCtaY == 4 || kCtaY == 8) && (kCtaZ == 4 || kCtaZ == 8), "a CTA is 8 x {4,8} x {4,8} corners");
constexpr int kBrickX = 2 * kSlots;            // a warp covers a kBrickX x 4 x 4 brick
constexpr int kFullThreads = 512 / kSlots;     // threads for a whole 8x8x8 block of corners
constexpr int kThreads = kFullThreads * kCtaY * kCtaZ / 64;
#ifdef MLSB_MLS_CTAS   // A/B builds: CTAs per SM at the actual CTA size
constexpr int kMinBlocks = MLSB_MLS_CTAS;
#else
constexpr int kMinBlocks = MLSB_MLS_MINBLOCKS * (kFullThreads / kThreads);
#endif
constexpr int kBatch = 32;           // splats fetched per warp per batch
constexpr float kRadiusCutoff = 0.99f;   // kernels/mls.cl:36
constexpr int kHitsCutoff = 4;           // kernels/mls.cl:37

struct MlsParams
{
    const mlsb_splat *splats;
    const int32_t *commands;
    const int32_t *start;
    const float *thresholds;    // per splat hit threshold on |p|^2 (hitThreshold(), common.cuh) or nullptr
    float *field;
    uint32_t pitch;
    uint32_t startShift;
    int3 offset;
    uint32_t zStride;
    int32_t zBias;
    uint32_t zFirst;
    float boundaryFactor;
    unsigned long long *work;   // [0] pair tests, [1] hits (COUNT builds only)
    uint8_t *blockActive;       // optional [bz][by][bx]: 1 where the block has splats (hint for marching)
    uint32_t activeDimX, activeDimY;
};

__device__ __forceinline__ uint32_t spread3(uint32_t v)
{
    v &= 0x3FF;
    v = (v | (v << 16)) & 0x030000FF;
    v = (v | (v << 8)) & 0x0300F00F;
    v = (v | (v << 4)) & 0x030C30C3;
    v = (v | (v << 2)) & 0x09249249;
    return v;
}

__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz)
{
    return __fmaf_rn(ax, bx, __fmaf_rn(ay, by, __fmul_rn(az, bz)));
}

__device__ __forceinline__ bool finite(float x) { return (__float_as_uint(x) & 0x7F800000u) != 0x7F800000u; }

struct Accum
{
    float sumW, sumWpx, sumWpy, sumWpz, sumWnx, sumWny, sumWnz, sumWpp, sumWpn;
    uint32_t hits;
};

// kernels/mls.cl:237-248
__device__ __forceinline__ float solveQuadratic(float a, float b, float c)
{
    float det = __fsub_rn(__fmul_rn(b, b), __fmul_rn(__fmul_rn(4.0f, a), c));
    float bdet = __fadd_rn(b, __fsqrt_rn(det));
    float x = __fdiv_rn(__fmul_rn(-2.0f, c), bdet);
    if (!finite(x)) x = __fdiv_rn(bdet, __fmul_rn(-2.0f, a));
    return finite(x) ? x : __uint_as_float(0x7FC00000u);
}

// kernels/mls.cl:394-424 (+ fitSphere :210-229, projectOriginSphere :263-267, fitPlane :198-203)
template <int SHAPE> __device__ __forceinline__ float finish(const Accum &f, float boundaryFactor)
{
    float result = __uint_as_float(0x7FC00000u);
    if (f.hits < kHitsCutoff) return result;
    if (SHAPE == MLSB_SHAPE_SPHERE)
    {
        float invSumW = __fdiv_rn(1.0f, f.sumW);
        float mx = __fmul_rn(f.sumWpx, invSumW), my = __fmul_rn(f.sumWpy, invSumW), mz = __fmul_rn(f.sumWpz, invSumW);
        float qNum = __fsub_rn(f.sumWpn, dot3(mx, my, mz, f.sumWnx, f.sumWny, f.sumWnz));
        float qDen = __fsub_rn(f.sumWpp, dot3(mx, my, mz, f.sumWpx, f.sumWpy, f.sumWpz));
        float q = __fdiv_rn(qNum, qDen);
        float lim = __fmul_rn(__fmul_rn(4 * FLT_EPSILON, (float) f.hits), fabsf(f.sumWpp));
        if (fabsf(qDen) < lim || !finite(q)) q = 0.0f;
        float a = __fmul_rn(0.5f, q);
        float bx = __fmul_rn(__fsub_rn(f.sumWnx, __fmul_rn(q, f.sumWpx)), invSumW);
        float by = __fmul_rn(__fsub_rn(f.sumWny, __fmul_rn(q, f.sumWpy)), invSumW);
        float bz = __fmul_rn(__fsub_rn(f.sumWnz, __fmul_rn(q, f.sumWpz)), invSumW);
        float c = __fmul_rn(__fsub_rn(__fmul_rn(-a, f.sumWpp), dot3(bx, by, bz, f.sumWpx, f.sumWpy, f.sumWpz)), invSumW);
        float b2 = dot3(bx, by, bz, bx, by, bz);
        float l = solveQuadratic(__fmul_rn(a, b2), b2, c);
        float ax = __fmul_rn(l, bx), ay = __fmul_rn(l, by), az = __fmul_rn(l, bz);
        float aa = dot3(ax, ay, az, ax, ay, az);
        if (aa < 3.0f)
        {
            float rhs = __fadd_rn(__fsub_rn(f.sumWpp, __fmul_rn(2.0f, dot3(f.sumWpx, f.sumWpy, f.sumWpz, ax, ay, az))),
                                  __fmul_rn(f.sumW, aa));
            if (qDen > __fmul_rn(boundaryFactor, rhs))
                result = __fmul_rn(-dot3(bx, by, bz, ax, ay, az), __fdiv_rn(1.0f, __fsqrt_rn(b2)));
        }
    }
    else
    {
        float mx = __fdiv_rn(f.sumWpx, f.sumW), my = __fdiv_rn(f.sumWpy, f.sumW), mz = __fdiv_rn(f.sumWpz, f.sumW);
        float rl = __fdiv_rn(1.0f, __fsqrt_rn(dot3(f.sumWnx, f.sumWny, f.sumWnz, f.sumWnx, f.sumWny, f.sumWnz)));
        float nx = __fmul_rn(f.sumWnx, rl), ny = __fmul_rn(f.sumWny, rl), nz = __fmul_rn(f.sumWnz, rl);
        float dist = -dot3(nx, ny, nz, mx, my, mz);
        float ax = __fmul_rn(nx, -dist), ay = __fmul_rn(ny, -dist), az = __fmul_rn(nz, -dist);
        float aa = dot3(ax, ay, az, ax, ay, az);
        if (aa < 3.0f)
        {
            float qDen = __fsub_rn(f.sumWpp, dot3(mx, my, mz, f.sumWpx, f.sumWpy, f.sumWpz));
            float rhs = __fadd_rn(__fsub_rn(f.sumWpp, __fmul_rn(2.0f, dot3(f.sumWpx, f.sumWpy, f.sumWpz, ax, ay, az))),
                                  __fmul_rn(f.sumW, aa));
            if (qDen > __fmul_rn(boundaryFactor, rhs)) result = dist;
        }
    }
    return result;
}

// PRE: the octree build left every splat's hit threshold in P.thresholds; otherwise (buffers bound by
// mlsb_mls_set_buffers) the kernel derives it from 1/r^2 for every splat it keeps.
template <int SHAPE, bool COUNT, bool PRE>
__global__ void __launch_bounds__(kThreads, kMinBlocks) processCornersKernel(MlsParams P)
{
    // per-warp staging buffers for the splats that survive the cull
    __shared__ float4 sPosT[kThreads / 32][kBatch];    // x, y, z, hit threshold on |p|^2
    __shared__ float4 sNormQ[kThreads / 32][kBatch];   // normal, quality
    __shared__ float sInvR[kThreads / 32][kBatch];     // 1/r^2

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31, quad = tid >> 5;
    // corner layout inside the block: the warp owns a kBrickX x 4 x 4 brick; a thread owns the
    // corners x = lx + 2j (j = slot), with y and z fixed
    constexpr int kBricksX = 8 / kBrickX, kBricksY = kCtaY / 4;
    const int qx = kBrickX * (int) (quad % kBricksX), qy = 4 * (int) ((quad / kBricksX) % kBricksY), qz = 4 * (int) (quad / (kBricksY * kBricksX));
    const int lx = qx + (int) (lane & 1);
    const int ly = qy + (int) ((lane >> 1) & 3);
    const int lz = qz + (int) ((lane >> 3) & 3);
    const int wx = blockIdx.x * kMlsWgs, wy = blockIdx.y * kCtaY, wz = blockIdx.z * kCtaZ + (int) P.zFirst;

    // (the code of any corner of the 8x8x8 block: the low three bits of each coordinate are shifted out)
    const uint32_t code = (spread3(wx) | (spread3(wy) << 1) | (spread3(wz) << 2)) >> P.startShift;
    int32_t pos = P.start[code];
    if (P.blockActive != nullptr && tid == 0)
        P.blockActive[((size_t) (wz >> 3) * P.activeDimY + (wy >> 3)) * P.activeDimX + blockIdx.x] = pos >= 0;

    float out[kSlots];
#pragma unroll
    for (int j = 0; j < kSlots; j++) out[j] = __uint_as_float(0x7FC00000u);

    if (pos >= 0)
    {
        const float cy = (float) (wy + ly + P.offset.y);
        const float cz = (float) (wz + lz + P.offset.z);
        float cx[kSlots];
#pragma unroll
        for (int j = 0; j < kSlots; j++) cx[j] = (float) (wx + lx + 2 * j + P.offset.x);

        Accum acc[kSlots];
#pragma unroll
        for (int j = 0; j < kSlots; j++)
        {
            acc[j].sumW = acc[j].sumWpx = acc[j].sumWpy = acc[j].sumWpz = 0.0f;
            acc[j].sumWnx = acc[j].sumWny = acc[j].sumWnz = acc[j].sumWpp = acc[j].sumWpn = 0.0f;
            acc[j].hits = 0;
        }
        unsigned long long tests = 0;

        // bounding box of this warp's brick of corners (exact small integers)
        const float bx0 = (float) (wx + qx + P.offset.x), bx1 = bx0 + (float) (kBrickX - 1);
        const float by0 = (float) (wy + qy + P.offset.y), by1 = by0 + 3.0f;
        const float bz0 = (float) (wz + qz + P.offset.z), bz1 = bz0 + 3.0f;
        float4 *const wPosT = sPosT[quad];
        float4 *const wNormQ = sNormQ[quad];
        float *const wInvR = sInvR[quad];

        // ---- every warp walks the chain on its own (no block barriers): 32 splats at a time are
        // fetched into registers, culled against the warp's brick, and the survivors are packed
        // into the warp's shared-memory buffer in chain order.  Unlike the reference (:336-358) a
        // batch may span several ranges; the per-corner visit order is still the chain order. ----
        int32_t end = P.commands[pos++];
        float4 rPos = make_float4(0.f, 0.f, 0.f, 0.f), rNq = rPos;
        float rThr = 0.0f;
        bool rHave = false;
        int nextCount = 0;
        // one-deep software pipeline: batch k+1 is in flight while batch k is processed (a two-deep
        // variant that also ran the id walk one batch further ahead measured 15 % slower on the
        // 70-bin workload)
        auto fetch = [&]() {
            int filled = 0;
            int32_t id = -1;
            while (filled < kBatch && pos >= 0)
            {
                const int take = min(end - pos, kBatch - filled);
                const int s = (int) lane - filled;
                if (s >= 0 && s < take) id = P.commands[pos + s];
                filled += take;
                pos += take;
                if (pos >= end)
                {
                    pos = P.commands[end];
                    end = pos >= 0 ? P.commands[pos++] : INT_MIN;
                }
            }
            nextCount = filled;
            rHave = id >= 0;
            if (rHave)
            {
                const float4 *sp = reinterpret_cast<const float4 *>(P.splats + id);
                rPos = __ldg(sp);
                rNq = __ldg(sp + 1);
                if (PRE) rThr = __ldg(P.thresholds + id);
            }
        };

        fetch();
        while (nextCount > 0)
        {
            if (COUNT) tests += (unsigned long long) kSlots * nextCount;
            // ---- cull: can the splat's ball reach this warp's brick at all? ----
            bool keep = false;
            if (rHave)
            {
                const float dx = fmaxf(fmaxf(bx0 - rPos.x, rPos.x - bx1), 0.0f);
                const float dy = fmaxf(fmaxf(by0 - rPos.y, rPos.y - by1), 0.0f);
                const float dz = fmaxf(fmaxf(bz0 - rPos.z, rPos.z - bz1), 0.0f);
                // a hit needs RN(|p|^2 * invr2) < 0.99 and |p|^2 >= dist^2 to the box for every corner
                // of the brick; the 1e-4 of slack dwarfs the few ulp of rounding on either side
                keep = (dx * dx + dy * dy + dz * dz) * rPos.w < 0.9901f;
            }
            const unsigned m = __ballot_sync(0xFFFFFFFFu, keep);
            const int n = __popc(m);
            if (keep)
            {
                const int p = __popc(m & ((1u << lane) - 1));
                wPosT[p] = make_float4(rPos.x, rPos.y, rPos.z, PRE ? rThr : hitThreshold(rPos.w));
                wNormQ[p] = rNq;
                wInvR[p] = rPos.w;
            }
            fetch();   // next batch: its global loads stay in flight while this one is processed
            __syncwarp();

            // ---- pair tests over the kept splats, in chain order ----
#if MLSB_MLS_VARIANT & 1
            for (int i = 0; i < n; i++)
            {
                const float4 s = wPosT[i];
#else
            float4 sNext = wPosT[0];
            for (int i = 0; i < n; i++)
            {
                const float4 s = sNext;
                if (i + 1 < n) sNext = wPosT[i + 1];
#endif
                const float pz = __fsub_rn(s.z, cz);
                const float py = __fsub_rn(s.y, cy);
                const float t = __fmaf_rn(py, py, __fmul_rn(pz, pz));
                float px[kSlots], pp[kSlots];
                bool hit[kSlots];
#pragma unroll
                for (int j = 0; j < kSlots; j++)
                {
                    px[j] = __fsub_rn(s.x, cx[j]);
                    pp[j] = __fmaf_rn(px[j], px[j], t);
                    hit[j] = pp[j] < s.w;   // <=> pp * invr2 < 0.99f, see hitThreshold()
                }
#if MLSB_MLS_VARIANT & 2
                // One uniform branch per splat; inside it the accumulation runs for every lane with
                // the weight (and pp, which alone could be non-finite) forced to zero where there is
                // no hit: adding w = 0 terms is exact, so no divergent region is needed.
                bool any = false;
#pragma unroll
                for (int j = 0; j < kSlots; j++) any |= hit[j];
                if (__any_sync(0xFFFFFFFFu, any))
                {
                    const float4 nq = wNormQ[i];
                    const float invr2 = wInvR[i];
                    const float u = __fmaf_rn(nq.y, py, __fmul_rn(nq.z, pz));   // shared part of dot3(n, p)
#pragma unroll
                    for (int j = 0; j < kSlots; j++)
                    {
                        const float d = __fmul_rn(pp[j], invr2);
                        float w = __fsub_rn(1.0f, d);
                        w = __fmul_rn(w, w);
                        w = __fmul_rn(w, w);
                        w = __fmul_rn(w, nq.w);
                        w = hit[j] ? w : 0.0f;
                        // (pp needs no such guard: w = 0 times a finite pp adds nothing, and pp is
                        // finite for every position a 2^20-cell grid can hold)
                        const float ppj = pp[j];
                        Accum &a = acc[j];
                        a.sumW = __fadd_rn(a.sumW, w);
                        a.sumWpx = __fmaf_rn(w, px[j], a.sumWpx);
                        a.sumWpy = __fmaf_rn(w, py, a.sumWpy);
                        a.sumWpz = __fmaf_rn(w, pz, a.sumWpz);
                        a.sumWnx = __fmaf_rn(w, nq.x, a.sumWnx);
                        a.sumWny = __fmaf_rn(w, nq.y, a.sumWny);
                        a.sumWnz = __fmaf_rn(w, nq.z, a.sumWnz);
                        a.sumWpp = __fmaf_rn(w, ppj, a.sumWpp);
                        a.sumWpn = __fmaf_rn(w, __fmaf_rn(nq.x, px[j], u), a.sumWpn);
                        // one predicated add (left to itself the compiler builds hits + 1, a
                        // predicated copy back and a move for one of the two slots)
                        asm("{\n\t.reg .pred p;\n\tsetp.lt.f32 p, %1, %2;\n\t@p add.u32 %0, %0, 1;\n\t}"
                            : "+r"(a.hits) : "f"(pp[j]), "f"(s.w));
                    }
                }
#else
                // after the cull nearly every kept splat hits somewhere in the brick, so there is
                // no brick-level vote: only one per slot (a 2x4x4 sub-brick)
                {
                    const float4 nq = wNormQ[i];
                    const float invr2 = wInvR[i];
                    const float u = __fmaf_rn(nq.y, py, __fmul_rn(nq.z, pz));   // shared part of dot3(n, p)
#pragma unroll
                    for (int j = 0; j < kSlots; j++)
                    {
                        if (__any_sync(0xFFFFFFFFu, hit[j]))
                        {
                            if (hit[j])
                            {
                                const float d = __fmul_rn(pp[j], invr2);
                                float w = __fsub_rn(1.0f, d);
                                w = __fmul_rn(w, w);
                                w = __fmul_rn(w, w);
                                w = __fmul_rn(w, nq.w);
                                Accum &a = acc[j];
                                a.sumW = __fadd_rn(a.sumW, w);
                                a.sumWpx = __fmaf_rn(w, px[j], a.sumWpx);
                                a.sumWpy = __fmaf_rn(w, py, a.sumWpy);
                                a.sumWpz = __fmaf_rn(w, pz, a.sumWpz);
                                a.sumWnx = __fmaf_rn(w, nq.x, a.sumWnx);
                                a.sumWny = __fmaf_rn(w, nq.y, a.sumWny);
                                a.sumWnz = __fmaf_rn(w, nq.z, a.sumWnz);
                                a.sumWpp = __fmaf_rn(w, pp[j], a.sumWpp);
                                a.sumWpn = __fmaf_rn(w, __fmaf_rn(nq.x, px[j], u), a.sumWpn);
                                a.hits++;
                            }
                        }
                    }
                }
#endif
            }
            __syncwarp();   // the buffer is overwritten by the next batch
        }

        unsigned long long hits = 0;
#pragma unroll
        for (int j = 0; j < kSlots; j++)
        {
            out[j] = finish<SHAPE>(acc[j], P.boundaryFactor);
            if (COUNT) hits += acc[j].hits;
        }
        if (COUNT)
        {
            for (int o = 16; o > 0; o >>= 1)
            {
                tests += __shfl_xor_sync(0xFFFFFFFFu, tests, o);
                hits += __shfl_xor_sync(0xFFFFFFFFu, hits, o);
            }
            if (lane == 0)
            {
                atomicAdd(&P.work[0], tests);
                atomicAdd(&P.work[1], hits);
            }
        }
    }

    // write_imagef(corners, (x, y + z*zStride + zBias)), kernels/mls.cl:429-432
    const long row = (long) (wy + ly) + (long) (wz + lz) * P.zStride + P.zBias;
    float *dst = P.field + row * P.pitch + wx + lx;
#pragma unroll
    for (int j = 0; j < kSlots; j++) dst[2 * j] = out[j];
}

// The same kernel with a TWO-LEVEL cull (MLSB_MLS_CULL=block; measured, not the default — see mlsCorners).
// Above, each of the eight warps walks the block's whole
// command chain on its own: it fetches every splat (36 bytes), tests it against its brick and packs
// the survivors — a quarter of the kernel's instructions and 40 % of its stall samples (ncu, C2) go
// into doing that eight times over.  Here the BLOCK walks the chain once, 256 entries at a time:
// thread t fetches entry t, tests it against all eight 4x4x4 bricks at once (the per-axis distances to
// the two half-intervals, combined per brick with exactly the arithmetic of the per-warp test, so the
// same splats reach the same bricks) and the entries that reach at least one brick are compacted, in
// chain order, into shared memory together with their 8-bit brick mask.  A warp then only lists the
// staged entries whose mask has its bit (a ballot per 32 entries) and pair-tests those straight out of
// the block's staging buffer.  Staging is double-buffered: two barriers per 256 chain entries, and the
// next batch's global loads are in flight while the current one is processed.  Per-corner
// accumulation order is still the chain order: the field is bit-identical.
namespace blockcull
{
constexpr int kThreads = kFullThreads;   // this kernel always takes a whole 8x8x8 block
constexpr int kStage = kThreads;   // chain entries per block batch
template <int SHAPE, bool COUNT, bool PRE>
__global__ void __launch_bounds__(kThreads, MLSB_MLS_MINBLOCKS) processCornersBlockCullKernel(MlsParams P)
{
    static_assert(kSlots == 2 && kThreads == 256, "the brick layout below is the 8 x (4x4x4) one");
    __shared__ float4 sPosT[2][kStage];      // x, y, z, hit threshold on |p|^2
    __shared__ float4 sNormQ[2][kStage];     // normal, quality
    __shared__ float sInvR[2][kStage];       // 1/r^2
    __shared__ uint8_t sMask[2][kStage];     // bricks the splat's ball can reach
    __shared__ uint16_t sIdx[kThreads / 32][kStage];   // per warp: staged entries of its brick, in order
    __shared__ uint32_t sWarpCount[2][kThreads / 32];

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31, quad = tid >> 5;
    const int qx = 4 * (int) (quad & 1), qy = 4 * (int) ((quad >> 1) & 1), qz = 4 * (int) (quad >> 2);
    const int lx = qx + (int) (lane & 1);
    const int ly = qy + (int) ((lane >> 1) & 3);
    const int lz = qz + (int) ((lane >> 3) & 3);
    const int wx = blockIdx.x * kMlsWgs, wy = blockIdx.y * kMlsWgs, wz = blockIdx.z * kMlsWgs + (int) P.zFirst;

    const uint32_t code = (spread3(wx) | (spread3(wy) << 1) | (spread3(wz) << 2)) >> P.startShift;
    int32_t pos = P.start[code];
    if (P.blockActive != nullptr && tid == 0)
        P.blockActive[((size_t) (wz >> 3) * P.activeDimY + blockIdx.y) * P.activeDimX + blockIdx.x] = pos >= 0;

    float out[kSlots];
#pragma unroll
    for (int j = 0; j < kSlots; j++) out[j] = __uint_as_float(0x7FC00000u);

    if (pos >= 0)   // uniform over the block: the barriers below are safe
    {
        const float cy = (float) (wy + ly + P.offset.y);
        const float cz = (float) (wz + lz + P.offset.z);
        float cx[kSlots];
#pragma unroll
        for (int j = 0; j < kSlots; j++) cx[j] = (float) (wx + lx + 2 * j + P.offset.x);

        Accum acc[kSlots];
#pragma unroll
        for (int j = 0; j < kSlots; j++)
        {
            acc[j].sumW = acc[j].sumWpx = acc[j].sumWpy = acc[j].sumWpz = 0.0f;
            acc[j].sumWnx = acc[j].sumWny = acc[j].sumWnz = acc[j].sumWpp = acc[j].sumWpn = 0.0f;
            acc[j].hits = 0;
        }
        unsigned long long tests = 0;

        // the block's low corner; brick (hx, hy, hz) spans [4h, 4h + 3] along each axis from there
        const float ox = (float) (wx + P.offset.x), oy = (float) (wy + P.offset.y), oz = (float) (wz + P.offset.z);

        int32_t end = P.commands[pos++];
        float4 rPos = make_float4(0.f, 0.f, 0.f, 0.f), rNq = rPos;
        float rThr = 0.0f;
        bool rHave = false;
        int nextCount = 0;
        // the next kStage entries of the chain, entry t to thread t; a batch may span several ranges
        auto fetch = [&]() {
            int filled = 0;
            int32_t id = -1;
            while (filled < kStage && pos >= 0)
            {
                const int take = min(end - pos, kStage - filled);
                const int s = (int) tid - filled;
                if (s >= 0 && s < take) id = P.commands[pos + s];
                filled += take;
                pos += take;
                if (pos >= end)
                {
                    pos = P.commands[end];
                    end = pos >= 0 ? P.commands[pos++] : INT_MIN;
                }
            }
            nextCount = filled;
            rHave = id >= 0;
            if (rHave)
            {
                const float4 *sp = reinterpret_cast<const float4 *>(P.splats + id);
                rPos = __ldg(sp);
                rNq = __ldg(sp + 1);
                if (PRE) rThr = __ldg(P.thresholds + id);
            }
        };

        fetch();
        for (int buf = 0; nextCount > 0; buf ^= 1)
        {
            if (COUNT) tests += (unsigned long long) kSlots * nextCount;   // per warp: its 64 corners against every entry
            // ---- block-level cull: which bricks can the splat's ball reach? ----
            uint32_t mask = 0;
            if (rHave)
            {
                float sx[2], sy[2], sz[2];
#pragma unroll
                for (int h = 0; h < 2; h++)
                {
                    const float x0 = ox + 4.0f * h, y0 = oy + 4.0f * h, z0 = oz + 4.0f * h;
                    const float dx = fmaxf(fmaxf(x0 - rPos.x, rPos.x - (x0 + 3.0f)), 0.0f);
                    const float dy = fmaxf(fmaxf(y0 - rPos.y, rPos.y - (y0 + 3.0f)), 0.0f);
                    const float dz = fmaxf(fmaxf(z0 - rPos.z, rPos.z - (z0 + 3.0f)), 0.0f);
                    sx[h] = dx * dx;
                    sy[h] = dy * dy;
                    sz[h] = dz * dz;
                }
                // a hit needs RN(|p|^2 * invr2) < 0.99 and |p|^2 >= dist^2 to the box for every corner
                // of the brick; the 1e-4 of slack dwarfs the few ulp of rounding on either side
#pragma unroll
                for (int q = 0; q < 8; q++)
                    if ((sx[q & 1] + sy[(q >> 1) & 1] + sz[q >> 2]) * rPos.w < 0.9901f) mask |= 1u << q;
            }
            const unsigned ball = __ballot_sync(0xFFFFFFFFu, mask != 0);
            if (lane == 0) sWarpCount[buf][quad] = __popc(ball);
            __syncthreads();
            uint32_t before = 0, n = 0;
#pragma unroll
            for (int w = 0; w < kThreads / 32; w++)
            {
                const uint32_t c = sWarpCount[buf][w];
                before += (uint32_t) w < quad ? c : 0u;
                n += c;
            }
            if (mask)
            {
                const uint32_t p = before + __popc(ball & ((1u << lane) - 1));
                sPosT[buf][p] = make_float4(rPos.x, rPos.y, rPos.z, PRE ? rThr : hitThreshold(rPos.w));
                sNormQ[buf][p] = rNq;
                sInvR[buf][p] = rPos.w;
                sMask[buf][p] = (uint8_t) mask;
            }
            __syncthreads();
            fetch();   // next batch: its global loads stay in flight while this one is processed

            // ---- this warp's entries, in chain order ----
            uint16_t *const wIdx = sIdx[quad];
            uint32_t m = 0;
            for (uint32_t e0 = 0; e0 < n; e0 += 32)
            {
                const uint32_t e = e0 + lane;
                const bool mine = e < n && ((sMask[buf][e] >> quad) & 1);
                const unsigned b = __ballot_sync(0xFFFFFFFFu, mine);
                if (mine) wIdx[m + __popc(b & ((1u << lane) - 1))] = (uint16_t) e;
                m += __popc(b);
            }
            __syncwarp();

            // ---- pair tests over them ----
            const float4 *const bPosT = sPosT[buf];
            const float4 *const bNormQ = sNormQ[buf];
            const float *const bInvR = sInvR[buf];
            for (uint32_t i = 0; i < m; i++)
            {
                const uint32_t e = wIdx[i];
                const float4 s = bPosT[e];
                const float pz = __fsub_rn(s.z, cz);
                const float py = __fsub_rn(s.y, cy);
                const float t = __fmaf_rn(py, py, __fmul_rn(pz, pz));
                float px[kSlots], pp[kSlots];
                bool hit[kSlots];
#pragma unroll
                for (int j = 0; j < kSlots; j++)
                {
                    px[j] = __fsub_rn(s.x, cx[j]);
                    pp[j] = __fmaf_rn(px[j], px[j], t);
                    hit[j] = pp[j] < s.w;   // <=> pp * invr2 < 0.99f, see hitThreshold()
                }
                bool any = false;
#pragma unroll
                for (int j = 0; j < kSlots; j++) any |= hit[j];
                if (__any_sync(0xFFFFFFFFu, any))
                {
                    const float4 nq = bNormQ[e];
                    const float invr2 = bInvR[e];
                    const float u = __fmaf_rn(nq.y, py, __fmul_rn(nq.z, pz));   // shared part of dot3(n, p)
#pragma unroll
                    for (int j = 0; j < kSlots; j++)
                    {
                        const float d = __fmul_rn(pp[j], invr2);
                        float w = __fsub_rn(1.0f, d);
                        w = __fmul_rn(w, w);
                        w = __fmul_rn(w, w);
                        w = __fmul_rn(w, nq.w);
                        w = hit[j] ? w : 0.0f;   // adding w = 0 terms is exact: no divergent region
                        const float ppj = pp[j];
                        Accum &a = acc[j];
                        a.sumW = __fadd_rn(a.sumW, w);
                        a.sumWpx = __fmaf_rn(w, px[j], a.sumWpx);
                        a.sumWpy = __fmaf_rn(w, py, a.sumWpy);
                        a.sumWpz = __fmaf_rn(w, pz, a.sumWpz);
                        a.sumWnx = __fmaf_rn(w, nq.x, a.sumWnx);
                        a.sumWny = __fmaf_rn(w, nq.y, a.sumWny);
                        a.sumWnz = __fmaf_rn(w, nq.z, a.sumWnz);
                        a.sumWpp = __fmaf_rn(w, ppj, a.sumWpp);
                        a.sumWpn = __fmaf_rn(w, __fmaf_rn(nq.x, px[j], u), a.sumWpn);
                        asm("{\n\t.reg .pred p;\n\tsetp.lt.f32 p, %1, %2;\n\t@p add.u32 %0, %0, 1;\n\t}"
                            : "+r"(a.hits) : "f"(pp[j]), "f"(s.w));
                    }
                }
            }
            // no barrier here: the next batch goes to the other buffer, and nobody can be two batches
            // ahead of a warp that is still reading this one (the two barriers of the next batch are in between)
        }

        unsigned long long hits = 0;
#pragma unroll
        for (int j = 0; j < kSlots; j++)
        {
            out[j] = finish<SHAPE>(acc[j], P.boundaryFactor);
            if (COUNT) hits += acc[j].hits;
        }
        if (COUNT)
        {
            for (int o = 16; o > 0; o >>= 1)
            {
                tests += __shfl_xor_sync(0xFFFFFFFFu, tests, o);
                hits += __shfl_xor_sync(0xFFFFFFFFu, hits, o);
            }
            if (lane == 0)
            {
                atomicAdd(&P.work[0], tests);
                atomicAdd(&P.work[1], hits);
            }
        }
    }

    // write_imagef(corners, (x, y + z*zStride + zBias)), kernels/mls.cl:429-432
    const long row = (long) (wy + ly) + (long) (wz + lz) * P.zStride + P.zBias;
    float *dst = P.field + row * P.pitch + wx + lx;
#pragma unroll
    for (int j = 0; j < kSlots; j++) dst[2 * j] = out[j];
}

}  // namespace blockcull

// Dependent-free FFMA stream: 8 independent accumulators per thread, every issue slot an FFMA.
__global__ void __launch_bounds__(256) fp32PeakKernel(float *out, float a, float b, int iters)
{
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = a + (float) (threadIdx.x + i);
    for (int it = 0; it < iters; it++)
    {
#pragma unroll
        for (int r = 0; r < 16; r++)
        {
#pragma unroll
            for (int i = 0; i < 8; i++) x[i] = __fmaf_rn(x[i], a, b);
        }
    }
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
    if (s == 12345.678f) out[0] = s;   // never true; keeps the loop alive
}

}  // namespace

// FP32 pipe peak of the current device in lane-instructions per second (one FFMA = one
// lane-instruction = 2 flop): the roofline denominator for processCorners, which
// MEASURED_PEAKS.json does not contain.
int measureFp32Peak(double *laneInstPerSecond, double *seconds)
{
    float *out = nullptr;
    MLSB_CUDA(cudaMalloc(&out, sizeof(float)));
    cudaEvent_t e0, e1;
    MLSB_CUDA(cudaEventCreate(&e0));
    MLSB_CUDA(cudaEventCreate(&e1));
    int sms = 0;
    MLSB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int blocks = sms * 8, iters = 4096;
    double best = 0.0, bestT = 0.0;
    for (int rep = 0; rep < 6; rep++)
    {
        MLSB_CUDA(cudaEventRecord(e0, 0));
        fp32PeakKernel<<<blocks, 256>>>(out, 0.999f, 0.001f, iters);
        MLSB_CUDA(cudaEventRecord(e1, 0));
        MLSB_CUDA(cudaEventSynchronize(e1));
        float ms = 0.0f;
        MLSB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        double inst = (double) blocks * 256 * iters * 16 * 8;
        double rate = inst / (ms * 1e-3);
        if (rep > 0 && rate > best) { best = rate; bestT = ms * 1e-3; }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *laneInstPerSecond = best;
    if (seconds) *seconds = bestT;
    return MLSB_OK;
}

int mlsAlloc(mlsb_worker *w)
{
    MLSB_CUDA(cudaMalloc(&w->mlsWork, 2 * sizeof(unsigned long long)));
    MLSB_CUDA(cudaMemset(w->mlsWork, 0, 2 * sizeof(unsigned long long)));
    return MLSB_OK;
}

void mlsFree(mlsb_worker *w) { cudaFree(w->mlsWork); }

int mlsCorners(mlsb_worker *w, float *field, uint32_t pitch, const mlsb_swathe &sw, bool recordActive)
{
    MLSB_REQUIRE(w->mlsBound, MLSB_ERR_STATE, "mlsb_mls_set has not been called");
    // src/mls.cpp:108-117
    uint32_t width = roundUp(sw.width, (uint32_t) kMlsWgs);
    uint32_t height = roundUp(sw.height, (uint32_t) kMlsWgs);
    MLSB_REQUIRE(sw.zStride >= height, MLSB_ERR_INVALID_ARGUMENT, "zStride smaller than the padded height");
    MLSB_REQUIRE(sw.zFirst <= sw.zLast, MLSB_ERR_INVALID_ARGUMENT, "empty swathe");
    MLSB_REQUIRE(sw.zFirst % kMlsWgs == 0, MLSB_ERR_INVALID_ARGUMENT, "zFirst must be a multiple of 8");
    MLSB_REQUIRE(pitch >= width, MLSB_ERR_LENGTH, "field pitch smaller than the padded width");
    MLSB_REQUIRE((long) sw.zFirst * sw.zStride + sw.zBias >= 0, MLSB_ERR_INVALID_ARGUMENT, "swathe starts before the image");

    MlsParams P;
    P.splats = w->mlsSplats;
    P.commands = w->mlsCommands;
    P.start = w->mlsStart;
    P.thresholds = w->mlsThresholds;
    P.field = field;
    P.pitch = pitch;
    P.startShift = w->mlsStartShift;
    P.offset = make_int3(w->mlsOffset[0], w->mlsOffset[1], w->mlsOffset[2]);
    P.zStride = sw.zStride;
    P.zBias = sw.zBias;
    P.zFirst = sw.zFirst;
    P.boundaryFactor = w->boundaryFactor;
    P.work = w->mlsWork;
    P.blockActive = recordActive ? w->blockActive : nullptr;
    P.activeDimX = P.activeDimY = roundUp(w->fieldDim / kMlsWgs, 4u);   // rows are read as 32-bit words
    const uint32_t blocksZ = divUp(sw.zLast - sw.zFirst + 1, (uint32_t) kMlsWgs);
    dim3 grid(width / kMlsWgs, height / kCtaY, blocksZ * (kMlsWgs / kCtaZ));
    dim3 fullGrid(width / kMlsWgs, height / kMlsWgs, blocksZ);
    ScopedStat st(w, "kernel.mls.processCorners.time", 1, 1);
    const bool count = w->countWork;
    cudaStream_t s = w->stream;
    if (w->mlsStream)
    {
        MLSB_CUDA(cudaEventRecord(w->mlsBefore, w->stream));
        MLSB_CUDA(cudaStreamWaitEvent(w->mlsStream, w->mlsBefore, 0));
        s = w->mlsStream;
    }
    const bool pre = P.thresholds != nullptr;
    // MLSB_MLS_CULL=block selects the two-level cull.  It is bit-identical and executes fewer instructions, but
    // measured on C2 it is no faster (lone worker 20.4 against 19.4 ms per step, bench 356 against 358 Msplats/s):
    // its two barriers per 256 chain entries tie the eight warps of a block to the slowest brick, and the
    // independent per-warp walks hide their own fetch latency better.  The per-warp kernel stays the default.
    static const bool blockCull = [] { const char *e = getenv("MLSB_MLS_CULL"); return e != nullptr && strcmp(e, "block") == 0; }();
#define MLSB_LAUNCH(SHAPE)                                                                                  \
    do {                                                                                                    \
        if (blockCull)                                                                                      \
        {                                                                                                   \
            if (count && pre) blockcull::processCornersBlockCullKernel<SHAPE, true, true><<<fullGrid, kFullThreads, 0, s>>>(P);    \
            else if (count) blockcull::processCornersBlockCullKernel<SHAPE, true, false><<<fullGrid, kFullThreads, 0, s>>>(P);     \
            else if (pre) blockcull::processCornersBlockCullKernel<SHAPE, false, true><<<fullGrid, kFullThreads, 0, s>>>(P);       \
            else blockcull::processCornersBlockCullKernel<SHAPE, false, false><<<fullGrid, kFullThreads, 0, s>>>(P);               \
        }                                                                                                   \
        else if (count && pre) processCornersKernel<SHAPE, true, true><<<grid, kThreads, 0, s>>>(P);        \
        else if (count) processCornersKernel<SHAPE, true, false><<<grid, kThreads, 0, s>>>(P);              \
        else if (pre) processCornersKernel<SHAPE, false, true><<<grid, kThreads, 0, s>>>(P);                \
        else processCornersKernel<SHAPE, false, false><<<grid, kThreads, 0, s>>>(P);                        \
    } while (0)
    if (w->cfg.shape == MLSB_SHAPE_SPHERE) MLSB_LAUNCH(MLSB_SHAPE_SPHERE);
    else MLSB_LAUNCH(MLSB_SHAPE_PLANE);
#undef MLSB_LAUNCH
    w->launches++;
    MLSB_CUDA(cudaGetLastError());
    if (w->mlsStream)
    {
        MLSB_CUDA(cudaEventRecord(w->mlsAfter, w->mlsStream));
        MLSB_CUDA(cudaStreamWaitEvent(w->stream, w->mlsAfter, 0));
    }
    return MLSB_OK;
}

}  // namespace mlsb
