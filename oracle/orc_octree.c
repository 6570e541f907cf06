/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see orc_common.h).
 *
 * Restatement of kernels/octree.cl and the driver order in
 * src/splat_tree_cl.cpp:269-335 (SplatTreeCL::enqueueBuild).
 */
#include "orc_common.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>

/* ---- clogs semantics ------------------------------------------------------------------ */

void orc_radixsort_u32_u32(uint32_t *keys, uint32_t *values, size_t n, unsigned bits)
{
    if (n == 0) return;
    if (bits == 0 || bits > 32) bits = 32;
    uint32_t *k2 = (uint32_t *) malloc(n * sizeof(uint32_t));
    uint32_t *v2 = (uint32_t *) malloc(n * sizeof(uint32_t));
    uint32_t *ks = keys, *vs = values, *kd = k2, *vd = v2;
    for (unsigned shift = 0; shift < bits; shift += 8)
    {
        unsigned w = bits - shift < 8 ? bits - shift : 8;
        uint32_t mask = (1u << w) - 1;
        size_t hist[257];
        memset(hist, 0, sizeof(hist));
        for (size_t i = 0; i < n; i++)
            hist[((ks[i] >> shift) & mask) + 1]++;
        for (unsigned d = 0; d < 256; d++)
            hist[d + 1] += hist[d];
        for (size_t i = 0; i < n; i++)
        {
            size_t p = hist[(ks[i] >> shift) & mask]++;
            kd[p] = ks[i];
            vd[p] = vs[i];
        }
        uint32_t *t;
        t = ks; ks = kd; kd = t;
        t = vs; vs = vd; vd = t;
    }
    if (ks != keys)
    {
        memcpy(keys, ks, n * sizeof(uint32_t));
        memcpy(values, vs, n * sizeof(uint32_t));
    }
    free(k2);
    free(v2);
}

void orc_radixsort_u64_f4(uint64_t *keys, float *values4, size_t n, unsigned bits)
{
    if (n == 0) return;
    if (bits == 0 || bits > 64) bits = 64;
    uint64_t *k2 = (uint64_t *) malloc(n * sizeof(uint64_t));
    float *v2 = (float *) malloc(n * 4 * sizeof(float));
    uint64_t *ks = keys, *kd = k2;
    float *vs = values4, *vd = v2;
    for (unsigned shift = 0; shift < bits; shift += 8)
    {
        unsigned w = bits - shift < 8 ? bits - shift : 8;
        uint64_t mask = (1ull << w) - 1;
        size_t hist[257];
        memset(hist, 0, sizeof(hist));
        for (size_t i = 0; i < n; i++)
            hist[((ks[i] >> shift) & mask) + 1]++;
        for (unsigned d = 0; d < 256; d++)
            hist[d + 1] += hist[d];
        for (size_t i = 0; i < n; i++)
        {
            size_t p = hist[(ks[i] >> shift) & mask]++;
            kd[p] = ks[i];
            memcpy(vd + 4 * p, vs + 4 * i, 4 * sizeof(float));
        }
        uint64_t *t = ks; ks = kd; kd = t;
        float *u = vs; vs = vd; vd = u;
    }
    if (ks != keys)
    {
        memcpy(keys, ks, n * sizeof(uint64_t));
        memcpy(values4, vs, n * 4 * sizeof(float));
    }
    free(k2);
    free(v2);
}

/* ---- kernels/octree.cl helpers -------------------------------------------------------- */

/* kernels/octree.cl:50-55 */
ORC_EXPORT int orc_level_shift(int lox, int loy, int loz, int hix, int hiy, int hiz)
{
    int dx = hix - lox, dy = hiy - loy, dz = hiz - loz;
    int big = dx > dy ? dx : dy;
    big = big > dz ? big : dz;
    return big > 1 ? 32 - __builtin_clz((unsigned) (big - 1)) : 0;
}

/* kernels/octree.cl:60-66 */
ORC_EXPORT float orc_point_box_dist2(const float pos[3], const float lo[3], const float hi[3])
{
    float dist[3];
    for (int i = 0; i < 3; i++)
    {
        float m = hi[i] < pos[i] ? hi[i] : pos[i];      /* min(hi, pos) */
        float nearest = lo[i] > m ? lo[i] : m;          /* max(lo, .)   */
        dist[i] = nearest - pos[i];
    }
    /* OpenCL dot(): taken as x*x + y*y + z*z evaluated left to right, unfused. */
    return dist[0] * dist[0] + dist[1] * dist[1] + dist[2] * dist[2];
}

/* kernels/octree.cl:121-135 (same function as kernels/mls.cl:150-164) */
ORC_EXPORT uint32_t orc_make_code(int x, int y, int z)
{
    uint32_t ans = 0;
    uint32_t scale = 1;
    y <<= 1;
    z <<= 2;
    while (x != 0 || y != 0 || z != 0)
    {
        uint32_t bits = (uint32_t) ((x & 1) | (y & 2) | (z & 4));
        ans += bits * scale;
        scale <<= 3;
        x >>= 1;
        y >>= 1;
        z >>= 1;
    }
    return ans;
}

/* kernels/octree.cl:79-89 */
static void prepare(int ilo[3], int *shift, int minShift, int maxShift,
                    const float pr[4], const int bias[3])
{
    int lo[3], hi[3];
    for (int i = 0; i < 3; i++)
    {
        lo[i] = (int) floorf(pr[i] - pr[3]);   /* convert_int3_rtn */
        hi[i] = (int) floorf(pr[i] + pr[3]);
    }
    int s = orc_level_shift(lo[0], lo[1], lo[2], hi[0], hi[1], hi[2]);
    s = s < minShift ? minShift : s;           /* clamp(x, lo, hi) = min(max(x, lo), hi) */
    s = s > maxShift ? maxShift : s;
    *shift = s;
    for (int i = 0; i < 3; i++)
    {
        int v = lo[i] - bias[i];
        v = v > 0 ? v : 0;
        ilo[i] = v >> s;
    }
}

/* kernels/octree.cl:100-110 */
static int good_entry(const int cell[3], int shift, const float position[3], float radius2,
                      const int bias[3])
{
    float vblo[3], vbhi[3];
    for (int i = 0; i < 3; i++)
    {
        int blo = (int) ((unsigned) cell[i] << shift) + bias[i];
        int bhi = (int) ((unsigned) (cell[i] + 1) << shift) + bias[i];
        vblo[i] = (float) blo;
        vbhi[i] = (float) bhi;
    }
    return orc_point_box_dist2(position, vblo, vbhi) < radius2;
}

/* kernels/octree.cl:159-214.  keys/values are splat-major: slot j of splat i at [8*i + j].
 * Mutates splats[first + i].radius to 1/r^2 (kernels/octree.cl:193). */
ORC_EXPORT void orc_write_entries(uint32_t *keys, uint32_t *values, OrcSplat *splats,
                                  const int bias[3], uint32_t minShift, uint32_t maxShift,
                                  uint32_t firstSplat, uint32_t numSplats)
{
    uint32_t levelOffsets[32];
    {
        uint32_t pos = 0;
        uint32_t add = 1u << (3 * (maxShift - minShift));
        for (uint32_t i = minShift; i <= maxShift; i++)
        {
            levelOffsets[i] = pos;
            pos += add;
            add >>= 3;
        }
    }
    for (uint32_t g = 0; g < numSplats; g++)
    {
        uint32_t pos = g * 8;
        uint32_t gid = g + firstSplat;
        float pr[4] = { splats[gid].position[0], splats[gid].position[1],
                        splats[gid].position[2], splats[gid].radius };
        int ilo[3], shift;
        prepare(ilo, &shift, (int) minShift, (int) maxShift, pr, bias);

        float radius2 = pr[3] * pr[3];
        splats[gid].radius = 1.0f / radius2;
        radius2 *= 1.00001f;
        uint32_t levelOffset = levelOffsets[shift];
        int bound = 1 << (maxShift - shift);
        for (int oz = 0; oz < 2; oz++)
            for (int oy = 0; oy < 2; oy++)
                for (int ox = 0; ox < 2; ox++)
                {
                    int addr[3] = { ilo[0] + ox, ilo[1] + oy, ilo[2] + oz };
                    uint32_t key = orc_make_code(addr[0], addr[1], addr[2]) + levelOffset;
                    int isect = good_entry(addr, shift, pr, radius2, bias);
                    isect &= (addr[0] < bound) && (addr[1] < bound) && (addr[2] < bound);
                    key = isect ? key : UINT32_MAX;
                    values[pos] = gid;
                    keys[pos] = key;
                    pos++;
                }
    }
}

/* kernels/octree.cl:230-239 (numKeys - 1 work items) */
static void count_commands(uint32_t *indicator, const uint32_t *keys, size_t numKeys)
{
    for (size_t pos = 0; pos + 1 < numKeys; pos++)
        indicator[pos] = keys[pos] != keys[pos + 1] ? 3 : 1;
}

/* clogs::Scan with offset: in-place exclusive prefix sum (src/splat_tree_cl.cpp:312-313) */
static void scan_u32(uint32_t *buf, size_t n, uint32_t offset)
{
    uint32_t sum = offset;
    for (size_t i = 0; i < n; i++)
    {
        uint32_t v = buf[i];
        buf[i] = sum;
        sum += v;
    }
}

/* kernels/octree.cl:256-279 */
static void write_splat_ids(int32_t *commands, int32_t *start, int32_t *jumpPos,
                            const uint32_t *commandMap, const uint32_t *keys,
                            const uint32_t *splatIds, size_t numEntries)
{
    for (size_t pos = 0; pos < numEntries; pos++)
    {
        uint32_t curKey = keys[pos];
        if (curKey != UINT32_MAX)
        {
            uint32_t cpos = commandMap[pos];
            commands[cpos] = (int32_t) splatIds[pos];
            uint32_t prevKey = pos > 0 ? keys[pos - 1] : UINT32_MAX;
            uint32_t nextKey = pos < numEntries - 1 ? keys[pos + 1] : UINT32_MAX;
            if (prevKey != curKey)
                start[curKey] = (int32_t) (cpos - 1);
            if (curKey != nextKey)
                jumpPos[curKey] = (int32_t) (cpos + 1);
        }
    }
}

/* kernels/octree.cl:292-341 (writeStart / writeStartTop) */
static void write_start(int32_t *start, int32_t *commands, const int32_t *jumpPos,
                        uint32_t curOffset, int havePrev, uint32_t prevOffset, size_t numCodes)
{
    for (size_t code = 0; code < numCodes; code++)
    {
        size_t pos = code + curOffset;
        int32_t jp = jumpPos[pos];
        int32_t prev = havePrev ? start[prevOffset + (code >> 3)] : -1;
        if (jp >= 0)
        {
            commands[jp] = prev;
            commands[start[pos]] = jp;
        }
        else
            start[pos] = prev;
    }
}

/* Size of the start array: src/splat_tree_cl.cpp:289-296 */
ORC_EXPORT uint32_t orc_octree_num_start(uint32_t maxLevels, uint32_t subsamplingShift)
{
    uint32_t maxShift = maxLevels + subsamplingShift - 1;
    uint32_t minShift = subsamplingShift < maxShift ? subsamplingShift : maxShift;
    uint32_t pos = 0;
    for (uint32_t i = minShift; i <= maxShift; i++)
        pos += 1u << (3 * (maxShift - i));
    return pos;
}

/*
 * SplatTreeCL::enqueueBuild, src/splat_tree_cl.cpp:269-335.
 *
 * start:    numStart entries (orc_octree_num_start)
 * commands: 8*numSplats + 2*min(numStart, 8*numSplats) entries (src/splat_tree_cl.cpp:120);
 *           pre-filled by the caller if it wants to detect unwritten slots
 * sortedKeys/sortedValues/commandMap: optional (may be NULL) 8*numSplats-entry outputs of
 *           the intermediate stages.
 * Returns the number of command slots in use (1 + last written index), 0 if none.
 */
ORC_EXPORT uint32_t orc_octree_build(OrcSplat *splats, uint32_t firstSplat, uint32_t numSplats,
                                     const int offset[3], uint32_t subsamplingShift,
                                     uint32_t maxLevels,
                                     int32_t *start, int32_t *commands,
                                     uint32_t *sortedKeys, uint32_t *sortedValues,
                                     uint32_t *commandMapOut)
{
    uint32_t maxShift = maxLevels + subsamplingShift - 1;
    uint32_t minShift = subsamplingShift < maxShift ? subsamplingShift : maxShift;
    uint32_t levelOffsets[32];
    uint32_t pos = 0;
    for (uint32_t i = minShift; i <= maxShift; i++)
    {
        levelOffsets[i] = pos;
        pos += 1u << (3 * (maxShift - i));
    }
    uint32_t numStart = pos;
    size_t numEntries = (size_t) numSplats * 8;

    uint32_t *keys = (uint32_t *) malloc((numEntries + 1) * sizeof(uint32_t));
    uint32_t *values = (uint32_t *) malloc((numEntries + 1) * sizeof(uint32_t));
    uint32_t *commandMap = (uint32_t *) malloc((numEntries + 1) * sizeof(uint32_t));
    int32_t *jumpPos = (int32_t *) malloc((size_t) numStart * sizeof(int32_t));

    orc_write_entries(keys, values, splats, offset, minShift, maxShift, firstSplat, numSplats);
    orc_radixsort_u32_u32(keys, values, numEntries, 3 * (maxShift - minShift) + 1);
    if (numEntries > 0)
        commandMap[numEntries - 1] = 0;  /* not written by countCommands; value unused by an exclusive scan */
    count_commands(commandMap, keys, numEntries);
    scan_u32(commandMap, numEntries, 1);
    for (uint32_t i = 0; i < numStart; i++)
        jumpPos[i] = -1;
    write_splat_ids(commands, start, jumpPos, commandMap, keys, values, numEntries);
    for (int i = (int) maxShift; i >= (int) minShift; i--)
    {
        size_t levelSize = (size_t) 1 << (3 * (maxShift - i));
        int havePrev = i != (int) maxShift;
        write_start(start, commands, jumpPos, levelOffsets[i], havePrev,
                    havePrev ? levelOffsets[i + 1] : 0, levelSize);
    }

    uint32_t used = 0;
    for (uint32_t i = 0; i < numStart; i++)
        if (jumpPos[i] >= 0 && (uint32_t) jumpPos[i] + 1 > used)
            used = (uint32_t) jumpPos[i] + 1;

    if (sortedKeys) memcpy(sortedKeys, keys, numEntries * sizeof(uint32_t));
    if (sortedValues) memcpy(sortedValues, values, numEntries * sizeof(uint32_t));
    if (commandMapOut) memcpy(commandMapOut, commandMap, numEntries * sizeof(uint32_t));
    free(keys);
    free(values);
    free(commandMap);
    free(jumpPos);
    return used;
}
