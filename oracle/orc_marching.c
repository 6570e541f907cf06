/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see orc_common.h).
 *
 * Restatement of kernels/marching.cl, the table construction in
 * src/marching.cpp:50-252 (Marching::makeTables), the weld sequence of
 * Marching::shipOut (src/marching.cpp:553-625) and kernels/scale_bias.cl.
 *
 * The reference compacts occupied cells with atomic_inc (kernels/marching.cl:113), so its
 * triangle order is nondeterministic; here cells are visited in x-fastest linear order.
 */
#include "orc_common.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define NUM_EDGES 19
#define NUM_CUBES 256
#define NUM_TETRAHEDRA 6
#define KEY_AXIS_BITS 21
#define KEY_EXTERNAL_FLAG (1ull << 63)

/* src/marching.cpp:50-71: the 19 cell edges as corner pairs (corner k sits at offset
 * (k&1, (k>>1)&1, k>>2)); every pair (a,b) with a<b whose offsets are componentwise <=. */
static const uint8_t edgeIndices[NUM_EDGES][2] = {
    {0, 1}, {0, 2}, {0, 3}, {1, 3}, {2, 3}, {0, 4}, {0, 5}, {1, 5}, {4, 5}, {0, 6},
    {2, 6}, {4, 6}, {0, 7}, {1, 7}, {2, 7}, {3, 7}, {4, 7}, {5, 7}, {6, 7}
};

/* src/marching.cpp:73-81: six tetrahedra around the 0-7 diagonal */
static const uint8_t tetrahedronIndices[NUM_TETRAHEDRA][4] = {
    {0, 7, 1, 3}, {0, 7, 3, 2}, {0, 7, 2, 6}, {0, 7, 6, 4}, {0, 7, 4, 5}, {0, 7, 5, 1}
};

typedef struct
{
    uint8_t count[NUM_CUBES][2];      /* countTable: (vertices, indices) */
    uint16_t start[NUM_CUBES + 1][2]; /* startTable: offsets into data; .y already biased */
    uint8_t data[8192];               /* dataTable: edge ids, then local indices */
    uint32_t key[2432][3];            /* keyTable */
    uint32_t numVertexEntries, numDataEntries;
    int built;
} Tables;

static Tables T;

static unsigned find_edge(unsigned v0, unsigned v1)
{
    if (v0 > v1) { unsigned t = v0; v0 = v1; v1 = t; }
    for (unsigned i = 0; i < NUM_EDGES; i++)
        if (edgeIndices[i][0] == v0 && edgeIndices[i][1] == v1)
            return i;
    abort();
}

static unsigned parity4(const uint8_t v[4])
{
    unsigned p = 0;
    for (int i = 0; i < 4; i++)
        for (int j = i + 1; j < 4; j++)
            if (v[i] > v[j]) p ^= 1;
    return p;
}

/* lexicographic successor of a 4-permutation (what std::next_permutation does) */
static int next_perm4(uint8_t v[4])
{
    int i = 2;
    while (i >= 0 && v[i] >= v[i + 1]) i--;
    if (i < 0) return 0;
    int j = 3;
    while (v[j] <= v[i]) j--;
    uint8_t t = v[i]; v[i] = v[j]; v[j] = t;
    for (int a = i + 1, b = 3; a < b; a++, b--) { t = v[a]; v[a] = v[b]; v[b] = t; }
    return 1;
}

/* src/marching.cpp:109-252 */
static void build_tables(void)
{
    if (T.built) return;
    static uint8_t vertexTable[8192], indexTable[8192];
    uint32_t nv = 0, ni = 0;
    for (unsigned code = 0; code < NUM_CUBES; code++)
    {
        uint32_t sv = nv, si = ni;
        uint8_t triangles[64];
        unsigned ntri = 0;
        for (unsigned j = 0; j < NUM_TETRAHEDRA; j++)
        {
            uint8_t vtx[4];
            int out[8] = {0};
            unsigned outside = 0;
            for (int k = 0; k < 4; k++)
            {
                vtx[k] = tetrahedronIndices[j][k];
                out[vtx[k]] = (code >> vtx[k]) & 1;
                outside += out[vtx[k]];
            }
            unsigned baseParity = parity4(vtx);
            if (outside > 2)
            {
                baseParity ^= 1;
                for (int k = 0; k < 4; k++) out[vtx[k]] = !out[vtx[k]];
            }
            /* sort ascending, then walk permutations in lexicographic order; take the
             * first rotation (same parity as the tetrahedron) in a canonical pattern */
            for (int a = 0; a < 4; a++)
                for (int b = a + 1; b < 4; b++)
                    if (vtx[a] > vtx[b]) { uint8_t t = vtx[a]; vtx[a] = vtx[b]; vtx[b] = t; }
            do
            {
                if (parity4(vtx) != baseParity) continue;
                unsigned mask = 0;
                for (int k = 0; k < 4; k++) mask |= (unsigned) out[vtx[k]] << k;
                unsigned t0 = vtx[0], t1 = vtx[1], t2 = vtx[2], t3 = vtx[3];
                if (mask == 0)
                    break;
                else if (mask == 1)
                {
                    triangles[ntri++] = find_edge(t0, t1);
                    triangles[ntri++] = find_edge(t0, t3);
                    triangles[ntri++] = find_edge(t0, t2);
                    break;
                }
                else if (mask == 3)
                {
                    triangles[ntri++] = find_edge(t0, t2);
                    triangles[ntri++] = find_edge(t1, t2);
                    triangles[ntri++] = find_edge(t1, t3);
                    triangles[ntri++] = find_edge(t1, t3);
                    triangles[ntri++] = find_edge(t0, t3);
                    triangles[ntri++] = find_edge(t0, t2);
                    break;
                }
            } while (next_perm4(vtx));
        }
        int edgeCompact[NUM_EDGES];
        int pool = 0;
        for (unsigned e = 0; e < NUM_EDGES; e++)
        {
            int used = 0;
            for (unsigned t = 0; t < ntri; t++) used |= triangles[t] == e;
            if (used)
            {
                edgeCompact[e] = pool++;
                for (int axis = 0; axis < 3; axis++)
                    T.key[nv][axis] = ((edgeIndices[e][0] >> axis) & 1) + ((edgeIndices[e][1] >> axis) & 1);
                vertexTable[nv++] = (uint8_t) e;
            }
        }
        for (unsigned t = 0; t < ntri; t++)
            indexTable[ni++] = (uint8_t) edgeCompact[triangles[t]];
        T.start[code][0] = (uint16_t) sv;
        T.start[code][1] = (uint16_t) si;
        T.count[code][0] = (uint8_t) (nv - sv);
        T.count[code][1] = (uint8_t) (ni - si);
    }
    T.start[NUM_CUBES][0] = (uint16_t) nv;
    T.start[NUM_CUBES][1] = (uint16_t) ni;
    for (unsigned i = 0; i <= NUM_CUBES; i++)
        T.start[i][1] = (uint16_t) (T.start[i][1] + nv);
    memcpy(T.data, vertexTable, nv);
    memcpy(T.data + nv, indexTable, ni);
    T.numVertexEntries = nv;
    T.numDataEntries = nv + ni;
    T.built = 1;
}

/* Copies out the four lookup tables.  count: 512 B; start: 257*2 u16; data: 8192 B;
 * key: 2432*3 u32.  sizes[0] = vertex entries, sizes[1] = total data entries
 * (asserted 2432 and 8192 at src/marching.cpp:248-251). */
ORC_EXPORT void orc_marching_tables(uint8_t *count, uint16_t *start, uint8_t *data, uint32_t *key,
                                    uint32_t sizes[2])
{
    build_tables();
    if (count) memcpy(count, T.count, sizeof(T.count));
    if (start) memcpy(start, T.start, sizeof(T.start));
    if (data) memcpy(data, T.data, sizeof(T.data));
    if (key) memcpy(key, T.key, sizeof(T.key));
    sizes[0] = T.numVertexEntries;
    sizes[1] = T.numDataEntries;
}

/* kernels/marching.cl:41-66 */
static unsigned cube_code(const float iso[8])
{
    unsigned c = 0;
    for (int i = 0; i < 8; i++)
        c |= (iso[i] >= 0.0f ? 1u : 0u) << i;
    return c;
}

static int cube_valid(const float iso[8])
{
    for (int i = 0; i < 8; i++)
        if (!orc_isfinite(iso[i])) return 0;
    return 1;
}

static void load_iso(float iso[8], const float *image, uint32_t pitch,
                     uint32_t x, uint32_t y, uint32_t z, uint32_t zStride, int32_t zBias)
{
    /* kernels/marching.cl:94-106 */
    long y0 = (long) y + (long) zStride * z + zBias;
    long y1 = y0 + zStride;
    iso[0] = image[y0 * pitch + x];
    iso[1] = image[y0 * pitch + x + 1];
    iso[2] = image[(y0 + 1) * pitch + x];
    iso[3] = image[(y0 + 1) * pitch + x + 1];
    iso[4] = image[y1 * pitch + x];
    iso[5] = image[y1 * pitch + x + 1];
    iso[6] = image[(y1 + 1) * pitch + x];
    iso[7] = image[(y1 + 1) * pitch + x + 1];
}

/* kernels/marching.cl:148-154 */
ORC_EXPORT uint64_t orc_compute_key(const uint32_t coords[3], const uint32_t top[3])
{
    uint64_t key = ((uint64_t) coords[2] << (2 * KEY_AXIS_BITS)) | ((uint64_t) coords[1] << KEY_AXIS_BITS) | (uint64_t) coords[0];
    if (coords[0] == 0 || coords[1] == 0
        || coords[0] == top[0] || coords[1] == top[1] || coords[2] == top[2])
        key |= KEY_EXTERNAL_FLAG;
    return key;
}

/* kernels/marching.cl:130-138, FP_CONTRACT OFF with an explicit fma */
static void interp(float out[3], float iso0, float iso1, const uint32_t cell[3], unsigned c0, unsigned c1)
{
    float inv = 1.0f / (iso0 - iso1);
    float t = iso0 * inv;
    for (int axis = 0; axis < 3; axis++)
    {
        uint32_t o0 = (c0 >> axis) & 1, o1 = (c1 >> axis) & 1;
        uint32_t delta = o1 - o0;
        out[axis] = __builtin_fmaf(t, (float) delta, (float) (cell[axis] + o0));
    }
}

/* kernels/marching.cl:295-326.  Exposed separately for the reference's known-answer test
 * (test/test_marching.cpp:401-479).  outKeys / outVertices slots that the kernel does not
 * write are left untouched. */
ORC_EXPORT void orc_compact_vertices(float *outVertices, uint64_t *outKeys, uint32_t *indexRemap,
                                     uint32_t *firstExternal, const uint32_t *vertexUnique,
                                     const float *inVertices4, const uint64_t *inKeys,
                                     uint64_t minExternalKey, uint64_t keyOffset, uint32_t n)
{
    for (uint32_t gid = 0; gid < n; gid++)
    {
        uint32_t u = vertexUnique[gid];
        const float *v = inVertices4 + 4 * (size_t) gid;
        uint64_t key = inKeys[gid];
        uint64_t nextKey = inKeys[gid + 1];
        int ext = key >= minExternalKey;
        if (key != nextKey)
        {
            outVertices[3 * (size_t) u + 0] = v[0];
            outVertices[3 * (size_t) u + 1] = v[1];
            outVertices[3 * (size_t) u + 2] = v[2];
            if (ext)
            {
                outKeys[u] = (key & (KEY_EXTERNAL_FLAG - 1)) + keyOffset;
                if (u == 0)
                    *firstExternal = 0;
            }
            else if (nextKey >= minExternalKey)
                *firstExternal = u + 1;
        }
        uint32_t originalIndex;
        memcpy(&originalIndex, &v[3], 4);
        indexRemap[originalIndex] = u;
    }
}

/* kernels/scale_bias.cl:32-40 */
ORC_EXPORT void orc_scale_bias(float *vertices, uint32_t numVertices, float scale, const float bias[3])
{
    for (uint32_t i = 0; i < numVertices; i++)
        for (int a = 0; a < 3; a++)
            vertices[3 * (size_t) i + a] = __builtin_fmaf(vertices[3 * (size_t) i + a], scale, bias[a]);
}

/*
 * One shipped-out block of Marching::generate covering the cells between slices zFirst and
 * zLast (corners zFirst..zLast): genOccupied (:84-120) -> scan -> generateElements (:184-258)
 * -> sentinel + sort by key -> countUniqueVertices (:271-279) -> scan -> compactVertices ->
 * reindex (:334-340), following Marching::addSlices / shipOut, src/marching.cpp:553-743,
 * with top.z = 2*zFirst (zTop) and minExternalKey = zLast << 43 (zMax).
 * With zFirst = 0, zLast = depth-1 this is the whole bin in a single shipOut.
 *
 * Vertices come out in grid-global coordinates (cell + keyOffset), i.e. before the
 * ScaleBiasFilter.  outKeys has one entry per welded vertex; entries of internal vertices
 * are left untouched (the reference leaves them undefined).
 * sizes = {numVertices, numTriangles, numInternalVertices}; optional stats =
 * {occupied cells, unwelded vertices, indices}.
 * Returns 0, or -1 if vcap/icap (capacities in vertices / indices) are too small.
 */
ORC_EXPORT int orc_marching_generate(const float *image, uint32_t pitch,
                                     uint32_t width, uint32_t height,
                                     uint32_t zStride, int32_t zBias,
                                     uint32_t zFirst, uint32_t zLast,
                                     const uint32_t keyOffset[3],
                                     float *outVertices, uint64_t *outKeys, uint32_t *outIndices,
                                     uint32_t vcap, uint32_t icap,
                                     uint32_t sizes[3], uint32_t *stats)
{
    build_tables();
    sizes[0] = sizes[1] = sizes[2] = 0;
    if (stats) stats[0] = stats[1] = stats[2] = 0;
    if (width < 2 || height < 2 || zLast <= zFirst)
        return 0;

    /* genOccupied */
    size_t maxCells = (size_t) (width - 1) * (height - 1) * (zLast - zFirst);
    uint32_t *cells = (uint32_t *) malloc(maxCells * 3 * sizeof(uint32_t) + 4);
    uint32_t *viCount = (uint32_t *) malloc(maxCells * 2 * sizeof(uint32_t) + 4);
    size_t compacted = 0;
    for (uint32_t z = zFirst; z < zLast; z++)
        for (uint32_t y = 0; y + 1 < height; y++)
            for (uint32_t x = 0; x + 1 < width; x++)
            {
                float iso[8];
                load_iso(iso, image, pitch, x, y, z, zStride, zBias);
                unsigned code = cube_code(iso);
                if (cube_valid(iso) && code != 0 && code != 255)
                {
                    cells[3 * compacted + 0] = x;
                    cells[3 * compacted + 1] = y;
                    cells[3 * compacted + 2] = z;
                    viCount[2 * compacted + 0] = T.count[code][0];
                    viCount[2 * compacted + 1] = T.count[code][1];
                    compacted++;
                }
            }

    /* scanElements (exclusive, offsets 0) */
    uint32_t nV = 0, nI = 0;
    for (size_t i = 0; i < compacted; i++)
    {
        uint32_t v = viCount[2 * i], n = viCount[2 * i + 1];
        viCount[2 * i] = nV;
        viCount[2 * i + 1] = nI;
        nV += v;
        nI += n;
    }
    if (stats) { stats[0] = (uint32_t) compacted; stats[1] = nV; stats[2] = nI; }
    if (nI > icap) { free(cells); free(viCount); return -1; }

    /* generateElements */
    float *verts = (float *) malloc(((size_t) nV + 1) * 4 * sizeof(float));
    uint64_t *keys = (uint64_t *) malloc(((size_t) nV + 1) * sizeof(uint64_t));
    const uint32_t top[3] = { 2 * (width - 1), 2 * (height - 1), 2 * zFirst };
    for (size_t g = 0; g < compacted; g++)
    {
        const uint32_t *cell = cells + 3 * g;
        uint32_t globalCell[3] = { cell[0] + keyOffset[0], cell[1] + keyOffset[1], cell[2] + keyOffset[2] };
        float iso[8];
        load_iso(iso, image, pitch, cell[0], cell[1], cell[2], zStride, zBias);
        float lverts[NUM_EDGES][3];
        for (unsigned e = 0; e < NUM_EDGES; e++)
            interp(lverts[e], iso[edgeIndices[e][0]], iso[edgeIndices[e][1]], globalCell,
                   edgeIndices[e][0], edgeIndices[e][1]);
        unsigned code = cube_code(iso);
        uint32_t vNext = viCount[2 * g], iNext = viCount[2 * g + 1];
        unsigned sx = T.start[code][0], sy = T.start[code][1];
        unsigned ex = T.start[code + 1][0], ey = T.start[code + 1][1];
        for (unsigned i = 0; i < ex - sx; i++)
        {
            float *v = verts + 4 * (size_t) (vNext + i);
            memcpy(v, lverts[T.data[sx + i]], 3 * sizeof(float));
            uint32_t idx = vNext + i;
            memcpy(&v[3], &idx, 4);
            uint32_t kc[3] = { 2 * cell[0] + T.key[sx + i][0], 2 * cell[1] + T.key[sx + i][1],
                               2 * cell[2] + T.key[sx + i][2] };
            keys[vNext + i] = orc_compute_key(kc, top);
        }
        for (unsigned i = 0; i < ey - sy; i++)
            outIndices[iNext + i] = vNext + T.data[sy + i];
    }

    /* shipOut */
    int rc = 0;
    if (nV > 0)
    {
        keys[nV] = UINT64_MAX;
        orc_radixsort_u64_f4(keys, verts, nV, 0);
        uint32_t *unique = (uint32_t *) malloc(((size_t) nV + 1) * sizeof(uint32_t));
        for (uint32_t i = 0; i < nV; i++)
            unique[i] = keys[i] != keys[i + 1] ? 1 : 0;
        unique[nV] = 0;
        uint32_t sum = 0;
        for (uint32_t i = 0; i <= nV; i++)
        {
            uint32_t v = unique[i];
            unique[i] = sum;
            sum += v;
        }
        uint32_t numWelded = unique[nV];
        if (numWelded > vcap)
            rc = -1;
        else
        {
            uint64_t minExternalKey = (uint64_t) zLast << (2 * KEY_AXIS_BITS + 1);
            uint64_t keyOffsetL = ((uint64_t) keyOffset[2] << (2 * KEY_AXIS_BITS + 1))
                | ((uint64_t) keyOffset[1] << (KEY_AXIS_BITS + 1))
                | ((uint64_t) keyOffset[0] << 1);
            uint32_t *indexRemap = (uint32_t *) malloc((size_t) nV * sizeof(uint32_t));
            uint32_t firstExternal = 0;
            orc_compact_vertices(outVertices, outKeys, indexRemap, &firstExternal, unique,
                                 verts, keys, minExternalKey, keyOffsetL, nV);
            for (uint32_t i = 0; i < nI; i++)
                outIndices[i] = indexRemap[outIndices[i]];
            sizes[0] = numWelded;
            sizes[1] = nI / 3;
            sizes[2] = firstExternal;
            free(indexRemap);
        }
        free(unique);
    }
    free(verts);
    free(keys);
    free(cells);
    free(viCount);
    return rc;
}
