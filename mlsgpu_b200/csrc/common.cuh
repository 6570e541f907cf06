// Internal definitions shared by the CUDA translation units of libmlsb200.
#pragma once

#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <map>
#include <string>
#include <vector>

#include "../../include/mlsb200.h"
#include "wire.h"

namespace mlsb
{

void setError(const char *fmt, ...);

#define MLSB_CUDA(call)                                                                       \
    do {                                                                                      \
        cudaError_t err_ = (call);                                                            \
        if (err_ != cudaSuccess) {                                                            \
            mlsb::setError("%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(err_)); \
            return MLSB_ERR_CUDA;                                                             \
        }                                                                                     \
    } while (0)

#define MLSB_REQUIRE(cond, status, msg)                                                       \
    do {                                                                                      \
        if (!(cond)) {                                                                        \
            mlsb::setError("%s:%d: %s (%s)", __FILE__, __LINE__, msg, #cond);                 \
            return status;                                                                    \
        }                                                                                     \
    } while (0)

#define MLSB_TRY(expr)                                                                        \
    do {                                                                                      \
        int st_ = (expr);                                                                     \
        if (st_ != MLSB_OK) return st_;                                                       \
    } while (0)

template <typename T> static inline T divUp(T a, T b) { return (a + b - 1) / b; }
template <typename T> static inline T roundUp(T a, T b) { return divUp(a, b) * b; }

constexpr int kMaxLevels = 10;             // SplatTreeCL::MAX_LEVELS, src/splat_tree_cl.h:76
constexpr int kMlsWgs = 8;                 // MlsFunctor::wgs, src/mls.cpp:53
constexpr unsigned kKeyAxisBits = 21;      // Marching::KEY_AXIS_BITS, src/marching.h:113
constexpr unsigned kMaxCellVertices = 13;  // src/marching.h:88
constexpr unsigned kMaxCellIndices = 36;   // src/marching.h:92
constexpr unsigned kMaxCellBytes = 872;    // src/marching.h:97-100

// The hit test of kernels/mls.cl:373-374 is  d = pp * invr2;  d < 0.99f.  x -> RN(x * w) is
// monotone for w > 0, so the set of pp >= 0 that pass is exactly { pp < T } for one float T per
// splat: the smallest T with RN(T * w) >= 0.99f.  Comparing against T saves the multiply in the
// pair test (d itself is only needed for a hit) without changing a single decision.
// w = +0 (infinite radius) passes every finite pp: T = +inf.  w = +inf or NaN passes nothing: T = 0.
// Computed once per splat by the octree build (a side array), or by the MLS kernel itself when it
// is bound to buffers that did not come from our build.
#ifdef __CUDACC__
__device__ __forceinline__ float hitThreshold(float w)
{
    const float cutoff = 0.99f;   // RADIUS_CUTOFF, kernels/mls.cl:36
    if (!(w > 0.0f)) return w == 0.0f ? __uint_as_float(0x7F800000u) : 0.0f;
    if (__float_as_uint(w) == 0x7F800000u) return 0.0f;
    float t = __fdiv_rn(cutoff, w);
    if ((__float_as_uint(t) & 0x7F800000u) == 0x7F800000u) t = __uint_as_float(0x7F7FFFFFu);
    for (int k = 0; k < 8 && t > 0.0f && __fmul_rn(t, w) >= cutoff; k++) t = __uint_as_float(__float_as_uint(t) - 1);
    for (int k = 0; k < 8 && __fmul_rn(t, w) < cutoff; k++)
    {
        if (__float_as_uint(t) >= 0x7F7FFFFFu) return __uint_as_float(0x7F800000u);
        t = __uint_as_float(__float_as_uint(t) + 1);
    }
    return t;
}
#endif

// One timed interval recorded with a pair of CUDA events on the worker's stream.
struct TimedSpan
{
    cudaEvent_t begin, end;
    int stat;   // index into Worker::statNames
};

struct StatAccum
{
    double sum = 0.0;
    double sum2 = 0.0;   // sum of squares, for the sample standard deviation the reference prints
    uint64_t count = 0;
};

// Device-side layout of the radix sort scratch (see sort.cu).
struct SortScratch
{
    void *base = nullptr;          // the allocation: [caller's prefix | hist | tickets | status]
    uint32_t *hist = nullptr;      // [passes][256] global digit histograms
    uint32_t *status = nullptr;    // [passes][tiles][256] decoupled look-back words
    uint32_t *tickets = nullptr;   // [passes] dynamic tile counters
    size_t statusWordsPerPass = 0;
    uint32_t maxPasses = 0;
    size_t maxItems = 0;
};

}  // namespace mlsb

struct mlsb_worker
{
    mlsb_worker_config cfg{};
    int device = 0;
    cudaStream_t stream = nullptr;      // compute (the cl::CommandQueue of workers.cpp:215)
    cudaStream_t copyStream = nullptr;  // H2D staging (CopyGroup's queue, workers.cpp:101)
    // The long MLS kernel runs on a low-priority stream of its own, fenced by two events, and
    // everything else on a high-priority one, so that with several workers per GPU the short
    // latency-bound kernels of one bin are not queued behind another bin's 30 000 MLS blocks
    // (C2, 8 workers: 297 -> 308 Msplats/s; MLSB_PRIORITY=0 puts everything on one stream)
    cudaStream_t mlsStream = nullptr;
    cudaEvent_t mlsBefore = nullptr, mlsAfter = nullptr;
    cudaEvent_t copyEvent = nullptr;
    bool copyPending = false;

    // ---- splats (DeviceWorkerGroup::WorkItem::splats) ----
    mlsb_splat *splats = nullptr;

    // ---- octree (SplatTreeCL members, src/splat_tree_cl.cpp:114-129) ----
    uint32_t numStartMax = 0;
    uint32_t numStart = 0;
    uint32_t numCommandsMax = 0;
    int32_t *start = nullptr;
    int32_t *runStart = nullptr;   // start slot of each non-empty key, else -1 (plays jumpPos' role too)
    int32_t *jumpPos = nullptr;
    int32_t *commands = nullptr;
    uint32_t *entryKeys[2] = {nullptr, nullptr};
    uint32_t *entryValues[2] = {nullptr, nullptr};
    int32_t *runFirst = nullptr, *runLast = nullptr;
    uint32_t *runIndex = nullptr;
    // scratch of the fused build, zeroed by ONE memset per bin together with the sort's (they share an
    // allocation): look-back words of the entry tiles, {ticket, numValid}, per-key entry counts
    uint64_t *entryStatus = nullptr;
    uint32_t entryTilesMax = 0;
    uint32_t *entryCounters = nullptr;   // [0] tile ticket, [1] number of accepted entries
    uint32_t *keyCount = nullptr;        // only when the octree is small enough for the cluster plan kernel
    uint32_t *tileBucket = nullptr;      // [entry tile][bucket]: entries per 256-key bucket, then their base in the bucketed layout
    float *hitThresholds = nullptr;      // per splat: the |p|^2 below which the splat counts (mls.cu), by the build
    uint64_t *scanStatus = nullptr;   // decoupled look-back words for device-wide scans
    uint32_t scanStatusWords = 0;
    uint32_t *scanTicket = nullptr;
    mlsb::SortScratch sort;

    // ---- MLS (MlsFunctor kernel arguments, src/mls.cpp:76-94) ----
    const mlsb_splat *mlsSplats = nullptr;
    const int32_t *mlsCommands = nullptr;
    const int32_t *mlsStart = nullptr;
    const float *mlsThresholds = nullptr;   // per-splat hit thresholds when the octree came from our build
    int32_t mlsOffset[3] = {0, 0, 0};
    uint32_t mlsStartShift = 0;
    float boundaryFactor = 0.0f;
    bool mlsBound = false;
    unsigned long long *mlsWork = nullptr;   // device: [0] pair tests, [1] hits (profiling only)

    // ---- marching (Marching members, src/marching.cpp:380-405) ----
    uint32_t fieldDim = 0;        // corners per axis the field was allocated for, rounded to 8
    float *field = nullptr;       // [z][y][x], pitch fieldDim
    uint8_t *blockActive = nullptr;   // [bz][by][bx] per 8x8x8 block of the field: written by the MLS kernel
    uint64_t vertexSpace = 0, indexSpace = 0;
    uint32_t *latMask = nullptr;      // lattice of edge midpoints: 1 bit per possible vertex of a block
    uint4 *latRecord = nullptr;       // per non-empty lattice word: (mask, welded vertices in front of it: unflagged, flagged, -)
    uint64_t *latStatus = nullptr;    // look-back words of the lattice scan; they follow marchCounters in one allocation
    uint32_t latStatusWords = 0;
    size_t latticeWords = 0;
    uint32_t *indices = nullptr;
    float *welded = nullptr;          // packed float3
    uint64_t *weldedKeys = nullptr;
    float *weldedFlagged = nullptr;   // vertices carrying the external flag: they follow all others in the output
    uint64_t *weldedKeysFlagged = nullptr;
    uint32_t *marchCounters = nullptr;   // device: see marching.cu
    uint64_t *tileTotals = nullptr;      // per marching tile: totals, then exclusive bases
    uint32_t *tileList = nullptr;        // non-empty marching tiles
    uint32_t maxTiles = 0;
    float scale = 1.0f;
    float bias[3] = {0.0f, 0.0f, 0.0f};
    void *hostBlob = nullptr;     // pinned readback buffer
    size_t hostBlobBytes = 0;
    uint32_t *hostCounters = nullptr;   // pinned
    uint8_t *tables = nullptr;          // device copy of the marching LUTs

    // MeshFilterChain (src/mesh_filter.h:69-116): device-side filters between the marching output and the read-back
    std::vector<std::pair<mlsb_mesh_filter, void *>> meshFilters;

    // ---- statistics ----
    int profiling = 0;        // 0 off, 1 processCorners only, 2 every kernel
    bool countWork = false;   // MLS kernels also count pair tests / hits (slower; never timed)
    std::vector<std::string> statNames;
    std::map<std::string, int> statIndex;
    std::vector<mlsb::StatAccum> stats;
    std::vector<mlsb::TimedSpan> pendingSpans;
    std::vector<cudaEvent_t> eventPool;
    uint64_t launches = 0;
    uint32_t shipouts = 0;    // blocks shipped by the current generate call (statistic marching.shipouts)
    uint64_t marchWork[5] = {0, 0, 0, 0, 0};   // corners, cells, unwelded vertices, indices, welded vertices
    // host waits: spin (lowest latency) or block on an event (frees the core; MLSB_SYNC=block)
    bool blockingSync = false;
    bool yieldingSync = false;
    // mesh read-back by a kernel writing pinned host memory (one host wait per block) instead of the copy engine
    // (two waits, no SM): MLSB_READBACK=pack, off by default (see packMeshKernel for the measurements)
    bool packReadback = false;
    // MLSB_READBACK=compact: the block crosses PCIe in a 48-byte-per-splat wire format (8 bytes per vertex, 8 per
    // triangle instead of 12 and 12) and the worker thread expands it into the HostKeyMesh blob (unpack.cpp)
    bool compactReadback = false;
    uint2 *compactVerts = nullptr, *compactVertsFlagged = nullptr;   // (t, lattice point) per vertex, by rank
    uint64_t *wireBlob = nullptr;      // device: [external keys][vertices][triangles], one 64-bit word each
    size_t wireWords = 0;
    void *hostWire = nullptr;          // pinned landing buffer of the wire blob (grows on demand)
    size_t hostWireBytes = 0;
    // MLSB_SYNC=yield: poll cudaStreamQuery with sched_yield between polls
    cudaEvent_t syncEvent = nullptr;
};

namespace mlsb
{

// RAII helper: times everything enqueued on the worker stream during its lifetime under a
// reference statistic name (Statistics::timeEvent, src/statistics_cl.cpp:62-160).
class ScopedStat
{
public:
    ScopedStat(mlsb_worker *w, const char *name, int kernels = 1, int level = 2);
    ~ScopedStat();

private:
    mlsb_worker *w;
    int idx;
    cudaEvent_t begin;
};

int statId(mlsb_worker *w, const char *name);
// One host-side sample of a named statistic (marching.shipouts, marching.overflow: src/marching.cpp:350-352).
void addSample(mlsb_worker *w, const char *name, double value);
int waitStream(mlsb_worker *w);   // host wait for the worker's compute stream
void collectStats(mlsb_worker *w);   // requires the stream to be idle

// sort.cu
// prefixBytes: room in front of the sort's own scratch, in the same allocation (*prefix), for state the
// caller wants zeroed by the same memset
int sortScratchAlloc(SortScratch &s, size_t maxItems, uint32_t maxPasses, size_t prefixBytes = 0, void **prefix = nullptr);
void sortScratchFree(SortScratch &s);
// Stable LSD radix sort of (key,value) u32 pairs on the low `bits` bits.  Input in keys[0] /
// values[0]; returns in *result which of the two buffers holds the output.
int radixSortPairs(mlsb_worker *w, uint32_t *keys[2], uint32_t *values[2], uint32_t n, uint32_t bits,
                   int *result, const char *statName);
// Same, with the element count min(*nDev, n) read on the device (grids are sized for n).
// identityValues: values[0] is not read; the input value of element i is i.
int radixSortPairsDeviceCount(mlsb_worker *w, uint32_t *keys[2], uint32_t *values[2], uint32_t n,
                              const uint32_t *nDev, uint32_t bits, int *result, const char *statName,
                              bool identityValues);

// The last pass of a sort whose digit histograms were written by the caller (s.hist) into scratch the
// caller has zeroed: instead of the sorted pairs it writes the octree's command list,
// commands[1 + position + 2 * runIndex[key]] = value.  n = host-side bound, *nDev the real count.
int radixSortToCommands(mlsb_worker *w, uint32_t *keys[2], uint32_t *values[2], uint32_t n, const uint32_t *nDev,
                        uint32_t bits, int32_t *commands, const uint32_t *runIndex, const char *statName);
size_t sortScratchBytes(size_t maxItems, uint32_t maxPasses);   // what sortScratchAlloc allocates behind the prefix
// bytes of sort scratch (from s.hist on) a sort of n items with `passes` passes touches
size_t sortScratchBytesUsed(const SortScratch &s, uint32_t n, uint32_t passes);

// octree.cu
int octreeAlloc(mlsb_worker *w);
void octreeFree(mlsb_worker *w);
int octreeBuild(mlsb_worker *w, uint32_t firstSplat, uint32_t numSplats, const int32_t offset[3],
                uint32_t subsamplingShift);

// mls.cu
int measureFp32Peak(double *laneInstPerSecond, double *seconds);
int mlsAlloc(mlsb_worker *w);
void mlsFree(mlsb_worker *w);
int mlsCorners(mlsb_worker *w, float *field, uint32_t pitch, const mlsb_swathe &sw, bool recordActive = false);

// marching.cu
int marchingAlloc(mlsb_worker *w);
void marchingFree(mlsb_worker *w);
// where the field comes from (MlsFunctor bound to the worker / a host array / any generator) and where
// the blocks go (host blobs / device meshes)
struct MarchIo
{
    bool useMls = false;
    const float *hostField = nullptr;
    mlsb_field_generator generate = nullptr;
    void *generateUser = nullptr;
    mlsb_mesh_sink sink = nullptr;
    void *user = nullptr;
    mlsb_device_mesh_sink deviceSink = nullptr;
    void *deviceUser = nullptr;
};
int marchingGenerate(mlsb_worker *w, const MarchIo &io, const uint32_t size[3], const uint32_t keyOffset[3]);
int runMeshFilters(mlsb_worker *w, mlsb_device_mesh &mesh);
// MLSB_READBACK: 0 copy engine (default), 1 pack kernel into pinned memory, 2 compact wire format + host unpack
int readbackModeFromEnv();

int marchingReadDeviceMesh(mlsb_worker *w, const mlsb_device_mesh *mesh, void *hostBlob);
uint64_t marchingDeviceBytes(const mlsb_worker_config &c, uint64_t *maxAllocation, uint64_t *pinnedHost);
uint64_t octreeDeviceBytes(const mlsb_worker_config &c, uint64_t *maxAllocation);
int marchingTablesHost(uint8_t *count, uint16_t *start, uint8_t *data, uint32_t *key);

}  // namespace mlsb
