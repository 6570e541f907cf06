// extern "C" entry points declared in include/mlsb200.h, worker lifetime and statistics.
#include "common.cuh"
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <new>
#include <sched.h>

namespace mlsb
{

static thread_local char gError[1024] = "";

void setError(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(gError, sizeof(gError), fmt, ap);
    va_end(ap);
}

int statId(mlsb_worker *w, const char *name)
{
    auto it = w->statIndex.find(name);
    if (it != w->statIndex.end()) return it->second;
    int id = (int) w->statNames.size();
    w->statNames.push_back(name);
    w->stats.emplace_back();
    w->statIndex[name] = id;
    return id;
}

void addSample(mlsb_worker *w, const char *name, double value)
{
    StatAccum &a = w->stats[statId(w, name)];
    a.sum += value;
    a.sum2 += value * value;
    a.count++;
}

static cudaEvent_t takeEvent(mlsb_worker *w)
{
    if (!w->eventPool.empty())
    {
        cudaEvent_t e = w->eventPool.back();
        w->eventPool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

ScopedStat::ScopedStat(mlsb_worker *w, const char *name, int, int level) : w(w), idx(-1), begin(nullptr)
{
    if (w->profiling < level) return;
    idx = statId(w, name);
    begin = takeEvent(w);
    cudaEventRecord(begin, w->stream);
}

ScopedStat::~ScopedStat()
{
    if (idx < 0) return;
    cudaEvent_t end = takeEvent(w);
    cudaEventRecord(end, w->stream);
    w->pendingSpans.push_back(TimedSpan{begin, end, idx});
}

void collectStats(mlsb_worker *w)
{
    for (const TimedSpan &sp : w->pendingSpans)
    {
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, sp.begin, sp.end) == cudaSuccess)
        {
            const double seconds = ms * 1e-3;
            w->stats[sp.stat].sum += seconds;
            w->stats[sp.stat].sum2 += seconds * seconds;
            w->stats[sp.stat].count++;
        }
        w->eventPool.push_back(sp.begin);
        w->eventPool.push_back(sp.end);
    }
    w->pendingSpans.clear();
}

int waitStream(mlsb_worker *w)
{
    if (w->blockingSync)
    {
        MLSB_CUDA(cudaEventRecord(w->syncEvent, w->stream));
        MLSB_CUDA(cudaEventSynchronize(w->syncEvent));
    }
    else if (w->yieldingSync)
    {
        // poll, giving the core away between polls: for more worker threads than cores
        cudaError_t e;
        while ((e = cudaStreamQuery(w->stream)) == cudaErrorNotReady) sched_yield();
        MLSB_CUDA(e);
    }
    else
        MLSB_CUDA(cudaStreamSynchronize(w->stream));
    return MLSB_OK;
}

// src/mls.cpp:137-144
static float boundaryFactorOf(float limit)
{
    const float pi = 3.14159265358979323846f;
    const float boundaryScale = (sqrtf(6.0f) * 512) / (693 * pi);
    const float gamma = boundaryScale * limit;
    return 1.0f - gamma * gamma;
}

static int validateConfig(const mlsb_worker_config *c)
{
    MLSB_REQUIRE(c != nullptr, MLSB_ERR_INVALID_ARGUMENT, "null config");
    // src/splat_tree_cl.cpp:63-64, :104-105
    MLSB_REQUIRE(c->maxLevels >= 1 && c->maxLevels <= (uint32_t) kMaxLevels, MLSB_ERR_LENGTH, "maxLevels out of range");
    MLSB_REQUIRE(c->maxSplats >= 1 && c->maxSplats <= (1u << 31) / 16, MLSB_ERR_LENGTH, "maxSplats out of range");
    MLSB_REQUIRE(c->subsampling + c->maxLevels - 1 < 32, MLSB_ERR_INVALID_ARGUMENT, "levels + subsampling too large");
    MLSB_REQUIRE(c->shape == MLSB_SHAPE_SPHERE || c->shape == MLSB_SHAPE_PLANE, MLSB_ERR_INVALID_ARGUMENT, "unknown shape");
    MLSB_REQUIRE(c->maxCells >= 1, MLSB_ERR_INVALID_ARGUMENT, "maxCells must be positive");
    // src/marching.cpp:311-315 (the reference's MAX_DIMENSION is 8192; the lattice words and tile layout here
    // carry 512 corners per axis), :316-318 (mesh memory for at least one slice of worst-case cells)
    MLSB_REQUIRE(c->maxCells + 1 <= MLSB_MAX_DIMENSION, MLSB_ERR_INVALID_ARGUMENT, "maxCells + 1 must be in [2, 512]");
    MLSB_REQUIRE(c->meshBytes >= (uint64_t) c->maxCells * c->maxCells * kMaxCellBytes, MLSB_ERR_INVALID_ARGUMENT,
                 "meshBytes smaller than one slice of worst-case cells");
    MLSB_REQUIRE(c->meshBytes / kMaxCellBytes * kMaxCellIndices < (1ull << 31), MLSB_ERR_LENGTH, "meshBytes too large");
    return MLSB_OK;
}

// Copy of device-resident splats into a worker's buffer (mlsb_splats_upload_async): 16 bytes per thread and
// iteration, four independent loads in flight.
__global__ void __launch_bounds__(256) copyWordsKernel(uint4 *__restrict__ dst, const uint4 *__restrict__ src, uint32_t words)
{
    const uint32_t base = blockIdx.x * 1024 + threadIdx.x;
    uint4 v[4];
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (base + k * 256 < words) v[k] = src[base + k * 256];
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (base + k * 256 < words) dst[base + k * 256] = v[k];
}

}  // namespace mlsb

using namespace mlsb;

extern "C" {

int mlsb_device_count(int *count)
{
    MLSB_REQUIRE(count != nullptr, MLSB_ERR_INVALID_ARGUMENT, "null pointer");
    *count = 0;
    MLSB_CUDA(cudaGetDeviceCount(count));
    return MLSB_OK;
}

const char *mlsb_last_error_string(void) { return gError; }

const char *mlsb_version(void) { return "mlsb200 0.1 (sm_100a)"; }

// words of the scan scratch shared by the general octree path (computed in one place for create and usage)
static uint64_t scanStatusWordsFor(const mlsb_worker_config *c)
{
    const uint64_t maxStart = (1ull << (3 * c->maxLevels)) / 7;
    return std::max<uint64_t>(divUp<uint64_t>(maxStart, 2048) + 2, divUp<uint64_t>(c->maxSplats, 256) + 2);
}

int mlsb_worker_resource_usage_ex(const mlsb_worker_config *c, mlsb_resource_usage *usage)
{
    MLSB_TRY(validateConfig(c));
    MLSB_REQUIRE(usage != nullptr, MLSB_ERR_INVALID_ARGUMENT, "null pointer");
    uint64_t maxOct = 0, maxMarch = 0, pinned = 0;
    uint64_t total = (uint64_t) c->maxSplats * sizeof(mlsb_splat);                  // DeviceWorkerGroup::WorkItem::splats
    total += octreeDeviceBytes(*c, &maxOct);
    total += marchingDeviceBytes(*c, &maxMarch, &pinned);
    total += 2 * sizeof(unsigned long long);                                         // MLS work counters
    total += scanStatusWordsFor(c) * sizeof(uint64_t) + 4 * sizeof(uint32_t);        // scan status + tickets
    usage->deviceBytes = total;
    usage->maxAllocationBytes = std::max<uint64_t>({maxOct, maxMarch, (uint64_t) c->maxSplats * sizeof(mlsb_splat)});
    usage->pinnedHostBytes = pinned;
    return MLSB_OK;
}

int mlsb_worker_resource_usage(const mlsb_worker_config *c, uint64_t *deviceBytes)
{
    MLSB_REQUIRE(deviceBytes != nullptr, MLSB_ERR_INVALID_ARGUMENT, "null pointer");
    mlsb_resource_usage u;
    MLSB_TRY(mlsb_worker_resource_usage_ex(c, &u));
    *deviceBytes = u.deviceBytes;
    return MLSB_OK;
}

int mlsb_validate_device(int device)
{
    int ndev = 0;
    MLSB_CUDA(cudaGetDeviceCount(&ndev));
    MLSB_REQUIRE(device >= 0 && device < ndev, MLSB_ERR_INVALID_ARGUMENT, "device ordinal out of range");
    cudaDeviceProp p;
    MLSB_CUDA(cudaGetDeviceProperties(&p, device));
    if (p.major != 10)
    {
        setError("device %d (%s) has compute capability %d.%d: this library is built for sm_100a only", device, p.name, p.major, p.minor);
        return MLSB_ERR_CUDA;
    }
    if (p.sharedMemPerBlockOptin < 96 * 1024)
    {
        setError("device %d offers %zu bytes of shared memory per block, the large sort tiles need 96 KiB", device, p.sharedMemPerBlockOptin);
        return MLSB_ERR_CUDA;
    }
    int clusters = 0;
    MLSB_CUDA(cudaDeviceGetAttribute(&clusters, cudaDevAttrClusterLaunch, device));
    if (!clusters)
    {
        setError("device %d cannot launch thread-block clusters (the octree's plan kernel is one)", device);
        return MLSB_ERR_CUDA;
    }
    return MLSB_OK;
}

int mlsb_worker_create(const mlsb_worker_config *config, mlsb_worker **out)
{
    MLSB_REQUIRE(out != nullptr, MLSB_ERR_INVALID_ARGUMENT, "null pointer");
    *out = nullptr;
    MLSB_TRY(validateConfig(config));
    int ndev = 0;
    MLSB_CUDA(cudaGetDeviceCount(&ndev));
    MLSB_REQUIRE(ndev > 0, MLSB_ERR_CUDA, "no CUDA device: mlsb200 has no CPU fallback");
    MLSB_REQUIRE(config->device >= 0 && config->device < ndev, MLSB_ERR_INVALID_ARGUMENT, "device ordinal out of range");
    MLSB_TRY(mlsb_validate_device(config->device));
    MLSB_CUDA(cudaSetDevice(config->device));

    mlsb_worker *w = new (std::nothrow) mlsb_worker();
    MLSB_REQUIRE(w != nullptr, MLSB_ERR_LENGTH, "out of host memory");
    w->cfg = *config;
    w->device = config->device;
    w->boundaryFactor = boundaryFactorOf(config->boundaryLimit);
    int st = MLSB_OK;
    auto fail = [&](int code) { mlsb_worker_destroy(w); return code; };
    const char *priority = getenv("MLSB_PRIORITY");
    const bool split = priority == nullptr || strcmp(priority, "0") != 0;   // on unless MLSB_PRIORITY=0
    int least = 0, greatest = 0;
    if (split) cudaDeviceGetStreamPriorityRange(&least, &greatest);
    if (cudaStreamCreateWithPriority(&w->stream, cudaStreamNonBlocking, split ? greatest : 0) != cudaSuccess
        || (split && (cudaStreamCreateWithPriority(&w->mlsStream, cudaStreamNonBlocking, least) != cudaSuccess
                      || cudaEventCreateWithFlags(&w->mlsBefore, cudaEventDisableTiming) != cudaSuccess
                      || cudaEventCreateWithFlags(&w->mlsAfter, cudaEventDisableTiming) != cudaSuccess))
        || cudaStreamCreateWithFlags(&w->copyStream, cudaStreamNonBlocking) != cudaSuccess
        || cudaEventCreateWithFlags(&w->copyEvent, cudaEventDisableTiming) != cudaSuccess)
    {
        setError("stream creation failed: %s", cudaGetErrorString(cudaGetLastError()));
        return fail(MLSB_ERR_CUDA);
    }
    {
        const char *mode = getenv("MLSB_SYNC");
        w->blockingSync = mode != nullptr && strcmp(mode, "block") == 0;
        w->yieldingSync = mode != nullptr && strcmp(mode, "yield") == 0;
        const char *readback = getenv("MLSB_READBACK");
        w->packReadback = readback != nullptr && strcmp(readback, "pack") == 0;
        if (cudaEventCreateWithFlags(&w->syncEvent, cudaEventDisableTiming | cudaEventBlockingSync) != cudaSuccess)
        {
            setError("event creation failed: %s", cudaGetErrorString(cudaGetLastError()));
            return fail(MLSB_ERR_CUDA);
        }
    }
    if (cudaMalloc(&w->splats, (size_t) config->maxSplats * sizeof(mlsb_splat)) != cudaSuccess)
    {
        setError("cudaMalloc(splats) failed: %s", cudaGetErrorString(cudaGetLastError()));
        return fail(MLSB_ERR_CUDA);
    }
    if ((st = octreeAlloc(w)) != MLSB_OK) return fail(st);
    if ((st = mlsAlloc(w)) != MLSB_OK) return fail(st);
    if ((st = marchingAlloc(w)) != MLSB_OK) return fail(st);
    // shared scan scratch (the sort's lives with the octree's, see octreeAlloc)
    const uint64_t words = scanStatusWordsFor(config);
    w->scanStatusWords = (uint32_t) words;
    if (cudaMalloc(&w->scanStatus, words * sizeof(uint64_t)) != cudaSuccess
        || cudaMalloc(&w->scanTicket, 4 * sizeof(uint32_t)) != cudaSuccess)
    {
        setError("cudaMalloc(scan) failed: %s", cudaGetErrorString(cudaGetLastError()));
        return fail(MLSB_ERR_CUDA);
    }
    *out = w;
    return MLSB_OK;
}

void mlsb_worker_destroy(mlsb_worker *w)
{
    if (!w) return;
    cudaSetDevice(w->device);
    if (w->stream) cudaStreamSynchronize(w->stream);
    if (w->copyStream) cudaStreamSynchronize(w->copyStream);
    if (w->mlsStream) cudaStreamSynchronize(w->mlsStream);
    collectStats(w);
    for (cudaEvent_t e : w->eventPool) cudaEventDestroy(e);
    cudaFree(w->splats);
    octreeFree(w);
    mlsFree(w);
    marchingFree(w);
    cudaFree(w->scanStatus);
    cudaFree(w->scanTicket);
    if (w->copyEvent) cudaEventDestroy(w->copyEvent);
    if (w->syncEvent) cudaEventDestroy(w->syncEvent);
    if (w->mlsBefore) cudaEventDestroy(w->mlsBefore);
    if (w->mlsAfter) cudaEventDestroy(w->mlsAfter);
    if (w->stream) cudaStreamDestroy(w->stream);
    if (w->mlsStream) cudaStreamDestroy(w->mlsStream);
    if (w->copyStream) cudaStreamDestroy(w->copyStream);
    delete w;
}

void *mlsb_worker_stream(mlsb_worker *w) { return w ? (void *) w->stream : nullptr; }

int mlsb_worker_synchronize(mlsb_worker *w)
{
    MLSB_REQUIRE(w != nullptr, MLSB_ERR_INVALID_ARGUMENT, "null worker");
    MLSB_CUDA(cudaStreamSynchronize(w->copyStream));
    MLSB_CUDA(cudaStreamSynchronize(w->stream));
    w->copyPending = false;
    collectStats(w);
    return MLSB_OK;
}

int mlsb_host_alloc(void **ptr, size_t bytes)
{
    MLSB_REQUIRE(ptr != nullptr, MLSB_ERR_INVALID_ARGUMENT, "null pointer");
    MLSB_CUDA(cudaMallocHost(ptr, bytes ? bytes : 1));
    return MLSB_OK;
}

int mlsb_host_free(void *ptr)
{
    MLSB_CUDA(cudaFreeHost(ptr));
    return MLSB_OK;
}

int mlsb_splats_upload_async(mlsb_worker *w, const mlsb_splat *host, uint32_t dstFirst, uint32_t n)
{
    MLSB_REQUIRE(w != nullptr && (host != nullptr || n == 0), MLSB_ERR_INVALID_ARGUMENT, "null pointer");
    MLSB_REQUIRE((uint64_t) dstFirst + n <= w->cfg.maxSplats, MLSB_ERR_LENGTH, "more splats than the worker's buffer holds");
    MLSB_CUDA(cudaSetDevice(w->device));
    if (n)
    {
        // A source that already lives in device memory is copied by a kernel: a device-to-device cudaMemcpyAsync
        // may be queued on the copy engine that is busy with other workers' mesh read-backs at PCIe speed (the
        // HBM-resident bench pass was SLOWER than the pass from pinned host memory: 33.2 against 26.8 ms per step
        // on two GPUs), while the host-to-device engine is otherwise idle.
        cudaPointerAttributes attr;
        const bool onDevice = cudaPointerGetAttributes(&attr, host) == cudaSuccess && attr.type == cudaMemoryTypeDevice;
        if (!onDevice) (void) cudaGetLastError();   // unregistered host memory may report an error: not one of ours
        if (onDevice && (reinterpret_cast<uintptr_t>(host) & 15) == 0)
        {
            const uint32_t words = n * (uint32_t) (sizeof(mlsb_splat) / sizeof(uint4));
            copyWordsKernel<<<(words + 1023) / 1024, 256, 0, w->copyStream>>>(reinterpret_cast<uint4 *>(w->splats + dstFirst),
                                                                              reinterpret_cast<const uint4 *>(host), words);
            w->launches++;
            MLSB_CUDA(cudaGetLastError());
        }
        else
            MLSB_CUDA(cudaMemcpyAsync(w->splats + dstFirst, host, (size_t) n * sizeof(mlsb_splat), cudaMemcpyDefault,
                                      w->copyStream));
    }
    MLSB_CUDA(cudaEventRecord(w->copyEvent, w->copyStream));
    w->copyPending = true;
    return MLSB_OK;
}

int mlsb_splats_device(mlsb_worker *w, mlsb_splat **devicePtr)
{
    MLSB_REQUIRE(w != nullptr && devicePtr != nullptr, MLSB_ERR_INVALID_ARGUMENT, "null pointer");
    *devicePtr = w->splats;
    return MLSB_OK;
}

int mlsb_octree_build(mlsb_worker *w, uint32_t firstSplat, uint32_t numSplats, const uint32_t size[3],
                      const int32_t offset[3], uint32_t subsamplingShift)
{
    MLSB_REQUIRE(w != nullptr && size != nullptr && offset != nullptr, MLSB_ERR_INVALID_ARGUMENT, "null pointer");
    // src/splat_tree_cl.cpp:277-281
    MLSB_REQUIRE(numSplats <= w->cfg.maxSplats, MLSB_ERR_LENGTH, "numSplats exceeds maxSplats");
    MLSB_REQUIRE((uint64_t) firstSplat + numSplats <= w->cfg.maxSplats, MLSB_ERR_LENGTH, "splat range outside the buffer");
    MLSB_REQUIRE(w->cfg.maxLevels + subsamplingShift - 1 < 32, MLSB_ERR_INVALID_ARGUMENT, "levels + subsampling too large");
    uint64_t maxSize = 1ull << (w->cfg.maxLevels + subsamplingShift - 1);
    MLSB_REQUIRE(size[0] <= maxSize && size[1] <= maxSize && size[2] <= maxSize, MLSB_ERR_LENGTH, "region larger than the octree");
    MLSB_CUDA(cudaSetDevice(w->device));
    if (w->copyPending)
    {
        MLSB_CUDA(cudaStreamWaitEvent(w->stream, w->copyEvent, 0));
        w->copyPending = false;
    }
    return octreeBuild(w, firstSplat, numSplats, offset, subsamplingShift);
}

int mlsb_octree_get(mlsb_worker *w, const int32_t **start, uint32_t *numStart, const int32_t **commands,
                    uint32_t *numCommands, uint32_t *numLevels)
{
    MLSB_REQUIRE(w != nullptr, MLSB_ERR_INVALID_ARGUMENT, "null worker");
    if (start) *start = w->start;
    if (numStart) *numStart = w->numStart;
    if (commands) *commands = w->commands;
    if (numCommands) *numCommands = w->numCommandsMax;
    if (numLevels) *numLevels = w->cfg.maxLevels;
    return MLSB_OK;
}

int mlsb_mls_set(mlsb_worker *w, const int32_t offset[3], uint32_t subsamplingShift)
{
    MLSB_REQUIRE(w != nullptr, MLSB_ERR_INVALID_ARGUMENT, "null worker");
    MLSB_TRY(mlsb_mls_set_buffers(w, offset, w->splats, w->commands, w->start, subsamplingShift));
    w->mlsThresholds = w->hitThresholds;   // written by the octree build for exactly these splats
    return MLSB_OK;
}

int mlsb_mls_set_buffers(mlsb_worker *w, const int32_t offset[3], const mlsb_splat *splats, const int32_t *commands,
                         const int32_t *start, uint32_t subsamplingShift)
{
    MLSB_REQUIRE(w != nullptr && offset != nullptr && splats != nullptr && commands != nullptr && start != nullptr,
                 MLSB_ERR_INVALID_ARGUMENT, "null pointer");
    // MlsFunctor::subsamplingMin, src/mls.cpp:54: one start[] lookup per 8x8x8 work-group
    MLSB_REQUIRE(subsamplingShift >= 3, MLSB_ERR_INVALID_ARGUMENT, "subsampling must be at least 3");
    w->mlsSplats = splats;
    w->mlsCommands = commands;
    w->mlsStart = start;
    w->mlsThresholds = nullptr;
    memcpy(w->mlsOffset, offset, sizeof(w->mlsOffset));
    w->mlsStartShift = 3 * subsamplingShift;
    w->mlsBound = true;
    return MLSB_OK;
}

int mlsb_mls_set_boundary_limit(mlsb_worker *w, float limit)
{
    MLSB_REQUIRE(w != nullptr, MLSB_ERR_INVALID_ARGUMENT, "null worker");
    w->boundaryFactor = boundaryFactorOf(limit);
    return MLSB_OK;
}

int mlsb_mls_corners(mlsb_worker *w, float *deviceField, uint32_t pitch, const mlsb_swathe *swathe)
{
    MLSB_REQUIRE(w != nullptr && deviceField != nullptr && swathe != nullptr, MLSB_ERR_INVALID_ARGUMENT, "null pointer");
    MLSB_CUDA(cudaSetDevice(w->device));
    return mlsCorners(w, deviceField, pitch, *swathe);
}

int mlsb_set_scale_bias(mlsb_worker *w, float scale, const float bias[3])
{
    MLSB_REQUIRE(w != nullptr && bias != nullptr, MLSB_ERR_INVALID_ARGUMENT, "null pointer");
    w->scale = scale;
    memcpy(w->bias, bias, sizeof(w->bias));
    return MLSB_OK;
}

int mlsb_marching_generate(mlsb_worker *w, const uint32_t size[3], const uint32_t keyOffset[3], mlsb_mesh_sink sink,
                           void *user)
{
    MLSB_REQUIRE(w != nullptr && size != nullptr && keyOffset != nullptr, MLSB_ERR_INVALID_ARGUMENT, "null pointer");
    MLSB_CUDA(cudaSetDevice(w->device));
    MarchIo io;
    io.useMls = true;
    io.sink = sink;
    io.user = user;
    int st = marchingGenerate(w, io, size, keyOffset);
    collectStats(w);
    return st;
}

int mlsb_marching_generate_host_field(mlsb_worker *w, const float *hostField, const uint32_t size[3],
                                      const uint32_t keyOffset[3], mlsb_mesh_sink sink, void *user)
{
    MLSB_REQUIRE(w != nullptr && size != nullptr && keyOffset != nullptr, MLSB_ERR_INVALID_ARGUMENT, "null pointer");
    MLSB_CUDA(cudaSetDevice(w->device));
    MarchIo io;
    io.hostField = hostField;
    io.sink = sink;
    io.user = user;
    int st = marchingGenerate(w, io, size, keyOffset);
    collectStats(w);
    return st;
}

int mlsb_marching_generate_cb(mlsb_worker *w, mlsb_field_generator generate, void *generatorUser, const uint32_t size[3],
                              const uint32_t keyOffset[3], mlsb_mesh_sink sink, void *user)
{
    MLSB_REQUIRE(w != nullptr && generate != nullptr && size != nullptr && keyOffset != nullptr, MLSB_ERR_INVALID_ARGUMENT, "null pointer");
    MLSB_CUDA(cudaSetDevice(w->device));
    MarchIo io;
    io.generate = generate;
    io.generateUser = generatorUser;
    io.sink = sink;
    io.user = user;
    int st = marchingGenerate(w, io, size, keyOffset);
    collectStats(w);
    return st;
}

int mlsb_marching_generate_device(mlsb_worker *w, mlsb_field_generator generate, void *generatorUser, const uint32_t size[3],
                                  const uint32_t keyOffset[3], mlsb_device_mesh_sink sink, void *user)
{
    MLSB_REQUIRE(w != nullptr && sink != nullptr && size != nullptr && keyOffset != nullptr, MLSB_ERR_INVALID_ARGUMENT, "null pointer");
    MLSB_CUDA(cudaSetDevice(w->device));
    MarchIo io;
    io.useMls = generate == nullptr;
    io.generate = generate;
    io.generateUser = generatorUser;
    io.deviceSink = sink;
    io.deviceUser = user;
    int st = marchingGenerate(w, io, size, keyOffset);
    collectStats(w);
    return st;
}

int mlsb_worker_add_mesh_filter(mlsb_worker *w, mlsb_mesh_filter filter, void *user)
{
    MLSB_REQUIRE(w != nullptr && filter != nullptr, MLSB_ERR_INVALID_ARGUMENT, "null pointer");
    w->meshFilters.push_back(std::make_pair(filter, user));
    return MLSB_OK;
}

int mlsb_worker_clear_mesh_filters(mlsb_worker *w)
{
    MLSB_REQUIRE(w != nullptr, MLSB_ERR_INVALID_ARGUMENT, "null worker");
    w->meshFilters.clear();
    return MLSB_OK;
}

int mlsb_mesh_filter_chain_apply(mlsb_worker *w, const mlsb_device_mesh *mesh, mlsb_device_mesh_sink sink, void *user)
{
    MLSB_REQUIRE(w != nullptr && mesh != nullptr && sink != nullptr, MLSB_ERR_INVALID_ARGUMENT, "null pointer");
    MLSB_REQUIRE(mesh->sizes.numInternalVertices <= mesh->sizes.numVertices, MLSB_ERR_INVALID_ARGUMENT,
                 "more internal vertices than vertices");
    MLSB_CUDA(cudaSetDevice(w->device));
    mlsb_device_mesh current = *mesh;
    MLSB_TRY(runMeshFilters(w, current));
    const int rc = sink(&current, (void *) w->stream, user);
    MLSB_TRY(waitStream(w));
    if (rc != 0)
    {
        setError("the device mesh sink returned %d", rc);
        return MLSB_ERR_STATE;
    }
    return MLSB_OK;
}

int mlsb_device_mesh_read(mlsb_worker *w, const mlsb_device_mesh *mesh, void *hostBlob)
{
    MLSB_REQUIRE(w != nullptr && mesh != nullptr && hostBlob != nullptr, MLSB_ERR_INVALID_ARGUMENT, "null pointer");
    MLSB_CUDA(cudaSetDevice(w->device));
    return marchingReadDeviceMesh(w, mesh, hostBlob);
}

uint64_t mlsb_mesh_host_bytes(const mlsb_mesh_sizes *s)
{
    // MeshSizes::getHostBytes, src/mesh.h:76-81
    return 3 * sizeof(float) * s->numVertices + 3 * sizeof(uint32_t) * s->numTriangles
           + sizeof(uint64_t) * (s->numVertices - s->numInternalVertices);
}

int mlsb_marching_tables(uint8_t *count, uint16_t *start, uint8_t *data, uint32_t *key)
{
    return marchingTablesHost(count, start, data, key);
}

int mlsb_worker_process_bin(mlsb_worker *w, uint32_t firstSplat, uint32_t numSplats, const uint32_t size[3],
                            const int32_t offset[3], mlsb_mesh_sink sink, void *user)
{
    MLSB_REQUIRE(w != nullptr && size != nullptr && offset != nullptr, MLSB_ERR_INVALID_ARGUMENT, "null pointer");
    // src/workers.cpp:237-261
    uint32_t expanded[3], keyOffset[3];
    for (int i = 0; i < 3; i++)
    {
        MLSB_REQUIRE(offset[i] >= 0, MLSB_ERR_INVALID_ARGUMENT, "bin offset must be non-negative");
        expanded[i] = roundUp(size[i], (uint32_t) kMlsWgs);
        keyOffset[i] = (uint32_t) offset[i];
    }
    MLSB_TRY(mlsb_octree_build(w, firstSplat, numSplats, expanded, offset, w->cfg.subsampling));
    MLSB_TRY(mlsb_mls_set(w, offset, w->cfg.subsampling));
    return mlsb_marching_generate(w, size, keyOffset, sink, user);
}

int mlsb_worker_set_profiling(mlsb_worker *w, int enabled)
{
    MLSB_REQUIRE(w != nullptr, MLSB_ERR_INVALID_ARGUMENT, "null worker");
    MLSB_CUDA(cudaStreamSynchronize(w->stream));
    collectStats(w);
    w->profiling = enabled < 0 ? 0 : (enabled > 2 ? 2 : enabled);
    return MLSB_OK;
}

int mlsb_worker_get_statistic(mlsb_worker *w, const char *name, double *sum, uint64_t *count)
{
    MLSB_REQUIRE(w != nullptr && name != nullptr, MLSB_ERR_INVALID_ARGUMENT, "null pointer");
    MLSB_CUDA(cudaStreamSynchronize(w->stream));
    collectStats(w);
    auto it = w->statIndex.find(name);
    double s = 0.0;
    uint64_t c = 0;
    if (it != w->statIndex.end())
    {
        s = w->stats[it->second].sum;
        c = w->stats[it->second].count;
    }
    if (sum) *sum = s;
    if (count) *count = c;
    return MLSB_OK;
}

int mlsb_worker_visit_statistics(mlsb_worker *w, mlsb_statistic_visitor visit, void *user)
{
    MLSB_REQUIRE(w != nullptr && visit != nullptr, MLSB_ERR_INVALID_ARGUMENT, "null pointer");
    MLSB_CUDA(cudaStreamSynchronize(w->stream));
    collectStats(w);
    for (const auto &entry : w->statIndex)   // name order
    {
        const StatAccum &a = w->stats[entry.second];
        if (a.count) visit(user, entry.first.c_str(), a.sum, a.sum2, a.count);
    }
    return MLSB_OK;
}

int mlsb_worker_reset_statistics(mlsb_worker *w)
{
    MLSB_REQUIRE(w != nullptr, MLSB_ERR_INVALID_ARGUMENT, "null worker");
    MLSB_CUDA(cudaStreamSynchronize(w->stream));
    collectStats(w);
    for (StatAccum &a : w->stats) a = StatAccum();
    for (uint64_t &v : w->marchWork) v = 0;
    MLSB_CUDA(cudaMemset(w->mlsWork, 0, 2 * sizeof(unsigned long long)));
    return MLSB_OK;
}

int mlsb_worker_set_work_counting(mlsb_worker *w, int enabled)
{
    MLSB_REQUIRE(w != nullptr, MLSB_ERR_INVALID_ARGUMENT, "null worker");
    w->countWork = enabled != 0;
    return MLSB_OK;
}

int mlsb_worker_get_mls_work(mlsb_worker *w, uint64_t *pairTests, uint64_t *hits)
{
    MLSB_REQUIRE(w != nullptr, MLSB_ERR_INVALID_ARGUMENT, "null worker");
    unsigned long long h[2] = {0, 0};
    MLSB_CUDA(cudaStreamSynchronize(w->stream));
    MLSB_CUDA(cudaMemcpy(h, w->mlsWork, sizeof(h), cudaMemcpyDeviceToHost));
    if (pairTests) *pairTests = h[0];
    if (hits) *hits = h[1];
    return MLSB_OK;
}

int mlsb_worker_get_marching_work(mlsb_worker *w, uint64_t *corners, uint64_t *cells, uint64_t *unwelded, uint64_t *indices,
                                  uint64_t *welded)
{
    MLSB_REQUIRE(w != nullptr, MLSB_ERR_INVALID_ARGUMENT, "null worker");
    if (corners) *corners = w->marchWork[0];
    if (cells) *cells = w->marchWork[1];
    if (unwelded) *unwelded = w->marchWork[2];
    if (indices) *indices = w->marchWork[3];
    if (welded) *welded = w->marchWork[4];
    return MLSB_OK;
}

int mlsb_measure_fp32_peak(int device, double *laneInstPerSecond)
{
    MLSB_REQUIRE(laneInstPerSecond != nullptr, MLSB_ERR_INVALID_ARGUMENT, "null pointer");
    MLSB_CUDA(cudaSetDevice(device));
    return measureFp32Peak(laneInstPerSecond, nullptr);
}

uint64_t mlsb_worker_launch_count(mlsb_worker *w) { return w ? w->launches : 0; }

int mlsb_device_alloc(void **ptr, size_t bytes)
{
    MLSB_REQUIRE(ptr != nullptr, MLSB_ERR_INVALID_ARGUMENT, "null pointer");
    MLSB_CUDA(cudaMalloc(ptr, bytes ? bytes : 1));
    return MLSB_OK;
}

int mlsb_device_free(void *ptr)
{
    MLSB_CUDA(cudaFree(ptr));
    return MLSB_OK;
}

int mlsb_memcpy_h2d(void *dst, const void *src, size_t bytes)
{
    MLSB_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
    return MLSB_OK;
}

int mlsb_memcpy_d2h(void *dst, const void *src, size_t bytes)
{
    MLSB_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return MLSB_OK;
}

int mlsb_memcpy_d2d_async(void *dst, const void *src, size_t bytes, void *stream)
{
    if (bytes == 0) return MLSB_OK;
    MLSB_REQUIRE(dst != nullptr && src != nullptr, MLSB_ERR_INVALID_ARGUMENT, "null pointer");
    MLSB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
    return MLSB_OK;
}

int mlsb_device_memset(void *dst, int value, size_t bytes)
{
    MLSB_CUDA(cudaMemset(dst, value, bytes));
    return MLSB_OK;
}

}  // extern "C"
