// Device worker group: the reference's CopyGroup + DeviceWorkerGroup (src/workers.h:136-438,
// src/workers.cpp:87-418) as one native thread pool.  Every worker thread owns an mlsb_worker
// (its own CUDA streams and scratch), takes bins from a bounded queue, stages the splats through
// its own pinned buffer (memcpy + asynchronous H2D on the copy stream) and runs the per-bin path;
// while one worker waits for a mesh readback the others keep the GPU busy, which is what the
// reference gets from --device-threads and its out-of-order copy queue.
#include "common.cuh"

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

struct mlsb_group
{
    struct Item
    {
        const mlsb_splat *splats;
        uint32_t numSplats;
        uint32_t size[3];
        int32_t offset[3];
        uint64_t tag;
        uint64_t sequence;   // position in submission order
    };

    std::vector<mlsb_worker *> workers;
    std::vector<mlsb_splat *> staging;   // pinned, one per worker (CLH::PinnedMemory in CopyGroup)
    std::vector<std::thread> threads;
    std::deque<Item> queue;
    std::mutex mutex;
    std::condition_variable canPush, canPop, idle;
    size_t maxQueue = 0;
    size_t inFlight = 0;
    uint64_t submitted = 0;          // bins submitted so far
    uint64_t releasedUpTo = 0;       // the splat memory of every bin with sequence < this is no longer read
    std::vector<uint64_t> releasedLater;   // released out of order (workers overtake each other)
    bool stopping = false;
    bool stage = true;

    mlsb_group_sink sink = nullptr;
    void *user = nullptr;
    std::mutex sinkMutex;   // meshes are delivered serialised, like the single mesher thread ...
    bool concurrentSink = false;   // ... unless the consumer does its own locking

    std::mutex errorMutex;
    int status = MLSB_OK;
    std::string message;

    std::atomic<uint64_t> bins{0}, splats{0}, meshBytes{0};

    // --timeplot: one record per action, appended by the worker threads
    struct Event
    {
        uint32_t worker;
        const char *action;
        double t0, t1;
    };
    bool timeplot = false;
    std::mutex eventMutex;
    std::vector<Event> events;
};

namespace
{

struct SinkContext
{
    mlsb_group *group;
    uint64_t tag;
};

void groupTrampoline(const mlsb_mesh_sizes *sizes, const void *blob, void *user)
{
    SinkContext *ctx = static_cast<SinkContext *>(user);
    mlsb_group *g = ctx->group;
    g->meshBytes += mlsb_mesh_host_bytes(sizes);
    if (!g->sink) return;
    // A C++ consumer may throw (bad_alloc in a mesher, say): nothing may unwind through the C ABI and the
    // worker thread, so the failure is recorded and reported by mlsb_group_wait like any other.
    try
    {
        if (g->concurrentSink)
            g->sink(ctx->tag, sizes, blob, g->user);
        else
        {
            std::lock_guard<std::mutex> lock(g->sinkMutex);
            g->sink(ctx->tag, sizes, blob, g->user);
        }
    }
    catch (const std::exception &e)
    {
        std::lock_guard<std::mutex> lock(g->errorMutex);
        if (g->status == MLSB_OK)
        {
            g->status = MLSB_ERR_STATE;
            g->message = std::string("mesh sink failed: ") + e.what();
        }
    }
    catch (...)
    {
        std::lock_guard<std::mutex> lock(g->errorMutex);
        if (g->status == MLSB_OK)
        {
            g->status = MLSB_ERR_STATE;
            g->message = "mesh sink failed with an unknown exception";
        }
    }
}

// The caller's splat memory of bin `sequence` will not be read again.
void releaseSplats(mlsb_group *g, uint64_t sequence)
{
    std::lock_guard<std::mutex> lock(g->mutex);
    if (sequence != g->releasedUpTo)
    {
        g->releasedLater.push_back(sequence);
        return;
    }
    g->releasedUpTo++;
    for (bool found = true; found;)
    {
        found = false;
        for (size_t i = 0; i < g->releasedLater.size(); i++)
            if (g->releasedLater[i] == g->releasedUpTo)
            {
                g->releasedLater[i] = g->releasedLater.back();
                g->releasedLater.pop_back();
                g->releasedUpTo++;
                found = true;
                break;
            }
    }
}

double clockNow()
{
    static const std::chrono::steady_clock::time_point epoch = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - epoch).count();
}

void record(mlsb_group *g, size_t index, const char *action, double t0, double t1)
{
    if (!g->timeplot) return;
    std::lock_guard<std::mutex> lock(g->eventMutex);
    g->events.push_back(mlsb_group::Event{(uint32_t) index, action, t0, t1});
}

void workerLoop(mlsb_group *g, size_t index)
{
    mlsb_worker *w = g->workers[index];
    for (;;)
    {
        mlsb_group::Item item;
        const double tGet = g->timeplot ? clockNow() : 0.0;
        {
            std::unique_lock<std::mutex> lock(g->mutex);
            g->canPop.wait(lock, [g] { return g->stopping || !g->queue.empty(); });
            if (g->stopping || g->queue.empty()) return;   // on destroy, bins still queued are dropped, not processed
            item = g->queue.front();
            g->queue.pop_front();
            g->inFlight++;
        }
        g->canPush.notify_one();
        const double tCopy = g->timeplot ? clockNow() : 0.0;
        record(g, index, "get", tGet, tCopy);

        // CopyGroupBase::Worker::operator() + flush, src/workers.cpp:354-418
        const mlsb_splat *src = item.splats;
        if (g->stage && item.numSplats)
        {
            memcpy(g->staging[index], item.splats, (size_t) item.numSplats * sizeof(mlsb_splat));
            src = g->staging[index];
        }
        if (g->stage) releaseSplats(g, item.sequence);
        int st = mlsb_splats_upload_async(w, src, 0, item.numSplats);
        const double tCompute = g->timeplot ? clockNow() : 0.0;
        record(g, index, "copy", tCopy, tCompute);
        // DeviceWorkerGroupBase::Worker::operator(), src/workers.cpp:232-286
        SinkContext ctx{g, item.tag};
        if (st == MLSB_OK) st = mlsb_worker_process_bin(w, 0, item.numSplats, item.size, item.offset, groupTrampoline, &ctx);
        if (st != MLSB_OK)
        {
            // the bin may have failed before the compute stream waited for the upload (argument checks):
            // the staging buffer must not be overwritten while that copy is still in flight
            mlsb_worker_synchronize(w);
            std::lock_guard<std::mutex> lock(g->errorMutex);
            if (g->status == MLSB_OK)
            {
                g->status = st;
                g->message = mlsb_last_error_string();
            }
        }
        if (g->timeplot) record(g, index, "compute", tCompute, clockNow());
        if (!g->stage) releaseSplats(g, item.sequence);   // process_bin has waited for the upload
        g->bins++;
        g->splats += item.numSplats;
        {
            std::lock_guard<std::mutex> lock(g->mutex);
            g->inFlight--;
        }
        g->idle.notify_all();
    }
}

}  // namespace

using namespace mlsb;

extern "C" {

int mlsb_group_create(const mlsb_worker_config *config, uint32_t numWorkers, uint32_t queueDepth, int stageSplats,
                      mlsb_group_sink sink, void *user, mlsb_group **out)
{
    MLSB_REQUIRE(out != nullptr && config != nullptr, MLSB_ERR_INVALID_ARGUMENT, "null pointer");
    MLSB_REQUIRE(numWorkers >= 1 && numWorkers <= 64, MLSB_ERR_INVALID_ARGUMENT, "numWorkers must be in [1, 64]");
    *out = nullptr;
    mlsb_group *g = new mlsb_group();
    g->sink = sink;
    g->user = user;
    g->stage = stageSplats != 0;
    g->maxQueue = queueDepth ? queueDepth : 2 * numWorkers;
    for (uint32_t i = 0; i < numWorkers; i++)
    {
        mlsb_worker *w = nullptr;
        int st = mlsb_worker_create(config, &w);
        mlsb_splat *pinned = nullptr;
        if (st == MLSB_OK && g->stage && cudaMallocHost(&pinned, (size_t) config->maxSplats * sizeof(mlsb_splat)) != cudaSuccess)
        {
            setError("cudaMallocHost(staging) failed: %s", cudaGetErrorString(cudaGetLastError()));
            st = MLSB_ERR_CUDA;
        }
        if (st != MLSB_OK)
        {
            if (w) mlsb_worker_destroy(w);
            mlsb_group_destroy(g);
            return st;
        }
        g->workers.push_back(w);
        g->staging.push_back(pinned);
    }
    for (uint32_t i = 0; i < numWorkers; i++) g->threads.emplace_back(workerLoop, g, (size_t) i);
    *out = g;
    return MLSB_OK;
}

void mlsb_group_destroy(mlsb_group *g)
{
    if (!g) return;
    {
        std::lock_guard<std::mutex> lock(g->mutex);
        g->stopping = true;
    }
    g->canPop.notify_all();
    for (std::thread &t : g->threads) t.join();
    g->queue.clear();
    for (mlsb_worker *w : g->workers) mlsb_worker_destroy(w);
    for (mlsb_splat *p : g->staging)
        if (p) cudaFreeHost(p);
    delete g;
}

int mlsb_group_submit(mlsb_group *g, const mlsb_splat *splats, uint32_t numSplats, const uint32_t size[3],
                      const int32_t offset[3], uint64_t tag)
{
    MLSB_REQUIRE(g != nullptr && size != nullptr && offset != nullptr && (splats != nullptr || numSplats == 0),
                 MLSB_ERR_INVALID_ARGUMENT, "null pointer");
    MLSB_REQUIRE(numSplats <= g->workers[0]->cfg.maxSplats, MLSB_ERR_LENGTH, "bin holds more splats than a worker's buffer");
    mlsb_group::Item item;
    item.splats = splats;
    item.numSplats = numSplats;
    memcpy(item.size, size, sizeof(item.size));
    memcpy(item.offset, offset, sizeof(item.offset));
    item.tag = tag;
    {
        std::unique_lock<std::mutex> lock(g->mutex);
        g->canPush.wait(lock, [g] { return g->queue.size() < g->maxQueue; });   // WorkerGroup::get blocks likewise
        item.sequence = g->submitted++;
        g->queue.push_back(item);
    }
    g->canPop.notify_one();
    return MLSB_OK;
}

int mlsb_group_wait(mlsb_group *g)
{
    MLSB_REQUIRE(g != nullptr, MLSB_ERR_INVALID_ARGUMENT, "null group");
    {
        std::unique_lock<std::mutex> lock(g->mutex);
        g->idle.wait(lock, [g] { return g->queue.empty() && g->inFlight == 0; });
    }
    std::lock_guard<std::mutex> lock(g->errorMutex);
    int st = g->status;
    if (st != MLSB_OK) setError("%s", g->message.c_str());
    g->status = MLSB_OK;
    return st;
}

int mlsb_group_released(mlsb_group *g, uint64_t *count)
{
    MLSB_REQUIRE(g != nullptr && count != nullptr, MLSB_ERR_INVALID_ARGUMENT, "null pointer");
    std::lock_guard<std::mutex> lock(g->mutex);
    *count = g->releasedUpTo;
    return MLSB_OK;
}

int mlsb_group_set_concurrent_sink(mlsb_group *g, int enabled)
{
    MLSB_REQUIRE(g != nullptr, MLSB_ERR_INVALID_ARGUMENT, "null group");
    g->concurrentSink = enabled != 0;
    return MLSB_OK;
}

double mlsb_timeplot_now(void) { return clockNow(); }

int mlsb_group_set_timeplot(mlsb_group *g, int enabled)
{
    MLSB_REQUIRE(g != nullptr, MLSB_ERR_INVALID_ARGUMENT, "null group");
    clockNow();   // fixes the epoch
    g->timeplot = enabled != 0;
    return MLSB_OK;
}

int mlsb_group_timeplot(mlsb_group *g, mlsb_timeplot_visitor visit, void *user)
{
    MLSB_REQUIRE(g != nullptr && visit != nullptr, MLSB_ERR_INVALID_ARGUMENT, "null pointer");
    std::lock_guard<std::mutex> lock(g->eventMutex);
    char name[64];
    for (const mlsb_group::Event &e : g->events)
    {
        snprintf(name, sizeof(name), "device.%d.%u", g->workers[0]->cfg.device, e.worker);
        visit(user, name, e.action, e.t0, e.t1);
    }
    return MLSB_OK;
}

int mlsb_group_worker(mlsb_group *g, uint32_t index, mlsb_worker **worker)
{
    MLSB_REQUIRE(g != nullptr && worker != nullptr && index < g->workers.size(), MLSB_ERR_INVALID_ARGUMENT, "bad worker index");
    *worker = g->workers[index];
    return MLSB_OK;
}

int mlsb_group_counters(mlsb_group *g, uint64_t *bins, uint64_t *splats, uint64_t *meshBytes)
{
    MLSB_REQUIRE(g != nullptr, MLSB_ERR_INVALID_ARGUMENT, "null group");
    if (bins) *bins = g->bins;
    if (splats) *splats = g->splats;
    if (meshBytes) *meshBytes = g->meshBytes;
    return MLSB_OK;
}

}  // extern "C"
