/*
 * mlsb200 — C ABI of the B200-native mlsgpu reconstruction core.
 *
 * This is the drop-in boundary for the per-bin device path of bmerry/mlsgpu:
 * octree build -> per-corner MLS -> marching tetrahedra + weld -> scale/bias.
 * Each entry point names the reference interface it replaces (paths relative to the
 * upstream tree).  Plain pointers and sizes only; no C++/torch types; no exceptions cross
 * this boundary — every call returns an mlsb_status and the message of the last failure on
 * the calling thread is available from mlsb_last_error_string().
 *
 * Threading: one mlsb_worker <-> one host thread <-> one GPU, exactly like the reference's
 * per-worker (SplatTreeCL, MlsFunctor, Marching, ScaleBiasFilter) set created in
 * DeviceWorkerGroupBase::Worker::Worker (src/workers.cpp:207-225).  A worker owns all of its
 * device scratch, allocated once at creation for the worst case and recycled per bin
 * (src/splat_tree_cl.cpp:114-129, src/marching.cpp:380-405).
 *
 * There is no CPU fallback: every compute entry point fails with MLSB_ERR_CUDA when no
 * CUDA device is usable.
 */
#ifndef MLSB200_H
#define MLSB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MLSB_API __attribute__((visibility("default")))

typedef enum
{
    MLSB_OK = 0,
    MLSB_ERR_INVALID_ARGUMENT = 1,  /* std::invalid_argument in the reference (MLSGPU_ASSERT) */
    MLSB_ERR_LENGTH = 2,            /* std::length_error: a size exceeds what the worker was built for */
    MLSB_ERR_STATE = 3,             /* state_error: call sequence violated */
    MLSB_ERR_CUDA = 4,              /* cl::Error equivalent: the CUDA runtime reported a failure */
    MLSB_ERR_OVERFLOW = 5           /* mesh does not fit in mesh_bytes even one slice at a time */
} mlsb_status;

/* struct Splat, src/splat.h:40-45 (device view kernels/mls.cl:63-67).  32 bytes. */
typedef struct
{
    float position[3];   /* grid units (full-grid vertex coordinates, src/bucket_loader.cpp:79-85) */
    float radius;        /* grid units; the octree build overwrites it with 1/radius^2 */
    float normal[3];
    float quality;       /* 1/(smooth*r_world)^2, src/fast_ply.cpp:346-348 */
} mlsb_splat;

/* MlsShape, src/mls.h:45-49 */
typedef enum { MLSB_SHAPE_SPHERE = 0, MLSB_SHAPE_PLANE = 1 } mlsb_shape;

/* Marching::Swathe / ImageParams, src/marching.h:167-198: corner (x,y,z) lives at
 * field[(y + z*zStride + zBias) * pitch + x]. */
typedef struct
{
    uint32_t width, height;   /* corners carrying payload */
    uint32_t zStride;
    int32_t zBias;
    uint32_t zFirst, zLast;   /* closed interval of corner slices */
} mlsb_swathe;

/* MeshSizes, src/mesh.h:40-81 */
typedef struct
{
    uint64_t numVertices;
    uint64_t numTriangles;
    uint64_t numInternalVertices;
} mlsb_mesh_sizes;

/* What DeviceWorkerGroup hands to each Worker: --levels, --mem-bucket-splats / sizeof(Splat),
 * maxCells (bin side), per-worker mesh memory, --fit-shape, --fit-boundary-limit
 * (src/workers.h:214-350, src/mlsgpu_core.cpp:704-741). */
typedef struct
{
    int32_t device;           /* CUDA ordinal (replaces --cl-device) */
    uint32_t maxLevels;       /* SplatTreeCL maxLevels, 1..10 (src/splat_tree_cl.h:76) */
    uint32_t subsampling;     /* --subsampling, >= 3 for the MLS kernel (src/mls.cpp:54) */
    uint32_t maxSplats;       /* SplatTreeCL maxSplats: capacity of the device splat buffer */
    uint32_t maxCells;        /* bin side in cells; Marching gets maxCells+1 corners per axis */
    uint64_t meshBytes;       /* Marching meshMemory; vertex/index space = meshBytes/872 cells x 13 / x 36 */
    int32_t shape;            /* mlsb_shape */
    float boundaryLimit;      /* --fit-boundary-limit */
} mlsb_worker_config;

typedef struct mlsb_worker mlsb_worker;

/* SplatTreeCL::MAX_LEVELS and MAX_SPLATS (src/splat_tree_cl.h:76,88), Marching::MAX_DIMENSION as built here
 * (the reference's is 8192, src/marching.h:103; bins above 512 corners per side are not supported) and
 * MAX_GLOBAL_DIMENSION (src/marching.h:120). */
#define MLSB_MAX_LEVELS 10u
#define MLSB_MAX_SPLATS ((1u << 31) / 16u)
#define MLSB_MAX_DIMENSION 512u
#define MLSB_MAX_GLOBAL_DIMENSION ((1u << 20) - 1u)

/* CLH::ResourceUsage (src/clh.h) for one worker: what DeviceWorkerGroup::resourceUsage adds up
 * (src/workers.cpp:176-205) plus the pinned host memory the reference accounts for separately. */
typedef struct
{
    uint64_t deviceBytes;          /* total device memory a worker allocates */
    uint64_t maxAllocationBytes;   /* its largest single allocation (CL_DEVICE_MAX_MEM_ALLOC_SIZE check) */
    uint64_t pinnedHostBytes;      /* page-locked host memory: mesh read-back blob and counters */
} mlsb_resource_usage;

/* Marching::Generator (src/marching.h:204-253) as a callback: must ENQUEUE, on `stream` (a
 * cudaStream_t), work that fills deviceField — corner (x, y, z) at
 * deviceField[(y + z*zStride + zBias) * pitch + x] — for the corner slices zFirst..zLast of the swathe,
 * with NaN wherever the function is undefined, and return 0 (anything else aborts the generate with
 * MLSB_ERR_STATE).  Rows and columns beyond width/height up to the next multiple of 8 may be
 * written; nothing else may be (the post-conditions of :233-246). */
typedef int (*mlsb_field_generator)(void *user, float *deviceField, uint32_t pitch, const mlsb_swathe *swathe, void *stream);

/* DeviceKeyMesh (src/mesh.h:101-123): one shipped-out block still on the device.  vertices holds
 * numVertices packed float[3] (internal first, then external), vertexKeys numVertices uint64 of which
 * only [numInternalVertices, numVertices) are defined, triangles numTriangles x uint32[3] indexing
 * vertices.  The buffers are the worker's own and stay valid until the callback returns. */
typedef struct
{
    mlsb_mesh_sizes sizes;
    float *vertices;
    uint64_t *vertexKeys;
    uint32_t *triangles;
} mlsb_device_mesh;

/* Marching::OutputFunctor (src/marching.h:516-546) on the device side: called once per block before
 * any read-back.  It may modify the mesh in place with work enqueued on `stream` (a MeshFilter,
 * src/mesh_filter.h:55-116) and/or read it back itself with mlsb_device_mesh_read; whatever it
 * enqueues must have been waited for, or be on `stream`, when it returns.  A non-zero return aborts. */
typedef int (*mlsb_device_mesh_sink)(const mlsb_device_mesh *mesh, void *stream, void *user);

/* MeshFilter (src/mesh_filter.h:55-67): reads *in and describes its result in *out, with any device work
 * ENQUEUED on `stream` (a cudaStream_t).  *out arrives as a copy of *in: a filter that works in place only
 * edits out->sizes; one that writes elsewhere points out's arrays at its own device buffers, which must
 * stay valid until the block has been handed over.  Sizes are host values (as in DeviceKeyMesh), so a filter
 * that changes them must know the new counts on the host.  A non-zero return aborts the generate call with
 * MLSB_ERR_STATE. */
typedef int (*mlsb_mesh_filter)(void *user, const mlsb_device_mesh *in, mlsb_device_mesh *out, void *stream);

/* Receives one shipped-out block, the equivalent of Marching::OutputFunctor after
 * enqueueReadMesh (src/marching.h:516-546, src/mesh.cpp:62-102).  hostBlob has the HostKeyMesh
 * layout (src/mesh.cpp:51-60): [uint64 keys of external vertices][float[3] vertices]
 * [uint32[3] triangles], 8-byte aligned, mlsb_mesh_host_bytes(sizes) bytes; it is owned by the
 * worker and valid only during the callback.  Called zero or more times per generate,
 * serialised, on the calling thread. */
typedef void (*mlsb_mesh_sink)(const mlsb_mesh_sizes *sizes, const void *hostBlob, void *user);

/* ---- library / device ------------------------------------------------------------------ */

/* CLH::findDevices (src/clh.cpp:180-240): number of usable CUDA devices. */
MLSB_API int mlsb_device_count(int *count);
MLSB_API const char *mlsb_last_error_string(void);
MLSB_API const char *mlsb_version(void);

/* ---- worker lifetime -------------------------------------------------------------------- */

/* DeviceWorkerGroupBase::Worker::Worker, src/workers.cpp:207-225. */
MLSB_API int mlsb_worker_create(const mlsb_worker_config *config, mlsb_worker **out);
MLSB_API void mlsb_worker_destroy(mlsb_worker *w);
/* DeviceWorkerGroup::resourceUsage (src/workers.cpp:176-205): device bytes a worker allocates. */
MLSB_API int mlsb_worker_resource_usage(const mlsb_worker_config *config, uint64_t *deviceBytes);
MLSB_API int mlsb_worker_resource_usage_ex(const mlsb_worker_config *config, mlsb_resource_usage *usage);
/* SplatTreeCL::validateDevice / Marching::validateDevice / MlsFunctor's requirements
 * (src/splat_tree_cl.cpp:55-66, src/marching.cpp:296-309, src/clh.cpp:180-240) for a CUDA device:
 * MLSB_OK if the library can run on it, otherwise MLSB_ERR_CUDA with the reason in
 * mlsb_last_error_string() (compute capability other than 10.x — the kernels are built for sm_100a
 * only —, too little shared memory, no thread-block clusters). */
MLSB_API int mlsb_validate_device(int device);
/* The worker's own stream (the cl::CommandQueue of workers.cpp:215). */
MLSB_API void *mlsb_worker_stream(mlsb_worker *w);
MLSB_API int mlsb_worker_synchronize(mlsb_worker *w);

/* ---- splat staging (CopyGroup, src/workers.cpp:354-373) --------------------------------- */

/* Pinned host staging memory (CLH::PinnedMemory, src/clh.h). */
MLSB_API int mlsb_host_alloc(void **ptr, size_t bytes);
MLSB_API int mlsb_host_free(void *ptr);
/* Asynchronous H2D of n splats into the worker's splat buffer at element dstFirst, on the
 * worker's copy stream; the following octree build waits for it (copyEvent, workers.cpp:268). */
MLSB_API int mlsb_splats_upload_async(mlsb_worker *w, const mlsb_splat *host, uint32_t dstFirst, uint32_t n);
/* Device address of the worker's splat buffer (SplatTreeCL::getSplats). */
MLSB_API int mlsb_splats_device(mlsb_worker *w, mlsb_splat **devicePtr);

/* ---- octree: SplatTreeCL (src/splat_tree_cl.h:242-296) ---------------------------------- */

/* SplatTreeCL::enqueueBuild (src/splat_tree_cl.cpp:269-335) on the worker's stream.
 * size = corners of the region rounded up to 8; offset = bin low extent in the full grid. */
MLSB_API int mlsb_octree_build(mlsb_worker *w, uint32_t firstSplat, uint32_t numSplats,
                               const uint32_t size[3], const int32_t offset[3],
                               uint32_t subsamplingShift);
/* getStart/getCommands/getNumLevels (device pointers; test hook). numCommands = capacity. */
MLSB_API int mlsb_octree_get(mlsb_worker *w, const int32_t **start, uint32_t *numStart,
                             const int32_t **commands, uint32_t *numCommands, uint32_t *numLevels);

/* ---- MLS: MlsFunctor (src/mls.h:79-170) -------------------------------------------------- */

/* MlsFunctor::set(offset, tree, subsamplingShift), src/mls.cpp:90-94. */
MLSB_API int mlsb_mls_set(mlsb_worker *w, const int32_t offset[3], uint32_t subsamplingShift);
/* MlsFunctor::set(offset, splats, commands, start, shift), src/mls.cpp:76-88: explicit device
 * buffers (the hand-built octree of test/test_mls.cpp:416-514). */
MLSB_API int mlsb_mls_set_buffers(mlsb_worker *w, const int32_t offset[3], const mlsb_splat *splats,
                                  const int32_t *commands, const int32_t *start,
                                  uint32_t subsamplingShift);
/* MlsFunctor::setBoundaryLimit, src/mls.cpp:137-144. */
MLSB_API int mlsb_mls_set_boundary_limit(mlsb_worker *w, float limit);
/* MlsFunctor::enqueue -> processCorners (src/mls.cpp:101-135, kernels/mls.cl:299-433): fills a
 * device float image of row pitch `pitch` floats for the swathe, asynchronously. */
MLSB_API int mlsb_mls_corners(mlsb_worker *w, float *deviceField, uint32_t pitch, const mlsb_swathe *swathe);

/* ---- marching: Marching + ScaleBiasFilter (src/marching.h:564-608, src/mesh_filter.h) ---- */

/* ScaleBiasFilter::setScaleBias(scale, x, y, z), src/mesh_filter.cpp:77-84. */
MLSB_API int mlsb_set_scale_bias(mlsb_worker *w, float scale, const float bias[3]);
/* Marching::generate with the MlsFunctor bound by mlsb_mls_set as generator
 * (src/marching.cpp:745-824).  size = corners per axis of the bin, keyOffset = its low extent.
 * BLOCKING.  Calls sink once per shipped-out block. */
MLSB_API int mlsb_marching_generate(mlsb_worker *w, const uint32_t size[3], const uint32_t keyOffset[3],
                                    mlsb_mesh_sink sink, void *user);
/* Same with a host-supplied field ([depth][height][width] floats), the HostGenerator of
 * test/test_marching.cpp:62-129: marching-only tests. */
MLSB_API int mlsb_marching_generate_host_field(mlsb_worker *w, const float *hostField,
                                               const uint32_t size[3], const uint32_t keyOffset[3],
                                               mlsb_mesh_sink sink, void *user);
/* Marching::generate with ANY generator (src/marching.h:603-608): `generate` fills the worker's
 * field, once for the whole bin (the field is resident; there are no swathes to loop over). */
MLSB_API int mlsb_marching_generate_cb(mlsb_worker *w, mlsb_field_generator generate, void *generatorUser,
                                       const uint32_t size[3], const uint32_t keyOffset[3],
                                       mlsb_mesh_sink sink, void *user);
/* Marching::generate handing out DeviceKeyMesh blocks (the reference's OutputFunctor contract)
 * instead of host blobs; generate == NULL uses the MlsFunctor bound by mlsb_mls_set. */
MLSB_API int mlsb_marching_generate_device(mlsb_worker *w, mlsb_field_generator generate, void *generatorUser,
                                           const uint32_t size[3], const uint32_t keyOffset[3],
                                           mlsb_device_mesh_sink sink, void *user);
/* MeshFilterChain::addFilter (src/mesh_filter.h:96-99) for this worker: every block a generate or process_bin
 * call ships passes through the filters, in the order added, between the marching output (already scaled and
 * biased: the reference's ScaleBiasFilter is the first link of its chain, src/workers.cpp:224) and the sink —
 * the host sink then receives the FILTERED mesh, read back by enqueueReadMesh.  Blocks a filter empties are
 * still handed over, as MeshFilterChain does (test/test_mesh_filter.cpp:233-249). */
MLSB_API int mlsb_worker_add_mesh_filter(mlsb_worker *w, mlsb_mesh_filter filter, void *user);
MLSB_API int mlsb_worker_clear_mesh_filters(mlsb_worker *w);
/* MeshFilterChain::operator() on a mesh of the caller's (src/mesh_filter.cpp:44-64): runs the worker's filters on
 * `mesh` and hands the result to `sink` (then waits for the stream). */
MLSB_API int mlsb_mesh_filter_chain_apply(mlsb_worker *w, const mlsb_device_mesh *mesh, mlsb_device_mesh_sink sink, void *user);
/* enqueueReadMesh (src/mesh.cpp:62-102): reads a device mesh into a HostKeyMesh blob of
 * mlsb_mesh_host_bytes(&mesh->sizes) bytes (pinned memory for speed) and waits for it. */
MLSB_API int mlsb_device_mesh_read(mlsb_worker *w, const mlsb_device_mesh *mesh, void *hostBlob);
/* Bytes of the HostKeyMesh blob, MeshSizes::getHostBytes (src/mesh.h:76-81). */
MLSB_API uint64_t mlsb_mesh_host_bytes(const mlsb_mesh_sizes *sizes);
/* The four lookup tables of Marching::makeTables (src/marching.cpp:109-252), host copies:
 * count 256x2 u8, start 257x2 u16, data 8192 u8, key 2432x3 u32. */
MLSB_API int mlsb_marching_tables(uint8_t *count, uint16_t *start, uint8_t *data, uint32_t *key);

/* ---- one bin end to end: DeviceWorkerGroupBase::Worker::operator() ---------------------- */

/* src/workers.cpp:232-286 for one SubItem: octree build on splats [firstSplat, firstSplat+n)
 * of the worker's buffer (previously uploaded), MlsFunctor::set, Marching::generate.
 * size = bin corners per axis (numVertices), offset = bin low extent (also the key offset). */
MLSB_API int mlsb_worker_process_bin(mlsb_worker *w, uint32_t firstSplat, uint32_t numSplats,
                                     const uint32_t size[3], const int32_t offset[3],
                                     mlsb_mesh_sink sink, void *user);

/* ---- worker group: CopyGroup + DeviceWorkerGroup (src/workers.h:136-438) ------------------ */

typedef struct mlsb_group mlsb_group;

/* Like mlsb_mesh_sink, plus the tag given at submission (the ChunkId / bin identity that
 * OutputGeneratorBuilder carries, src/workers.h:488-509).  Calls are serialised across workers
 * but come from the group's worker threads. */
typedef void (*mlsb_group_sink)(uint64_t tag, const mlsb_mesh_sizes *sizes, const void *hostBlob, void *user);

/* numWorkers worker threads on config->device, each with its own mlsb_worker (the reference's
 * --device-threads), fed by a bounded queue of queueDepth bins (0 = 2 x numWorkers).
 * stageSplats != 0: every worker copies the bin's splats into its own pinned staging buffer before
 * the asynchronous H2D, so the caller's memory (pageable or pinned) is free again when
 * mlsb_group_wait returns or the queue slot is consumed — CopyGroupBase::Worker::operator(),
 * src/workers.cpp:377-418.  stageSplats == 0: the pointer given to submit is copied directly
 * (pinned host or device memory) and must stay valid until the bin has been processed. */
MLSB_API int mlsb_group_create(const mlsb_worker_config *config, uint32_t numWorkers, uint32_t queueDepth,
                               int stageSplats, mlsb_group_sink sink, void *user, mlsb_group **out);
MLSB_API void mlsb_group_destroy(mlsb_group *g);
/* Queues one bin (WorkerGroup::get + push, src/worker_group.h:137-199); blocks while the queue is
 * full.  size = corners per axis, offset = low extent of the bin in the full grid.
 * The splat memory must stay valid until the bin is released, see mlsb_group_released. */
MLSB_API int mlsb_group_submit(mlsb_group *g, const mlsb_splat *splats, uint32_t numSplats,
                               const uint32_t size[3], const int32_t offset[3], uint64_t tag);
/* count = the number of leading submits (in submission order) whose splat memory the group will
 * not read again: with stageSplats a bin is released as soon as a worker has copied it into its
 * staging buffer, otherwise when its upload has completed.  A loader that recycles gather buffers
 * (BucketLoader's circular buffer, src/bucket_loader.cpp:104-126) polls this before reusing one. */
MLSB_API int mlsb_group_released(mlsb_group *g, uint64_t *count);
/* Blocks until every queued bin has been processed; returns the first error any worker hit. */
MLSB_API int mlsb_group_wait(mlsb_group *g);
/* By default the sink is called by one worker at a time, like the reference's single mesher thread
 * (MesherGroup, src/workers.h:511-560).  enabled != 0: workers call it concurrently and the
 * consumer does its own locking (set before the first submit). */
MLSB_API int mlsb_group_set_concurrent_sink(mlsb_group *g, int enabled);
/* Worker i of the group (to set scale/bias, boundary limit, profiling; not while bins are queued). */
MLSB_API int mlsb_group_worker(mlsb_group *g, uint32_t index, mlsb_worker **worker);
/* Totals since creation: bins and splats processed, bytes of mesh handed to the sink. */
MLSB_API int mlsb_group_counters(mlsb_group *g, uint64_t *bins, uint64_t *splats, uint64_t *meshBytes);

/* ---- timeplot (--timeplot, src/timeplot.cpp:134-186) -------------------------------------- */

/* Seconds on the library's monotonic clock (the Timer::currentTime() the events below are stamped with). */
MLSB_API double mlsb_timeplot_now(void);
/* enabled != 0: every worker thread of the group records what it does per bin as the reference's
 * Timeplot::Action does — worker "device.<ordinal>.<index>", actions "get" (waiting for a bin),
 * "copy" (staging + H2D enqueue, CopyGroup's "write") and "compute" (the per-bin path including the
 * mesh sink).  Set before the first submit. */
MLSB_API int mlsb_group_set_timeplot(mlsb_group *g, int enabled);
/* Calls visit once per recorded event, in no particular order (call after mlsb_group_wait). */
typedef void (*mlsb_timeplot_visitor)(void *user, const char *worker, const char *action, double t0, double t1);
MLSB_API int mlsb_group_timeplot(mlsb_group *g, mlsb_timeplot_visitor visit, void *user);

/* ---- statistics (--statistics-cl, src/statistics_cl.cpp) -------------------------------- */

/* CUDA-event timing under the reference's statistic names (kernel.octree.*.time,
 * kernel.mls.processCorners.time, kernel.marching.*.time): level 0 off, 1 processCorners
 * only, 2 every kernel. */
MLSB_API int mlsb_worker_set_profiling(mlsb_worker *w, int level);
/* Sum (seconds) and count of a named statistic since the last reset. */
MLSB_API int mlsb_worker_get_statistic(mlsb_worker *w, const char *name, double *sum, uint64_t *count);
/* Every statistic of the worker that has samples, in name order: sum, sum of squares and count,
 * which is what Statistics::Variable holds (src/statistics.h:164-209) — the caller adds them to
 * its registry.  Besides the kernel times: "marching.shipouts" (blocks shipped per generate call)
 * and "marching.overflow" (blocks that did not fit the mesh memory and were split),
 * src/marching.cpp:350-352. */
typedef void (*mlsb_statistic_visitor)(void *user, const char *name, double sum, double sum2, uint64_t count);
MLSB_API int mlsb_worker_visit_statistics(mlsb_worker *w, mlsb_statistic_visitor visit, void *user);
MLSB_API int mlsb_worker_reset_statistics(mlsb_worker *w);
/* Work model of SURVEY 8(d): while counting is enabled every MLS launch also accumulates its
 * pair tests T and hits H on the device (slower kernel variant: enable only for an untimed
 * pass).  Counters are cleared by mlsb_worker_reset_statistics. */
MLSB_API int mlsb_worker_set_work_counting(mlsb_worker *w, int enabled);
MLSB_API int mlsb_worker_get_mls_work(mlsb_worker *w, uint64_t *pairTests, uint64_t *hits);
/* Marching's side of the same model, always counted (the per-block sizes are read back anyway):
 * corners classified (C), occupied cells (K), per-cell i.e. unwelded vertices (V), indices (I) and
 * welded vertices of all blocks shipped since the last mlsb_worker_reset_statistics. */
MLSB_API int mlsb_worker_get_marching_work(mlsb_worker *w, uint64_t *corners, uint64_t *cells, uint64_t *unwelded,
                                           uint64_t *indices, uint64_t *welded);
/* FP32 pipe peak of a device in lane-instructions per second, measured with a dependent-free
 * FFMA stream (the roofline denominator of processCorners). */
MLSB_API int mlsb_measure_fp32_peak(int device, double *laneInstPerSecond);
/* Total kernels launched by this worker since creation. */
MLSB_API uint64_t mlsb_worker_launch_count(mlsb_worker *w);

/* ---- raw device memory helpers for tests and bindings ----------------------------------- */
MLSB_API int mlsb_device_alloc(void **ptr, size_t bytes);
MLSB_API int mlsb_device_free(void *ptr);
MLSB_API int mlsb_memcpy_h2d(void *dst, const void *src, size_t bytes);
MLSB_API int mlsb_memcpy_d2h(void *dst, const void *src, size_t bytes);
MLSB_API int mlsb_device_memset(void *dst, int value, size_t bytes);
/* Device-to-device copy enqueued on `stream` (a cudaStream_t; what a MeshFilter or a device mesh sink uses,
 * CLH::enqueueCopyBuffer in test/test_mesh_filter.cpp:84-103). */
MLSB_API int mlsb_memcpy_d2d_async(void *dst, const void *src, size_t bytes, void *stream);

#ifdef __cplusplus
}
#endif

#endif /* MLSB200_H */
