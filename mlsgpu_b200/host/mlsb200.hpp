// C++ mirror of the reference's device-path classes over the mlsb200 C ABI.
//
// The reference's host is C++ (Boost + OpenCL C++ bindings).  A maintainer switching the per-bin
// device path to this library keeps the call sites of src/workers.cpp:207-286 and replaces the
// four OpenCL-backed classes with the ones below, which have the same names, the same methods
// and the same argument meaning; cl::CommandQueue / cl::Event disappear because a worker owns
// its CUDA streams and every enqueue is ordered on them, and errors surface as the same
// exception types the reference throws (std::length_error, std::invalid_argument, state_error,
// and mlsb::cuda_error where the reference throws cl::Error).
//
//   reference (file:line)                                   here
//   SplatTreeCL            src/splat_tree_cl.h:242-296      mlsb::SplatTreeCL
//   MlsFunctor             src/mls.h:79-170                 mlsb::MlsFunctor
//   Marching               src/marching.h:564-608           mlsb::Marching
//   ScaleBiasFilter        src/mesh_filter.h:118-168        mlsb::ScaleBiasFilter
//   MeshSizes / HostKeyMesh src/mesh.h:40-157               mlsb::MeshSizes / mlsb::HostKeyMesh
//   DeviceWorkerGroupBase::Worker  src/workers.cpp:207-286  mlsb::Worker (owns the handle)
//
// Header-only; needs only include/mlsb200.h and the C++11 standard library.
#ifndef MLSB200_HPP
#define MLSB200_HPP

#include <array>
#include <cstdint>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/mlsb200.h"

namespace mlsb
{

/// Thrown where the reference would throw cl::Error.
class cuda_error : public std::runtime_error
{
public:
    explicit cuda_error(const std::string &msg) : std::runtime_error(msg) {}
};

/// state_error of src/errors.h:40-47.
class state_error : public std::logic_error
{
public:
    explicit state_error(const std::string &msg) : std::logic_error(msg) {}
};

inline void check(int status)
{
    if (status == MLSB_OK) return;
    const std::string msg = mlsb_last_error_string();
    switch (status)
    {
    case MLSB_ERR_INVALID_ARGUMENT: throw std::invalid_argument(msg);
    case MLSB_ERR_LENGTH: throw std::length_error(msg);
    case MLSB_ERR_STATE: throw state_error(msg);
    case MLSB_ERR_OVERFLOW: throw std::length_error(msg);
    default: throw cuda_error(msg);
    }
}

enum MlsShape { MLS_SHAPE_SPHERE = MLSB_SHAPE_SPHERE, MLS_SHAPE_PLANE = MLSB_SHAPE_PLANE };   // src/mls.h:45-49

/// struct Splat, src/splat.h:40-45
typedef mlsb_splat Splat;

/// MeshSizes, src/mesh.h:40-81
class MeshSizes
{
public:
    MeshSizes() : sizes{0, 0, 0} {}
    explicit MeshSizes(const mlsb_mesh_sizes &s) : sizes(s) {}
    std::size_t numVertices() const { return sizes.numVertices; }
    std::size_t numTriangles() const { return sizes.numTriangles; }
    std::size_t numInternalVertices() const { return sizes.numInternalVertices; }
    std::size_t numExternalVertices() const { return sizes.numVertices - sizes.numInternalVertices; }
    std::size_t getHostBytes() const { return mlsb_mesh_host_bytes(&sizes); }

private:
    mlsb_mesh_sizes sizes;
};

/// HostKeyMesh, src/mesh.h:125-157: views into one blob laid out [keys][vertices][triangles].
struct HostKeyMesh : public MeshSizes
{
    const std::array<float, 3> *vertices;
    const std::array<std::uint32_t, 3> *triangles;
    const std::uint64_t *vertexKeys;   ///< keys of the EXTERNAL vertices only

    HostKeyMesh(const void *ptr, const mlsb_mesh_sizes &s) : MeshSizes(s)
    {
        vertexKeys = static_cast<const std::uint64_t *>(ptr);
        vertices = reinterpret_cast<const std::array<float, 3> *>(vertexKeys + numExternalVertices());
        triangles = reinterpret_cast<const std::array<std::uint32_t, 3> *>(vertices + numVertices());
    }
};

/// One worker = one host thread = one GPU: the object set that
/// DeviceWorkerGroupBase::Worker::Worker (src/workers.cpp:207-225) constructs.
class Worker
{
public:
    Worker(int device, unsigned maxLevels, unsigned subsampling, std::size_t maxSplats, unsigned maxCells,
           std::size_t meshMemory, MlsShape shape, float boundaryLimit)
    {
        mlsb_worker_config c;
        c.device = device;
        c.maxLevels = maxLevels;
        c.subsampling = subsampling;
        c.maxSplats = static_cast<std::uint32_t>(maxSplats);
        c.maxCells = maxCells;
        c.meshBytes = meshMemory;
        c.shape = shape;
        c.boundaryLimit = boundaryLimit;
        check(mlsb_worker_create(&c, &handle));
    }
    ~Worker() { mlsb_worker_destroy(handle); }
    Worker(const Worker &) = delete;
    Worker &operator=(const Worker &) = delete;

    mlsb_worker *get() const { return handle; }
    /// CopyGroupBase::Worker::flush, src/workers.cpp:354-373: async H2D from pinned staging.
    void uploadSplats(const Splat *host, std::size_t dstFirst, std::size_t n)
    {
        check(mlsb_splats_upload_async(handle, host, static_cast<std::uint32_t>(dstFirst), static_cast<std::uint32_t>(n)));
    }
    void finish() { check(mlsb_worker_synchronize(handle)); }

private:
    mlsb_worker *handle = nullptr;
};

/// SplatTreeCL, src/splat_tree_cl.h:242-296.
class SplatTreeCL
{
public:
    enum { MAX_LEVELS = 10 };                    // src/splat_tree_cl.h:76
    explicit SplatTreeCL(Worker &worker) : worker(worker) {}

    /// enqueueBuild, src/splat_tree_cl.cpp:269-335 (asynchronous on the worker's stream).
    void enqueueBuild(std::size_t firstSplat, std::size_t numSplats, const std::uint32_t size[3],
                      const std::int32_t offset[3], unsigned subsamplingShift)
    {
        check(mlsb_octree_build(worker.get(), static_cast<std::uint32_t>(firstSplat), static_cast<std::uint32_t>(numSplats),
                                size, offset, subsamplingShift));
    }
    const std::int32_t *getStart() const { return get().start; }        ///< device pointer
    const std::int32_t *getCommands() const { return get().commands; }  ///< device pointer
    const Splat *getSplats() const
    {
        mlsb_splat *p = nullptr;
        check(mlsb_splats_device(worker.get(), &p));
        return p;
    }
    std::size_t getNumLevels() const { return get().levels; }
    void clearSplats() {}   // the worker owns its splat buffer: nothing to release

private:
    struct Buffers { const std::int32_t *start, *commands; std::uint32_t levels; };
    Buffers get() const
    {
        Buffers b{nullptr, nullptr, 0};
        check(mlsb_octree_get(worker.get(), &b.start, nullptr, &b.commands, nullptr, &b.levels));
        return b;
    }
    Worker &worker;
    friend class MlsFunctor;
};

/// Marching::Swathe, src/marching.h:181-198
typedef mlsb_swathe Swathe;

/// MlsFunctor, src/mls.h:79-170 (a Marching::Generator).
class MlsFunctor
{
public:
    explicit MlsFunctor(Worker &worker) : worker(worker) {}
    /// set(offset, tree, subsamplingShift), src/mls.cpp:90-94
    void set(const std::int32_t offset[3], const SplatTreeCL &, unsigned subsamplingShift)
    {
        check(mlsb_mls_set(worker.get(), offset, subsamplingShift));
    }
    /// set(offset, splats, commands, start, subsamplingShift), src/mls.cpp:76-88 (device pointers)
    void set(const std::int32_t offset[3], const Splat *splats, const std::int32_t *commands, const std::int32_t *start,
             unsigned subsamplingShift)
    {
        check(mlsb_mls_set_buffers(worker.get(), offset, splats, commands, start, subsamplingShift));
    }
    void setBoundaryLimit(float limit) { check(mlsb_mls_set_boundary_limit(worker.get(), limit)); }
    const std::uint32_t *alignment() const
    {
        static const std::uint32_t wgs[3] = {8, 8, 8};   // src/mls.cpp:53
        return wgs;
    }
    /// enqueue(queue, distance, swathe, ...), src/mls.cpp:101-135: distance = device float image
    void enqueue(float *distance, std::uint32_t pitch, const Swathe &swathe)
    {
        check(mlsb_mls_corners(worker.get(), distance, pitch, &swathe));
    }

private:
    Worker &worker;
};

/// ScaleBiasFilter, src/mesh_filter.h:118-168 (applied inside the weld pass).
class ScaleBiasFilter
{
public:
    explicit ScaleBiasFilter(Worker &worker) : worker(worker) {}
    void setScaleBias(float scale, float x, float y, float z)
    {
        const float bias[3] = {x, y, z};
        check(mlsb_set_scale_bias(worker.get(), scale, bias));
    }

private:
    Worker &worker;
};

/// DeviceKeyMesh, src/mesh.h:101-123: a block still on the device.
typedef mlsb_device_mesh DeviceKeyMesh;

/// MeshFilterChain, src/mesh_filter.h:69-116: device-side filters between marching and the output functor.
/// The chain lives in the worker (DeviceWorkerGroup's workers each own one, src/workers.h:193); filters are
/// held by reference and must outlive it.  MeshFilter: void(stream, inMesh, outMesh) — outMesh starts as a copy
/// of inMesh; device work goes on the stream (a cudaStream_t); throw to abort.
class MeshFilterChain
{
public:
    typedef std::function<void(void *stream, const DeviceKeyMesh &inMesh, DeviceKeyMesh &outMesh)> MeshFilter;

    explicit MeshFilterChain(Worker &worker) : worker(worker) {}
    ~MeshFilterChain() { mlsb_worker_clear_mesh_filters(worker.get()); }
    MeshFilterChain(const MeshFilterChain &) = delete;
    MeshFilterChain &operator=(const MeshFilterChain &) = delete;

    void addFilter(const MeshFilter &filter)
    {
        filters.emplace_back(new MeshFilter(filter));
        check(mlsb_worker_add_mesh_filter(worker.get(), &MeshFilterChain::trampoline, filters.back().get()));
    }

private:
    static int trampoline(void *user, const mlsb_device_mesh *in, mlsb_device_mesh *out, void *stream)
    {
        try
        {
            (*static_cast<const MeshFilter *>(user))(stream, *in, *out);
            return 0;
        }
        catch (...)
        {
            return 1;
        }
    }
    Worker &worker;
    std::vector<std::unique_ptr<MeshFilter>> filters;
};

/// Marching, src/marching.h:564-608.
class Marching
{
public:
    enum { MAX_CELL_VERTICES = 13, MAX_CELL_INDICES = 36, MAX_CELL_BYTES = 872, KEY_AXIS_BITS = 21 };
    /// OutputFunctor, src/marching.h:516-546, after the readback of src/mesh.cpp:62-102.
    typedef std::function<void(const HostKeyMesh &)> OutputFunctor;

    explicit Marching(Worker &worker) : worker(worker) {}

    /// generate(queue, generator, output, size, keyOffset, events), src/marching.cpp:745-824.
    /// The generator is the MlsFunctor bound with MlsFunctor::set.  BLOCKING.
    void generate(MlsFunctor &, const OutputFunctor &output, const std::uint32_t size[3], const std::uint32_t keyOffset[3])
    {
        check(mlsb_marching_generate(worker.get(), size, keyOffset, &Marching::trampoline, const_cast<OutputFunctor *>(&output)));
    }
    /// generate with a host-supplied field (test/test_marching.cpp:62-129 HostGenerator).
    void generate(const float *hostField, const OutputFunctor &output, const std::uint32_t size[3], const std::uint32_t keyOffset[3])
    {
        check(mlsb_marching_generate_host_field(worker.get(), hostField, size, keyOffset, &Marching::trampoline,
                                                const_cast<OutputFunctor *>(&output)));
    }

private:
    static void trampoline(const mlsb_mesh_sizes *sizes, const void *blob, void *user)
    {
        (*static_cast<const OutputFunctor *>(user))(HostKeyMesh(blob, *sizes));
    }
    Worker &worker;
};

}  // namespace mlsb

#endif /* MLSB200_HPP */
