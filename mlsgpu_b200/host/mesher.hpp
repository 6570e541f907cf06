// Host mesher: welds the bins' meshes at their seams and prunes small components.
//
// An in-core restatement of what OOCMesher does with the blocks it receives (src/mesher.cpp):
//  * per block, a LOCAL union-find over the triangle edges groups its vertices into "clumps"
//    (computeLocalComponents / updateGlobalClumps, :220-278) — the block's nodes fit in cache, so
//    the 2 merges per triangle never touch the global arrays;
//  * vertices that are external to a block carry a 64-bit key and are unified across blocks through
//    a key -> vertex map; unifying two vertices also unifies their clumps and un-counts the shared
//    vertex (updateClumpKeyMap, :280-306);
//  * at write time components (root clumps) with fewer than floor(fit-prune x total vertices)
//    vertices are dropped and the rest are compacted (getStatistics :491-513, write :763-852).
// The reference streams blocks through temporary files to bound memory.  Here the mesh stays in host
// memory by default, one piece per block, and only the seam bookkeeping (external vertices) is
// serialised: the per-block work runs concurrently on the device workers' threads.  With
// enableSpill() (--tmp-dir) the mesher is out of core like the reference's (src/mesher.cpp:343-445):
// a completed block's vertices, clump labels and triangles go to ONE temporary file at once (pwrite
// from the worker's thread), host memory keeps the key maps, the clumps and a few counters per block,
// and stream() prunes, compacts and writes the output file by reading the blocks back in parallel.
//
// --split (src/mesher.h:169-179, testChunk in test/test_mesher.cpp:924-993): every block belongs
// to an output chunk; vertices are welded within a chunk only (a vertex on a chunk boundary is
// written to both files), while components — and therefore pruning — are global.
#ifndef MLSB200_MESHER_HPP
#define MLSB200_MESHER_HPP

#include <algorithm>
#include <array>
#include <atomic>
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <istream>
#include <map>
#include <memory>
#include <ostream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include <fcntl.h>
#include <unistd.h>

#include "../../include/mlsb200.h"

namespace mlsb
{

class Mesher
{
public:
    typedef std::array<float, 3> vertex_type;
    typedef std::array<std::uint32_t, 3> triangle_type;
    typedef std::array<std::uint32_t, 3> chunk_type;   // ChunkId::coords

    /// split = true: blocks of different chunks share components but not vertices.
    explicit Mesher(bool split = false) : split(split) {}
    ~Mesher()
    {
        if (spillFd >= 0) ::close(spillFd);
    }
    Mesher(const Mesher &) = delete;
    Mesher &operator=(const Mesher &) = delete;
    Mesher(Mesher &&other) noexcept { *this = std::move(other); }
    Mesher &operator=(Mesher &&other) noexcept
    {
        if (this == &other) return *this;
        if (spillFd >= 0) ::close(spillFd);
        split = other.split;
        chunkIndex = std::move(other.chunkIndex);
        chunks = std::move(other.chunks);
        clumps = std::move(other.clumps);
        globalKeyClump = std::move(other.globalKeyClump);
        totalVertices = other.totalVertices;
        totalTriangles = other.totalTriangles;
        spillFd = other.spillFd;
        spillSize.store(other.spillSize.load());
        keepInCore = other.keepInCore;
        inCoreBytes.store(other.inCoreBytes.load());
        other.spillFd = -1;
        return *this;
    }

    /// Out-of-core mode (--tmp-dir): from now on spill() moves completed blocks into an unnamed
    /// temporary file in `dir`.  Call before the first block.
    /// keepBytes: blocks stay in host memory until that much is held; only the ones after that go to the
    /// file (0: every block does).  A mesh that fits is then never written twice.
    void enableSpill(const std::string &dir, std::uint64_t keepBytes = 0)
    {
        keepInCore = keepBytes;
        if (totalVertices != 0 || spillFd >= 0) throw std::logic_error("enableSpill() comes before the first block");
        std::string path = (dir.empty() ? std::string(".") : dir) + "/mlsgpu-mesher-XXXXXX";
        std::vector<char> name(path.begin(), path.end());
        name.push_back('\0');
        spillFd = ::mkstemp(name.data());
        if (spillFd < 0) throw std::runtime_error("cannot create a temporary file in " + dir + ": " + std::strerror(errno));
        ::unlink(name.data());   // the file lives as long as the descriptor
    }
    bool spilling() const { return spillFd >= 0; }
    /// Bytes written to the temporary file so far.
    std::uint64_t spilledBytes() const { return spillSize.load(); }

    /// One shipped-out block (HostKeyMesh layout, src/mesh.cpp:51-60) on its way into the mesh.
    /// The blob pointers are only used by prepare/commit/complete, i.e. inside the sink call.
    struct Block
    {
        std::size_t nv = 0, ni = 0, nt = 0;
        const std::uint64_t *keys = nullptr;
        const vertex_type *verts = nullptr;
        const triangle_type *tris = nullptr;
        std::vector<std::uint32_t> clumpOfLocal;     // local clump index of every vertex of the blob
        std::vector<std::uint64_t> clumpVertices;    // vertices per local clump
        std::vector<std::uint32_t> remapExternal;    // global index of every external vertex of the blob

        // what the mesher keeps: the block's internal vertices followed by the external vertices it
        // was the first of its chunk to see, numbered firstVertex.. in the chunk's welded mesh; its
        // triangles in welded numbering; the (global) clump of every kept vertex is
        // firstClump + vertexClump[i]
        std::uint32_t firstVertex = 0, firstClump = 0;
        std::vector<vertex_type> vertices;
        std::vector<std::uint32_t> vertexClump;
        std::vector<triangle_type> triangles;
        // per local clump: kept vertices (above) and triangles of this block — what pruning needs to
        // size the output without looking at the arrays
        std::vector<std::uint32_t> clumpKept, clumpTriangles;
        // out-of-core mode: where [vertices][vertexClump][triangles] went, and how many there are
        std::int64_t spillOffset = -1;
        std::size_t numKept = 0;
    };

    /// Arrays whose storage is handed from block to block on a thread instead of going back to the
    /// allocator: a block's scratch arrays always (complete() returns them), and in out-of-core mode
    /// the kept arrays as well (spill() returns them) — a fresh 3 MB vector per block costs more in
    /// page faults than filling it does.
    struct Pool
    {
        std::vector<std::uint32_t> clumpOfLocal, remapExternal, vertexClump;
        std::vector<vertex_type> vertices;
        std::vector<triangle_type> triangles;
    };
    static Pool &pool()
    {
        static thread_local Pool p;
        return p;
    }
    template <typename T> static void take(std::vector<T> &dst, std::vector<T> &pooled)
    {
        dst.swap(pooled);
        dst.clear();
    }
    template <typename T> static void give(std::vector<T> &src, std::vector<T> &pooled)
    {
        src.clear();
        if (src.capacity() > pooled.capacity()) pooled.swap(src);
        std::vector<T>().swap(src);
    }

    /// Step 1, thread-safe: local components and a private copy of the internal vertices.
    static std::unique_ptr<Block> prepare(const mlsb_mesh_sizes &sizes, const void *blob)
    {
        std::unique_ptr<Block> block(new Block);
        Block &b = *block;
        take(b.clumpOfLocal, pool().clumpOfLocal);
        take(b.remapExternal, pool().remapExternal);
        take(b.vertices, pool().vertices);
        take(b.vertexClump, pool().vertexClump);
        take(b.triangles, pool().triangles);
        const std::size_t nv = b.nv = sizes.numVertices, nt = b.nt = sizes.numTriangles;
        const std::size_t ni = b.ni = sizes.numInternalVertices;
        b.keys = static_cast<const std::uint64_t *>(blob);
        b.verts = reinterpret_cast<const vertex_type *>(b.keys + (nv - ni));
        b.tris = reinterpret_cast<const triangle_type *>(b.verts + nv);

        // ---- local components ----
        static thread_local std::vector<std::int32_t> local;
        local.resize(nv);
        for (std::size_t i = 0; i < nv; i++) local[i] = (std::int32_t) i;
        {
            std::int32_t *const parent = local.data();
            for (std::size_t i = 0; i < nt; i++)
            {
                // two edges are enough: the third is redundant (src/mesher.cpp:230-234)
                uniteLocal(parent, (std::int32_t) b.tris[i][0], (std::int32_t) b.tris[i][1]);
                uniteLocal(parent, (std::int32_t) b.tris[i][1], (std::int32_t) b.tris[i][2]);
            }
        }
        // ---- one clump per local component ----
        b.clumpOfLocal.resize(nv);
        for (std::size_t i = 0; i < nv; i++)
            if (local[i] == (std::int32_t) i)
            {
                b.clumpOfLocal[i] = (std::uint32_t) b.clumpVertices.size();
                b.clumpVertices.push_back(0);
            }
        // a root is the smallest index of its component, so its entry is final before any of its
        // members reads it
        // (counters are kept in a register over runs of the same clump: one component per block is the
        // common case, and a read-modify-write of the same counter per element serialises on it)
        {
            std::uint32_t last = 0;
            std::uint64_t run = 0;
            for (std::size_t i = 0; i < nv; i++)
            {
                const std::uint32_t c = b.clumpOfLocal[findLocal(local, (std::int32_t) i)];
                b.clumpOfLocal[i] = c;
                if (c != last)
                {
                    b.clumpVertices[last] += run;
                    last = c;
                    run = 0;
                }
                run++;
            }
            if (run) b.clumpVertices[last] += run;
        }
        b.vertices.reserve(nv);
        b.vertices.assign(b.verts, b.verts + ni);
        b.vertexClump.reserve(nv);
        b.vertexClump.assign(b.clumpOfLocal.begin(), b.clumpOfLocal.begin() + ni);
        b.clumpKept.assign(b.clumpVertices.size(), 0);
        b.clumpTriangles.assign(b.clumpVertices.size(), 0);
        countRuns(b.clumpKept, ni, [&](std::size_t i) { return b.clumpOfLocal[i]; });
        countRuns(b.clumpTriangles, nt, [&](std::size_t i) { return b.clumpOfLocal[b.tris[i][0]]; });   // a triangle lies in one clump
        return block;
    }

    /// Step 2, serialised by the caller: numbers the block's vertices and clumps and resolves its
    /// external vertices through the key map.  The mesher takes ownership of the block.
    Block &commit(std::unique_ptr<Block> block, const chunk_type &chunkId = chunk_type{{0, 0, 0}})
    {
        Block &b = *block;
        const std::size_t nv = b.nv, ni = b.ni;
        auto found = chunkIndex.find(chunkId);
        if (found == chunkIndex.end())
        {
            found = chunkIndex.insert(std::make_pair(chunkId, chunks.size())).first;
            chunks.emplace_back(new Chunk);
            chunks.back()->id = chunkId;
        }
        Chunk &chunk = *chunks[found->second];
        // vertex and clump ids are 32-bit, as in the reference (it aborts with "too many connected
        // components", src/mesher.cpp:237-262): fail loudly instead of wrapping into a corrupt file
        if (chunk.totalVertices + nv > (std::size_t) UINT32_MAX)
            throw std::overflow_error("too many vertices in one output file (2^32): use --split");
        if (clumps.size() + b.clumpVertices.size() > (std::size_t) UINT32_MAX)
            throw std::overflow_error("too many connected components");
        b.firstClump = (std::uint32_t) clumps.size();
        for (std::size_t c = 0; c < b.clumpVertices.size(); c++)
            clumps.push_back(Clump{b.firstClump + (std::uint32_t) c, b.clumpVertices[c]});
        b.firstVertex = (std::uint32_t) chunk.totalVertices;
        b.remapExternal.resize(nv - ni);
        for (std::size_t i = ni; i < nv; i++)
        {
            const std::uint32_t clump = b.firstClump + b.clumpOfLocal[i];
            const std::uint64_t key = b.keys[i - ni];
            const Seen fresh = {b.firstVertex + (std::uint32_t) b.vertices.size(), clump};
            auto added = chunk.keyMap.insert(std::make_pair(key, fresh));
            std::uint32_t earlierClump = added.first->second.clump;
            bool seenBefore = !added.second;
            if (added.second)
            {
                b.vertices.push_back(b.verts[i]);
                b.vertexClump.push_back(b.clumpOfLocal[i]);
                b.clumpKept[b.clumpOfLocal[i]]++;
                if (split)
                {
                    // new to this chunk's file, but another chunk may hold the same vertex
                    auto global = globalKeyClump.insert(std::make_pair(key, clump));
                    seenBefore = !global.second;
                    earlierClump = global.first->second;
                }
            }
            if (seenBefore)
            {
                // the same vertex seen from two blocks: unify their clumps, count it once
                const std::uint32_t root = mergeClumps(clump, earlierClump);
                clumps[root].vertices--;
            }
            b.remapExternal[i - ni] = added.first->second.vertex;
        }
        b.numKept = b.vertices.size();
        chunk.totalVertices += b.vertices.size();
        chunk.totalTriangles += b.nt;
        totalVertices += b.vertices.size();
        totalTriangles += b.nt;
        chunk.blocks.push_back(std::move(block));
        return b;
    }

    /// Step 3, thread-safe (touches only the block): triangles in welded numbering.
    static void complete(Block &b)
    {
        const std::size_t ni = b.ni;
        const std::uint32_t firstVertex = b.firstVertex;
        const std::uint32_t *ext = b.remapExternal.data();
        auto remap = [=](std::uint32_t v) { return v < ni ? firstVertex + v : ext[v - ni]; };
        // (resize + indexed stores: reserve + push_back avoids writing the array twice but measured twice as slow)
        b.triangles.resize(b.nt);
        triangle_type *const out = b.triangles.data();
        const triangle_type *const in = b.tris;
        for (std::size_t i = 0; i < b.nt; i++)
            out[i] = triangle_type{{remap(in[i][0]), remap(in[i][1]), remap(in[i][2])}};
        // scratch that is not needed any more
        give(b.clumpOfLocal, pool().clumpOfLocal);
        give(b.remapExternal, pool().remapExternal);
        b.keys = nullptr;
        b.verts = nullptr;
        b.tris = nullptr;
    }

    /// Step 4, out-of-core mode only, thread-safe: the completed block's arrays move to the temporary
    /// file (one reservation of the file's tail, then positioned writes).
    void spill(Block &b)
    {
        if (spillFd < 0) return;
        const std::size_t bytesV = b.vertices.size() * sizeof(vertex_type), bytesC = b.vertexClump.size() * sizeof(std::uint32_t);
        const std::size_t bytesT = b.triangles.size() * sizeof(triangle_type);
        if (inCoreBytes.load() + bytesV + bytesC + bytesT <= keepInCore)
        {
            inCoreBytes.fetch_add(bytesV + bytesC + bytesT);   // within the budget: the block keeps its arrays
            return;
        }
        b.spillOffset = (std::int64_t) spillSize.fetch_add(bytesV + bytesC + bytesT);
        writeAt(b.vertices.data(), bytesV, b.spillOffset);
        writeAt(b.vertexClump.data(), bytesC, b.spillOffset + (std::int64_t) bytesV);
        writeAt(b.triangles.data(), bytesT, b.spillOffset + (std::int64_t) (bytesV + bytesC));
        give(b.vertices, pool().vertices);
        give(b.vertexClump, pool().vertexClump);
        give(b.triangles, pool().triangles);
    }

    /// All steps, for a single-threaded consumer.
    void add(const mlsb_mesh_sizes &sizes, const void *blob)
    {
        Block &b = commit(prepare(sizes, blob));
        complete(b);
        spill(b);
    }

    std::size_t numVertices() const { return totalVertices; }
    std::size_t numTriangles() const { return totalTriangles; }
    bool isSplit() const { return split; }

    /// --checkpoint (OOCMesher::checkpoint, src/mesher.cpp:854-874): everything finish() needs —
    /// clumps, and per chunk its blocks' vertices, clumps and triangles — so that pruning and writing
    /// can run later, possibly with another --fit-prune.  Call after the last block has completed.
    void save(std::ostream &out) const
    {
        if (spilling()) throw std::logic_error("--checkpoint keeps the mesh in one file already: not with --tmp-dir");
        out.write(checkpointMagic(), 8);
        put<std::uint64_t>(out, split);
        put<std::uint64_t>(out, totalVertices);
        put<std::uint64_t>(out, totalTriangles);
        put<std::uint64_t>(out, clumps.size());
        for (const Clump &c : clumps)
        {
            put<std::uint64_t>(out, c.parent);
            put<std::uint64_t>(out, c.vertices);
        }
        put<std::uint64_t>(out, chunks.size());
        for (const auto &entry : chunkIndex)
        {
            const Chunk &chunk = *chunks[entry.second];
            for (int a = 0; a < 3; a++) put<std::uint64_t>(out, chunk.id[a]);
            put<std::uint64_t>(out, chunk.totalVertices);
            put<std::uint64_t>(out, chunk.totalTriangles);
            put<std::uint64_t>(out, chunk.blocks.size());
            for (const std::unique_ptr<Block> &b : chunk.blocks)
            {
                put<std::uint64_t>(out, b->firstVertex);
                put<std::uint64_t>(out, b->firstClump);
                putVector(out, b->vertices);
                putVector(out, b->vertexClump);
                putVector(out, b->triangles);
            }
        }
        if (!out) throw std::runtime_error("could not write the checkpoint");
    }

    /// --resume (OOCMesher::resume, src/mesher.cpp:876-898): replaces the state with a saved one.
    void load(std::istream &in)
    {
        char magic[8];
        in.read(magic, 8);
        if (!in || std::memcmp(magic, checkpointMagic(), 8) != 0) throw std::runtime_error("not a mesher checkpoint");
        *this = Mesher(get<std::uint64_t>(in) != 0);
        totalVertices = (std::size_t) get<std::uint64_t>(in);
        totalTriangles = (std::size_t) get<std::uint64_t>(in);
        clumps.resize((std::size_t) get<std::uint64_t>(in));
        for (Clump &c : clumps)
        {
            c.parent = (std::uint32_t) get<std::uint64_t>(in);
            c.vertices = get<std::uint64_t>(in);
        }
        const std::uint64_t numChunks = get<std::uint64_t>(in);
        for (std::uint64_t i = 0; i < numChunks; i++)
        {
            chunks.emplace_back(new Chunk);
            Chunk &chunk = *chunks.back();
            for (int a = 0; a < 3; a++) chunk.id[a] = (std::uint32_t) get<std::uint64_t>(in);
            chunkIndex[chunk.id] = chunks.size() - 1;
            chunk.totalVertices = (std::size_t) get<std::uint64_t>(in);
            chunk.totalTriangles = (std::size_t) get<std::uint64_t>(in);
            chunk.blocks.resize((std::size_t) get<std::uint64_t>(in));
            for (std::unique_ptr<Block> &b : chunk.blocks)
            {
                b.reset(new Block);
                b->firstVertex = (std::uint32_t) get<std::uint64_t>(in);
                b->firstClump = (std::uint32_t) get<std::uint64_t>(in);
                getVector(in, b->vertices);
                getVector(in, b->vertexClump);
                getVector(in, b->triangles);
                b->nt = b->triangles.size();
                b->numKept = b->vertices.size();
            }
        }
        if (!in) throw std::runtime_error("truncated mesher checkpoint");
    }

    /// One output file: the welded mesh of a chunk in pieces (one per block); vertex indices run
    /// across the pieces in order.  The component counts describe the whole mesh, not the chunk.
    struct Result
    {
        chunk_type chunk = chunk_type{{0, 0, 0}};
        std::vector<std::vector<vertex_type>> vertexChunks;
        std::vector<std::vector<triangle_type>> triangleChunks;
        std::size_t numVertices = 0, numTriangles = 0;
        std::size_t totalComponents = 0, keptComponents = 0;
        std::uint64_t thresholdVertices = 0;
    };

    /// Prune and compact; one Result for the whole mesh (no --split).
    Result finish(double pruneThreshold)
    {
        std::vector<Result> all = finishChunks(pruneThreshold);
        if (all.empty())
        {
            Result r;
            countComponents(pruneThreshold, r, nullptr);
            return r;
        }
        return std::move(all[0]);
    }

    /// Prune and compact, one Result per chunk in chunk order.  pruneThreshold = --fit-prune
    /// (fraction of all welded vertices, counted over the whole mesh).
    std::vector<Result> finishChunks(double pruneThreshold)
    {
        if (spilling()) throw std::logic_error("an out-of-core mesher is finished with stream()");
        Result counts;
        std::vector<char> keepClump;
        countComponents(pruneThreshold, counts, &keepClump);
        std::vector<Result> out;
        for (const auto &entry : chunkIndex)   // std::map: sorted by chunk coordinates
        {
            Result r = counts;
            r.chunk = entry.first;
            finishChunk(*chunks[entry.second], keepClump, r);
            out.push_back(std::move(r));
        }
        return out;
    }

    /// Prune, compact and WRITE, chunk by chunk, without holding a chunk in memory (works in core as
    /// well).  `open(chunk, numVertices, numTriangles)` returns a writer with thread-safe
    /// `vertices(first, data, count)` and `triangles(first, data, count)` (PlyStreamWriter), or a null
    /// pointer to skip the chunk (--split writes only chunks with triangles); the Results carry the
    /// counts only.  OOCMesher::write, src/mesher.cpp:763-852.
    template <typename Open>
    std::vector<Result> stream(double pruneThreshold, Open open)
    {
        Result counts;
        std::vector<char> keepClump;
        countComponents(pruneThreshold, counts, &keepClump);
        const bool prune = counts.keptComponents != counts.totalComponents;
        std::vector<Result> out;
        for (const auto &entry : chunkIndex)
        {
            Result r = counts;
            r.chunk = entry.first;
            Chunk &chunk = *chunks[entry.second];
            const std::int64_t numBlocks = (std::int64_t) chunk.blocks.size();
            // kept vertices and triangles per block, from the per-clump counters
            std::vector<std::size_t> firstV(chunk.blocks.size() + 1, 0), firstT(chunk.blocks.size() + 1, 0);
            for (std::int64_t i = 0; i < numBlocks; i++)
            {
                const Block &b = *chunk.blocks[i];
                std::size_t kv = b.numKept, kt = b.nt;
                if (prune)
                {
                    kv = kt = 0;
                    for (std::size_t c = 0; c < b.clumpKept.size(); c++)
                        if (keepClump[b.firstClump + c]) kv += b.clumpKept[c], kt += b.clumpTriangles[c];
                }
                firstV[i + 1] = firstV[i] + kv;
                firstT[i + 1] = firstT[i] + kt;
            }
            r.numVertices = firstV.back();
            r.numTriangles = firstT.back();
            auto writer = open(r.chunk, r.numVertices, r.numTriangles);
            if (writer)
            {
                std::vector<std::uint32_t> newIndex;
                if (prune) newIndex.assign(chunk.totalVertices, UINT32_MAX);
                // (an exception must not leave a worksharing loop: every iteration catches its own, notes the
                // first message and lets the remaining iterations fall through)
                std::string failure;
                std::atomic<bool> failed(false);
                auto guarded = [&](const std::function<void()> &body) {
                    if (failed.load()) return;
                    try
                    {
                        body();
                    }
                    catch (const std::exception &e)
                    {
                        failed.store(true);
#pragma omp critical(mlsb_mesher_stream_failure)
                        if (failure.empty()) failure = e.what();
                    }
                };
#pragma omp parallel
                {
                    std::vector<vertex_type> verts;
                    std::vector<std::uint32_t> labels;
                    std::vector<triangle_type> tris;
#pragma omp for schedule(dynamic, 1)
                    for (std::int64_t i = 0; i < numBlocks; i++)
                        guarded([&] {
                            const Block &b = *chunk.blocks[i];
                            const vertex_type *v = fetch(b, 0, b.numKept, b.vertices, verts);
                            if (!prune)
                            {
                                writer->vertices(firstV[i], v, b.numKept);
                                return;
                            }
                            const std::uint32_t *l = fetch(b, b.numKept * sizeof(vertex_type), b.numKept, b.vertexClump, labels);
                            std::size_t n = 0;
                            if (v != verts.data())
                            {
                                verts.assign(v, v + b.numKept);   // in core: compact a copy, the block stays intact
                                v = verts.data();
                            }
                            for (std::size_t k = 0; k < b.numKept; k++)
                                if (keepClump[b.firstClump + l[k]])
                                {
                                    newIndex[b.firstVertex + k] = (std::uint32_t) (firstV[i] + n);
                                    verts[n++] = v[k];
                                }
                            writer->vertices(firstV[i], verts.data(), n);
                        });
                    // (the loop's implicit barrier: every newIndex is final before a triangle reads it)
#pragma omp for schedule(dynamic, 1)
                    for (std::int64_t i = 0; i < numBlocks; i++)
                        guarded([&] {
                            const Block &b = *chunk.blocks[i];
                            const triangle_type *t = fetch(b, b.numKept * (sizeof(vertex_type) + sizeof(std::uint32_t)), b.nt, b.triangles, tris);
                            if (!prune)
                            {
                                writer->triangles(firstT[i], t, b.nt);
                                return;
                            }
                            if (t != tris.data())
                            {
                                tris.assign(t, t + b.nt);
                                t = tris.data();
                            }
                            std::size_t n = 0;
                            for (std::size_t k = 0; k < b.nt; k++)
                                if (newIndex[t[k][0]] != UINT32_MAX)   // a triangle's vertices share one component
                                    tris[n++] = triangle_type{{newIndex[t[k][0]], newIndex[t[k][1]], newIndex[t[k][2]]}};
                            writer->triangles(firstT[i], tris.data(), n);
                        });
                }
                if (!failure.empty()) throw std::runtime_error(failure);
                writer->close();
            }
            out.push_back(std::move(r));
        }
        if (out.empty()) out.push_back(counts);
        return out;
    }

private:
    static const char *checkpointMagic() { return "MLSBCKP1"; }
    // positioned I/O on the temporary file
    void writeAt(const void *data, std::size_t bytes, std::int64_t offset) const
    {
        const char *p = static_cast<const char *>(data);
        while (bytes > 0)
        {
            const ssize_t n = ::pwrite(spillFd, p, bytes, (off_t) offset);
            if (n < 0 && errno == EINTR) continue;
            if (n <= 0) throw std::runtime_error(std::string("error writing the mesher's temporary file: ") + std::strerror(errno));
            p += n;
            offset += n;
            bytes -= (std::size_t) n;
        }
    }
    void readAt(void *data, std::size_t bytes, std::int64_t offset) const
    {
        char *p = static_cast<char *>(data);
        while (bytes > 0)
        {
            const ssize_t n = ::pread(spillFd, p, bytes, (off_t) offset);
            if (n < 0 && errno == EINTR) continue;
            if (n <= 0) throw std::runtime_error("error reading the mesher's temporary file");
            p += n;
            offset += n;
            bytes -= (std::size_t) n;
        }
    }
    /// `count` elements of a block's array: the array itself in core, else read from the temporary
    /// file (at `skip` bytes into the block's record) into `scratch`.
    template <typename T>
    const T *fetch(const Block &b, std::size_t skip, std::size_t count, const std::vector<T> &inCore, std::vector<T> &scratch) const
    {
        if (b.spillOffset < 0) return inCore.data();
        scratch.resize(count);
        if (count) readAt(scratch.data(), count * sizeof(T), b.spillOffset + (std::int64_t) skip);
        return scratch.data();
    }
    template <typename T> static void put(std::ostream &out, T value) { out.write(reinterpret_cast<const char *>(&value), sizeof(T)); }
    template <typename T> static T get(std::istream &in)
    {
        T value = T();
        in.read(reinterpret_cast<char *>(&value), sizeof(T));
        if (!in) throw std::runtime_error("truncated mesher checkpoint");
        return value;
    }
    template <typename T> static void putVector(std::ostream &out, const std::vector<T> &v)
    {
        put<std::uint64_t>(out, v.size());
        if (!v.empty()) out.write(reinterpret_cast<const char *>(v.data()), (std::streamsize) (v.size() * sizeof(T)));
    }
    template <typename T> static void getVector(std::istream &in, std::vector<T> &v)
    {
        v.resize((std::size_t) get<std::uint64_t>(in));
        if (!v.empty()) in.read(reinterpret_cast<char *>(v.data()), (std::streamsize) (v.size() * sizeof(T)));
    }

    struct Clump
    {
        std::uint32_t parent;
        std::uint64_t vertices;
    };
    struct Seen
    {
        std::uint32_t vertex;   // welded index of an external vertex within its chunk
        std::uint32_t clump;    // the clump it belonged to where it was first seen
    };
    struct Chunk
    {
        chunk_type id;
        std::unordered_map<std::uint64_t, Seen> keyMap;
        std::vector<std::unique_ptr<Block>> blocks;
        std::size_t totalVertices = 0, totalTriangles = 0;
    };

    /// getStatistics, src/mesher.cpp:491-513.  keepClump (optional): every clump's verdict.
    void countComponents(double pruneThreshold, Result &r, std::vector<char> *keepClump)
    {
        std::uint64_t total = 0;
        for (std::size_t c = 0; c < clumps.size(); c++)
            if (clumps[c].parent == c) total += clumps[c].vertices;
        r.thresholdVertices = (std::uint64_t) (total * pruneThreshold);
        std::vector<char> keep(clumps.size(), 0);
        for (std::size_t c = 0; c < clumps.size(); c++)
            if (clumps[c].parent == c)
            {
                r.totalComponents++;
                keep[c] = clumps[c].vertices >= r.thresholdVertices;
                r.keptComponents += keep[c];
            }
        if (keepClump && r.keptComponents != r.totalComponents)
        {
            keepClump->resize(clumps.size());
            for (std::size_t c = 0; c < clumps.size(); c++) (*keepClump)[c] = keep[findClump((std::uint32_t) c)];
        }
    }

    void finishChunk(Chunk &chunk, const std::vector<char> &keepClump, Result &r)
    {
        std::vector<std::unique_ptr<Block>> &blocks = chunk.blocks;
        const std::size_t totalVertices = chunk.totalVertices, totalTriangles = chunk.totalTriangles;
        const std::int64_t numBlocks = (std::int64_t) blocks.size();
        r.vertexChunks.resize(blocks.size());
        r.triangleChunks.resize(blocks.size());
        if (r.keptComponents == r.totalComponents)
        {
            // nothing to prune: hand the pieces over as they are
            for (std::int64_t i = 0; i < numBlocks; i++)
            {
                r.vertexChunks[i].swap(blocks[i]->vertices);
                r.triangleChunks[i].swap(blocks[i]->triangles);
            }
            r.numVertices = totalVertices;
            r.numTriangles = totalTriangles;
            return;
        }
        // pass 1: kept vertices per block; prefix; pass 2: new numbering; pass 3: triangles
        std::vector<std::size_t> firstKept(blocks.size() + 1, 0);
#pragma omp parallel for schedule(dynamic, 1)
        for (std::int64_t i = 0; i < numBlocks; i++)
        {
            const Block &b = *blocks[i];
            std::size_t kept = 0;
            for (std::uint32_t c : b.vertexClump) kept += keepClump[b.firstClump + c];
            firstKept[i + 1] = kept;
        }
        for (std::int64_t i = 0; i < numBlocks; i++) firstKept[i + 1] += firstKept[i];
        std::vector<std::uint32_t> newIndex(totalVertices, UINT32_MAX);
#pragma omp parallel for schedule(dynamic, 1)
        for (std::int64_t i = 0; i < numBlocks; i++)
        {
            const Block &b = *blocks[i];
            std::vector<vertex_type> &out = r.vertexChunks[i];
            out.reserve(firstKept[i + 1] - firstKept[i]);
            for (std::size_t v = 0; v < b.vertices.size(); v++)
                if (keepClump[b.firstClump + b.vertexClump[v]])
                {
                    newIndex[b.firstVertex + v] = (std::uint32_t) (firstKept[i] + out.size());
                    out.push_back(b.vertices[v]);
                }
        }
#pragma omp parallel for schedule(dynamic, 1)
        for (std::int64_t i = 0; i < numBlocks; i++)
        {
            std::vector<triangle_type> &out = r.triangleChunks[i];
            for (const triangle_type &t : blocks[i]->triangles)
                if (newIndex[t[0]] != UINT32_MAX)   // a triangle's vertices share one component
                {
                    triangle_type o = {newIndex[t[0]], newIndex[t[1]], newIndex[t[2]]};
                    out.push_back(o);
                }
        }
        r.numVertices = firstKept[blocks.size()];
        for (const std::vector<triangle_type> &c : r.triangleChunks) r.numTriangles += c.size();
    }

    static std::int32_t findLocal(std::vector<std::int32_t> &local, std::int32_t x)
    {
        while (local[x] != x)
        {
            local[x] = local[local[x]];
            x = local[x];
        }
        return x;
    }
    /// Union by Rem's algorithm (interleaved finds with splicing): parents always have the smaller
    /// index, so a root is the smallest vertex of its component, as prepare() relies on.
    static void uniteLocal(std::int32_t *parent, std::int32_t x, std::int32_t y)
    {
        while (parent[x] != parent[y])
        {
            if (parent[x] < parent[y]) std::swap(x, y);
            if (x == parent[x])
            {
                parent[x] = parent[y];
                return;
            }
            const std::int32_t z = parent[x];
            parent[x] = parent[y];
            x = z;
        }
    }
    /// counts[label(i)]++ for i < n, with the counter of the current run of equal labels in a register.
    template <typename Label>
    static void countRuns(std::vector<std::uint32_t> &counts, std::size_t n, Label label)
    {
        std::uint32_t last = 0, run = 0;
        for (std::size_t i = 0; i < n; i++)
        {
            const std::uint32_t c = label(i);
            if (c != last)
            {
                counts[last] += run;
                last = c;
                run = 0;
            }
            run++;
        }
        if (run) counts[last] += run;
    }
    std::uint32_t findClump(std::uint32_t x)
    {
        while (clumps[x].parent != x)
        {
            clumps[x].parent = clumps[clumps[x].parent].parent;
            x = clumps[x].parent;
        }
        return x;
    }
    std::uint32_t mergeClumps(std::uint32_t a, std::uint32_t b)
    {
        a = findClump(a);
        b = findClump(b);
        if (a == b) return a;
        const std::uint32_t lo = std::min(a, b), hi = std::max(a, b);
        clumps[hi].parent = lo;
        clumps[lo].vertices += clumps[hi].vertices;
        return lo;
    }

    bool split;
    std::map<chunk_type, std::size_t> chunkIndex;
    std::vector<std::unique_ptr<Chunk>> chunks;
    std::vector<Clump> clumps;
    std::unordered_map<std::uint64_t, std::uint32_t> globalKeyClump;   // split only: key -> clump where first seen
    std::size_t totalVertices = 0, totalTriangles = 0;
    int spillFd = -1;                          // out-of-core mode: the temporary file
    std::atomic<std::uint64_t> spillSize{0};   // its length (blocks reserve their place with a fetch_add)
    std::uint64_t keepInCore = 0;              // bytes of block arrays that may stay in host memory before blocks spill
    std::atomic<std::uint64_t> inCoreBytes{0};
};

/// ChunkNamer (src/mesher.cpp:88-96): base_XXXX_YYYY_ZZZZ.ply, at least four digits per coordinate.
inline std::string chunkName(const std::string &baseName, const Mesher::chunk_type &chunk)
{
    char buf[64];
    std::snprintf(buf, sizeof(buf), "_%04u_%04u_%04u.ply", chunk[0], chunk[1], chunk[2]);
    return baseName + buf;
}

}  // namespace mlsb

#endif
