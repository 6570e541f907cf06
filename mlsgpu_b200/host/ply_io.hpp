// Binary PLY input (splats) and output (mesh), the formats of src/fast_ply.{h,cpp}.
//
// Reader: little-endian binary PLY whose vertex element has float32 x, y, z, nx, ny, nz, radius in
// any order, possibly among other fixed-size properties (src/fast_ply.cpp:185-325).  decode()
// applies the conditioning of FastPly::Reader::decode (:334-350).
// Writer: float32 x, y, z vertices and "list uint8 uint32 vertex_indices" faces with the header
// padded to a multiple of four bytes by a trailing comment (:443-481).
#ifndef MLSB200_PLY_IO_HPP
#define MLSB200_PLY_IO_HPP

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fcntl.h>
#include <cerrno>
#include <limits>
#include <memory>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <sys/mman.h>
#include <unistd.h>
#include <vector>

#include "../../include/mlsb200.h"

namespace mlsb
{

class PlyReader
{
public:
    PlyReader(const std::string &path, float smooth, float maxRadius) : path(path), smooth(smooth), maxRadius(maxRadius)
    {
        file = std::fopen(path.c_str(), "rb");
        if (!file) throw std::runtime_error("cannot open " + path);
        try
        {
            readHeader();
        }
        catch (...)
        {
            std::fclose(file);
            throw;
        }
    }
    ~PlyReader()
    {
        if (file) std::fclose(file);
        if (preadFd >= 0) ::close(preadFd);
    }
    PlyReader(const PlyReader &) = delete;
    PlyReader &operator=(const PlyReader &) = delete;

    std::size_t size() const { return numVertices; }
    std::size_t vertexBytes() const { return vertexSize; }   // FastPly::Reader::getVertexSize

    /// Reads vertices [first, last) and appends the finite ones as conditioned splats
    /// (non-finite splats are dropped, src/splat_set.cpp:221-240).
    void read(std::size_t first, std::size_t last, std::vector<mlsb_splat> &out)
    {
        std::vector<char> buffer(std::min<std::size_t>(last - first, 1u << 16) * vertexSize);
        std::fseek(file, (long) (headerSize + first * vertexSize), SEEK_SET);
        while (first < last)
        {
            const std::size_t n = std::min<std::size_t>(last - first, buffer.size() / vertexSize);
            if (std::fread(buffer.data(), vertexSize, n, file) != n) throw std::runtime_error("short read in " + path);
            for (std::size_t i = 0; i < n; i++)
            {
                mlsb_splat s = decode(buffer.data() + i * vertexSize);
                bool finite = true;
                const float *f = reinterpret_cast<const float *>(&s);
                for (int k = 0; k < 8; k++) finite = finite && std::isfinite(f[k]);
                if (finite) out.push_back(s);
            }
            first += n;
        }
    }

    /// Decodes vertices [first, first + count) into out[0 .. count) — every slot is written, finite or
    /// not — with OpenMP over the vertices of a memory-mapped file (FileSet's reader thread + OpenMP
    /// decode, src/splat_set.cpp:212-219).  Returns false if the file cannot be mapped.
    bool readRaw(std::size_t first, std::size_t count, mlsb_splat *out) const
    {
        if (count == 0) return true;
        const int fd = ::open(path.c_str(), O_RDONLY);
        if (fd < 0) return false;
        const std::size_t begin = headerSize + first * vertexSize, bytes = count * vertexSize;
        const std::size_t page = (std::size_t) ::sysconf(_SC_PAGESIZE), aligned = begin / page * page;
        void *map = ::mmap(nullptr, bytes + (begin - aligned), PROT_READ, MAP_PRIVATE, fd, (off_t) aligned);
        ::close(fd);
        if (map == MAP_FAILED) return false;
        const char *base = static_cast<const char *>(map) + (begin - aligned);
#pragma omp parallel for schedule(static)
        for (std::int64_t i = 0; i < (std::int64_t) count; i++) out[i] = decode(base + (std::size_t) i * vertexSize);
        ::munmap(map, bytes + (begin - aligned));
        return true;
    }

    /// Decodes vertices [first, first + count) into out[0 .. count) on the calling thread (pread into
    /// `scratch`): safe to call from several threads at once, which is how the streaming front-end
    /// walks a file chunk by chunk.
    void readRange(std::size_t first, std::size_t count, mlsb_splat *out, std::vector<char> &scratch) const
    {
        if (count == 0) return;
        scratch.resize(count * vertexSize);
        if (preadFd < 0)
        {
            const int fd = ::open(path.c_str(), O_RDONLY);
            if (fd < 0) throw std::runtime_error("cannot open " + path);
            int expected = -1;
            if (!__atomic_compare_exchange_n(&preadFd, &expected, fd, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) ::close(fd);
        }
        std::size_t done = 0;
        while (done < scratch.size())
        {
            const ssize_t got = ::pread(preadFd, scratch.data() + done, scratch.size() - done, (off_t) (headerSize + first * vertexSize + done));
            if (got <= 0) throw std::runtime_error("short read in " + path);
            done += (std::size_t) got;
        }
        for (std::size_t i = 0; i < count; i++) out[i] = decode(scratch.data() + i * vertexSize);
    }

    static bool isFinite(const mlsb_splat &s)
    {
        const float *f = reinterpret_cast<const float *>(&s);
        bool finite = true;
        for (int k = 0; k < 8; k++) finite = finite && std::isfinite(f[k]);
        return finite;
    }

private:
    mlsb_splat decode(const char *p) const
    {
        mlsb_splat s;
        std::memcpy(&s.position[0], p + offsets[0], 4);
        std::memcpy(&s.position[1], p + offsets[1], 4);
        std::memcpy(&s.position[2], p + offsets[2], 4);
        std::memcpy(&s.normal[0], p + offsets[3], 4);
        std::memcpy(&s.normal[1], p + offsets[4], 4);
        std::memcpy(&s.normal[2], p + offsets[5], 4);
        std::memcpy(&s.radius, p + offsets[6], 4);
        s.radius = std::min(s.radius, maxRadius);
        s.radius *= smooth;
        s.quality = (float) (1.0 / (s.radius * s.radius));   // double division, as upstream
        return s;
    }

    static std::size_t typeSize(const std::string &t)
    {
        static const std::map<std::string, std::size_t> sizes = {
            {"char", 1}, {"int8", 1}, {"uchar", 1}, {"uint8", 1}, {"short", 2}, {"int16", 2}, {"ushort", 2}, {"uint16", 2},
            {"int", 4}, {"int32", 4}, {"uint", 4}, {"uint32", 4}, {"float", 4}, {"float32", 4}, {"double", 8}, {"float64", 8}};
        auto it = sizes.find(t);
        if (it == sizes.end()) throw std::runtime_error("unknown PLY type " + t);
        return it->second;
    }

    void readHeader()
    {
        static const char *names[7] = {"x", "y", "z", "nx", "ny", "nz", "radius"};
        std::string line;
        bool inVertex = false, sawVertex = false, little = false;
        bool have[7] = {};
        auto getline = [&]() {
            line.clear();
            int c;
            while ((c = std::fgetc(file)) != EOF && c != '\n') line.push_back((char) c);
            if (!line.empty() && line.back() == '\r') line.pop_back();
            return c != EOF;
        };
        if (!getline() || line != "ply") throw std::runtime_error(path + ": PLY signature missing");
        while (getline())
        {
            std::istringstream in(line);
            std::string tok;
            in >> tok;
            if (tok == "end_header") break;
            if (tok == "format")
            {
                std::string fmt;
                in >> fmt;
                if (fmt != "binary_little_endian") throw std::runtime_error(path + ": only binary_little_endian is supported");
                little = true;
            }
            else if (tok == "element")
            {
                std::string name;
                std::size_t count;
                in >> name >> count;
                if (inVertex) inVertex = false;
                if (name == "vertex")
                {
                    if (sawVertex) throw std::runtime_error(path + ": two vertex elements");
                    inVertex = sawVertex = true;
                    numVertices = count;
                }
                else if (!sawVertex)
                    throw std::runtime_error(path + ": the vertex element must come first");
            }
            else if (tok == "property" && inVertex)
            {
                std::string type, name;
                in >> type;
                if (type == "list") throw std::runtime_error(path + ": list property in vertex element");
                in >> name;
                for (int k = 0; k < 7; k++)
                    if (name == names[k])
                    {
                        if (type != "float" && type != "float32") throw std::runtime_error(path + ": property " + name + " must be float32");
                        offsets[k] = vertexSize;
                        have[k] = true;
                    }
                vertexSize += typeSize(type);
            }
        }
        if (!little) throw std::runtime_error(path + ": no format line");
        for (int k = 0; k < 7; k++)
            if (!have[k]) throw std::runtime_error(path + ": property " + names[k] + " is missing");
        headerSize = (std::size_t) std::ftell(file);
    }

    std::string path;
    std::FILE *file = nullptr;
    mutable int preadFd = -1;
    float smooth, maxRadius;
    std::size_t numVertices = 0, vertexSize = 0, headerSize = 0;
    std::size_t offsets[7] = {};
};

/// FastPly::Writer, src/fast_ply.cpp:398-564.
/// An output file of known size that is filled in pieces, from any thread, at any position: what the
/// out-of-core mesher writes into (Mesher::stream).  Made by PlyWriter::openStream.
class PlyStreamWriter
{
public:
    typedef std::array<float, 3> vertex_type;
    typedef std::array<std::uint32_t, 3> triangle_type;

    PlyStreamWriter(const std::string &path, const std::string &header, std::size_t numVertices, std::size_t numTriangles)
        : path(path), vertexOffset(header.size()), faceOffset(header.size() + 12 * numVertices)
    {
        fd = ::open(path.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0666);
        if (fd < 0) throw std::runtime_error("cannot create " + path);
        if (::ftruncate(fd, (off_t) (faceOffset + faceBytes * numTriangles)) != 0)
        {
            ::close(fd);
            throw std::runtime_error("cannot size " + path);
        }
        put(header.data(), header.size(), 0);
    }
    ~PlyStreamWriter()
    {
        if (fd >= 0) ::close(fd);
    }
    PlyStreamWriter(const PlyStreamWriter &) = delete;
    PlyStreamWriter &operator=(const PlyStreamWriter &) = delete;

    /// Vertices first .. first + count of the file.  Thread-safe.
    void vertices(std::size_t first, const vertex_type *data, std::size_t count)
    {
        put(data, 12 * count, vertexOffset + 12 * first);
    }
    /// Faces first .. first + count of the file.  Thread-safe.
    void triangles(std::size_t first, const triangle_type *data, std::size_t count)
    {
        const std::size_t slice = 1 << 16;
        std::vector<unsigned char> packed(faceBytes * std::min(slice, count));
        for (std::size_t done = 0; done < count; done += slice)
        {
            const std::size_t n = std::min(slice, count - done);
            for (std::size_t i = 0; i < n; i++)
            {
                packed[faceBytes * i] = 3;
                std::memcpy(packed.data() + faceBytes * i + 1, data[done + i].data(), 12);
            }
            put(packed.data(), faceBytes * n, faceOffset + faceBytes * (first + done));
        }
    }
    void close()
    {
        const int rc = ::close(fd);
        fd = -1;
        if (rc != 0) throw std::runtime_error("error writing " + path);
    }

private:
    static const std::size_t faceBytes = 13;   // uint8 count + 3 x uint32
    void put(const void *data, std::size_t bytes, std::size_t offset)
    {
        const char *p = static_cast<const char *>(data);
        while (bytes > 0)
        {
            const ssize_t n = ::pwrite(fd, p, bytes, (off_t) offset);
            if (n < 0 && errno == EINTR) continue;
            if (n <= 0) throw std::runtime_error("error writing " + path);
            p += n;
            offset += (std::size_t) n;
            bytes -= (std::size_t) n;
        }
    }
    std::string path;
    std::size_t vertexOffset, faceOffset;
    int fd = -1;
};

class PlyWriter
{
public:
    void addComment(const std::string &c) { comments.push_back(c); }

    /// The same file as write() would make for these counts, to be filled piece by piece.
    std::unique_ptr<PlyStreamWriter> openStream(const std::string &path, std::size_t numVertices, std::size_t numTriangles) const
    {
        return std::unique_ptr<PlyStreamWriter>(new PlyStreamWriter(path, makeHeader(numVertices, numTriangles), numVertices, numTriangles));
    }

    typedef std::array<float, 3> vertex_type;
    typedef std::array<std::uint32_t, 3> triangle_type;

    void write(const std::string &path, const std::vector<vertex_type> &vertices, const std::vector<triangle_type> &triangles) const
    {
        write(path, std::vector<std::vector<vertex_type>>(1, vertices), std::vector<std::vector<triangle_type>>(1, triangles));
    }

    /// Chunked mesh (Mesher::Result): the file is sized up front and mapped, and the chunks are
    /// packed into it in parallel; file systems that cannot map fall back to buffered writes.
    void write(const std::string &path, const std::vector<std::vector<vertex_type>> &vertexChunks,
               const std::vector<std::vector<triangle_type>> &triangleChunks) const
    {
        std::vector<std::size_t> vertexStart(vertexChunks.size() + 1, 0), triangleStart(triangleChunks.size() + 1, 0);
        for (std::size_t i = 0; i < vertexChunks.size(); i++) vertexStart[i + 1] = vertexStart[i] + vertexChunks[i].size();
        for (std::size_t i = 0; i < triangleChunks.size(); i++) triangleStart[i + 1] = triangleStart[i] + triangleChunks[i].size();
        const std::size_t numVertices = vertexStart.back(), numTriangles = triangleStart.back();
        const std::string header = makeHeader(numVertices, numTriangles);
        const std::size_t vertexOffset = header.size(), faceOffset = vertexOffset + 12 * numVertices;
        const std::size_t total = faceOffset + faceBytes * numTriangles;

        const int fd = ::open(path.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0666);
        if (fd < 0) throw std::runtime_error("cannot create " + path);
        void *map = MAP_FAILED;
        if (::ftruncate(fd, (off_t) total) == 0) map = ::mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        if (map != MAP_FAILED)
        {
            unsigned char *out = static_cast<unsigned char *>(map);
            std::memcpy(out, header.data(), header.size());
#pragma omp parallel for schedule(dynamic, 1)
            for (std::int64_t i = 0; i < (std::int64_t) vertexChunks.size(); i++)
                if (!vertexChunks[i].empty())
                    std::memcpy(out + vertexOffset + 12 * vertexStart[i], vertexChunks[i].data(), 12 * vertexChunks[i].size());
#pragma omp parallel for schedule(dynamic, 1)
            for (std::int64_t i = 0; i < (std::int64_t) triangleChunks.size(); i++)
                packFaces(triangleChunks[i].data(), triangleChunks[i].size(), out + faceOffset + faceBytes * triangleStart[i]);
            const bool bad = ::munmap(map, total) != 0;
            if (::close(fd) != 0 || bad) throw std::runtime_error("error writing " + path);
            return;
        }
        ::close(fd);
        std::FILE *f = std::fopen(path.c_str(), "wb");
        if (!f) throw std::runtime_error("cannot create " + path);
        std::fwrite(header.data(), 1, header.size(), f);
        for (const std::vector<vertex_type> &c : vertexChunks)
            if (!c.empty()) std::fwrite(c.data(), sizeof(c[0]), c.size(), f);
        std::vector<unsigned char> faces;
        for (const std::vector<triangle_type> &c : triangleChunks)
        {
            faces.resize(c.size() * faceBytes);
            packFaces(c.data(), c.size(), faces.data());
            if (!faces.empty()) std::fwrite(faces.data(), 1, faces.size(), f);
        }
        if (std::fclose(f) != 0) throw std::runtime_error("error writing " + path);
    }

    std::string makeHeader(std::size_t numVertices, std::size_t numTriangles) const
    {
        std::ostringstream out;
        out << "ply\nformat binary_little_endian 1.0\n";
        for (const std::string &c : comments) out << "comment " << c << '\n';
        out << "element vertex " << numVertices << '\n'
            << "property float32 x\nproperty float32 y\nproperty float32 z\n"
            << "element face " << numTriangles << '\n'
            << "property list uint8 uint32 vertex_indices\n"
            << "comment padding:";
        std::size_t size = (std::size_t) out.tellp() + 12;   // 12 for \nend_header\n
        while (size % 4 != 0)
        {
            out << 'X';
            size++;
        }
        out << "\nend_header\n";
        return out.str();
    }

private:
    static const std::size_t faceBytes = 13;   // uint8 count + 3 x uint32

    static void packFaces(const triangle_type *triangles, std::size_t count, unsigned char *out)
    {
        for (std::size_t i = 0; i < count; i++)
        {
            out[faceBytes * i] = 3;
            std::memcpy(out + faceBytes * i + 1, triangles[i].data(), 12);
        }
    }

    std::vector<std::string> comments;
};

}  // namespace mlsb

#endif
