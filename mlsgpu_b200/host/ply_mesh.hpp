// General PLY mesh reader for the validation tools (plymanifold, meshdiff): the job of the
// reference's generic PLY::Reader with its vertex and triangle builders (extras/plymanifold.cpp:30-133).
//
// Accepts ascii and binary_little_endian files whose "vertex" element has scalar x, y, z (any
// numeric type, among other properties) and whose "face" element has a list property
// vertex_indices of integral type; other elements and properties are skipped.  Error wording
// follows the builders' validateProperties.
#ifndef MLSB200_PLY_MESH_HPP
#define MLSB200_PLY_MESH_HPP

#include <array>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace mlsb
{

struct PlyMesh
{
    std::vector<std::array<float, 3>> vertices;
    std::vector<std::array<std::uint32_t, 3>> triangles;
    std::vector<std::string> comments;
};

class PlyFormatError : public std::runtime_error
{
public:
    explicit PlyFormatError(const std::string &msg) : std::runtime_error(msg) {}
};

namespace plydetail
{
enum Type { INT8, UINT8, INT16, UINT16, INT32, UINT32, FLOAT32, FLOAT64 };

inline Type parseType(const std::string &t)
{
    static const struct { const char *name; Type type; } names[] = {
        {"char", INT8}, {"int8", INT8}, {"uchar", UINT8}, {"uint8", UINT8}, {"short", INT16}, {"int16", INT16},
        {"ushort", UINT16}, {"uint16", UINT16}, {"int", INT32}, {"int32", INT32}, {"uint", UINT32}, {"uint32", UINT32},
        {"float", FLOAT32}, {"float32", FLOAT32}, {"double", FLOAT64}, {"float64", FLOAT64}};
    for (const auto &n : names)
        if (t == n.name) return n.type;
    throw PlyFormatError("Unknown type " + t);
}
inline std::size_t typeSize(Type t)
{
    static const std::size_t sizes[] = {1, 1, 2, 2, 4, 4, 4, 8};
    return sizes[t];
}
inline bool isIntegral(Type t) { return t <= UINT32; }

struct Property
{
    std::string name;
    bool isList = false;
    Type lengthType = UINT8, valueType = FLOAT32;
};
struct Element
{
    std::string name;
    std::size_t count = 0;
    std::vector<Property> properties;
};

// one value from a binary little-endian stream, as double (exact for every PLY type but huge
// float64, which no index or coordinate here needs)
inline double readBinary(const unsigned char *&p, const unsigned char *end, Type t)
{
    const std::size_t n = typeSize(t);
    if ((std::size_t) (end - p) < n) throw PlyFormatError("Unexpected end of file");
    double v = 0.0;
    switch (t)
    {
    case INT8: { std::int8_t x; std::memcpy(&x, p, 1); v = x; break; }
    case UINT8: { std::uint8_t x; std::memcpy(&x, p, 1); v = x; break; }
    case INT16: { std::int16_t x; std::memcpy(&x, p, 2); v = x; break; }
    case UINT16: { std::uint16_t x; std::memcpy(&x, p, 2); v = x; break; }
    case INT32: { std::int32_t x; std::memcpy(&x, p, 4); v = x; break; }
    case UINT32: { std::uint32_t x; std::memcpy(&x, p, 4); v = x; break; }
    case FLOAT32: { float x; std::memcpy(&x, p, 4); v = x; break; }
    case FLOAT64: { std::memcpy(&v, p, 8); break; }
    }
    p += n;
    return v;
}
}  // namespace plydetail

inline PlyMesh readPlyMesh(const std::string &path)
{
    using namespace plydetail;
    std::ifstream in(path.c_str(), std::ios::in | std::ios::binary);
    if (!in) throw std::runtime_error("Could not open " + path);
    std::string line;
    auto getline = [&]() {
        if (!std::getline(in, line)) return false;
        if (!line.empty() && line[line.size() - 1] == '\r') line.erase(line.size() - 1);
        return true;
    };
    if (!getline() || line != "ply") throw PlyFormatError("PLY signature missing");
    PlyMesh mesh;
    std::vector<Element> elements;
    bool ascii = false, haveFormat = false, ended = false;
    while (getline())
    {
        std::istringstream tokens(line);
        std::string tok;
        tokens >> tok;
        if (tok == "end_header")
        {
            ended = true;
            break;
        }
        else if (tok == "format")
        {
            std::string fmt, version;
            tokens >> fmt >> version;
            if (fmt == "ascii") ascii = true;
            else if (fmt != "binary_little_endian") throw PlyFormatError("Unsupported format " + fmt);
            haveFormat = true;
        }
        else if (tok == "comment")
            mesh.comments.push_back(line.size() > 8 ? line.substr(8) : std::string());
        else if (tok == "element")
        {
            Element e;
            if (!(tokens >> e.name >> e.count)) throw PlyFormatError("Malformed element line");
            elements.push_back(e);
        }
        else if (tok == "property")
        {
            if (elements.empty()) throw PlyFormatError("Property without element");
            Property p;
            std::string type;
            tokens >> type;
            if (type == "list")
            {
                std::string lengthType, valueType;
                tokens >> lengthType >> valueType;
                p.isList = true;
                p.lengthType = parseType(lengthType);
                p.valueType = parseType(valueType);
                if (!isIntegral(p.lengthType)) throw PlyFormatError("List length must have integral type");
            }
            else
                p.valueType = parseType(type);
            if (!(tokens >> p.name)) throw PlyFormatError("Malformed property line");
            elements.back().properties.push_back(p);
        }
    }
    if (!ended || !haveFormat) throw PlyFormatError("Incomplete header");

    // ---- validateProperties of the two builders ----
    const Element *vertexElement = nullptr, *faceElement = nullptr;
    for (const Element &e : elements)
    {
        if (e.name == "vertex" && !vertexElement) vertexElement = &e;
        if (e.name == "face" && !faceElement) faceElement = &e;
    }
    if (!vertexElement) throw PlyFormatError("Missing element vertex");
    if (!faceElement) throw PlyFormatError("Missing element face");
    for (const char *field : {"x", "y", "z"})
    {
        const Property *found = nullptr;
        for (const Property &p : vertexElement->properties)
            if (p.name == field) found = &p;
        if (!found) throw PlyFormatError(std::string("Missing property ") + field);
        if (found->isList) throw PlyFormatError(std::string("Property ") + field + " should not be a list");
    }
    {
        const Property *found = nullptr;
        for (const Property &p : faceElement->properties)
            if (p.name == "vertex_indices") found = &p;
        if (!found) throw PlyFormatError("Missing property vertex_indices");
        if (!found->isList) throw PlyFormatError("Property vertex_indices should be a list");
        if (!isIntegral(found->valueType)) throw PlyFormatError("Property vertex_indices should have integral type");
    }

    // ---- body ----
    std::vector<unsigned char> body;
    const unsigned char *cursor = nullptr, *end = nullptr;
    if (!ascii)
    {
        const std::streampos at = in.tellg();
        in.seekg(0, std::ios::end);
        const std::streampos fileEnd = in.tellg();
        in.seekg(at);
        body.resize((std::size_t) (fileEnd - at));
        if (!body.empty()) in.read(reinterpret_cast<char *>(body.data()), (std::streamsize) body.size());
        if (!in) throw PlyFormatError("Unexpected end of file");
        cursor = body.data();
        end = cursor + body.size();
    }
    auto value = [&](Type t) -> double {
        if (!ascii) return readBinary(cursor, end, t);
        double v;
        if (!(in >> v)) throw PlyFormatError("Unexpected end of file");
        return v;
    };
    for (const Element &e : elements)
    {
        const bool isVertex = &e == vertexElement, isFace = &e == faceElement;
        if (isVertex) mesh.vertices.resize(e.count);
        if (isFace) mesh.triangles.resize(e.count);
        for (std::size_t i = 0; i < e.count; i++)
            for (const Property &p : e.properties)
            {
                if (!p.isList)
                {
                    const double v = value(p.valueType);
                    if (isVertex && p.name.size() == 1 && p.name[0] >= 'x' && p.name[0] <= 'z')
                        mesh.vertices[i][p.name[0] - 'x'] = (float) v;
                    continue;
                }
                const double length = value(p.lengthType);
                if (length < 0) throw PlyFormatError("Negative list length");
                const bool indices = isFace && p.name == "vertex_indices";
                if (indices && length != 3) throw PlyFormatError("Face does not contain 3 vertices");
                for (std::size_t k = 0; k < (std::size_t) length; k++)
                {
                    const double v = value(p.valueType);
                    if (indices)
                    {
                        if (!(v >= 0 && v <= 4294967295.0)) throw PlyFormatError("Negative or out-of-range index");
                        mesh.triangles[i][k] = (std::uint32_t) v;
                    }
                }
            }
    }
    return mesh;
}

}  // namespace mlsb

#endif
