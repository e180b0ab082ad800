// read_mesh: the file readers trimem reaches through OpenMesh::IO::read_mesh (src/mesh.cpp:7-16):
// STL (binary and ASCII; the default input `<conf>.stl` of mc/config.py), OFF, OBJ and PLY (ascii / binary, both
// byte orders). Host-only.
//
// What matters for parity is the numbering the reader produces, because every id the Python
// side sees (vertex rows of points(), edge ids of the flip sweeps) follows from it:
//   * STL has no shared vertices: OpenMesh's STL reader welds corners with bitwise-equal float32
//     coordinates (a std::map keyed on the coordinate triple), numbering vertices in order of first
//     appearance, and skips facets that degenerate under the welding; coordinates are float32
//     widened to double (the docs example doc/python-module.rst:60-98 depends on exactly this);
//   * OFF / OBJ / PLY list vertices explicitly: ids are file order, faces in file order; polygons are
//     fan-triangulated from their first corner (TriMesh kernels triangulate on add_face).
// Faces then go through HalfedgeMesh::add_face in file order, i.e. the same construction as
// TriMesh(points, cells).
#include <array>
#include <cctype>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "halfedge_mesh.h"

namespace trimem_b200 {

namespace {

struct Raw {
    std::vector<double> xyz;
    std::vector<int> tri;
};

std::string lower_ext(const std::string& f) {
    const size_t p = f.find_last_of('.');
    std::string e = p == std::string::npos ? "" : f.substr(p + 1);
    for (auto& c : e) c = (char)std::tolower((unsigned char)c);
    return e;
}

struct Welder {
    std::map<std::array<float, 3>, int> ids;
    Raw* raw;
    int vertex(const float c[3]) {
        const std::array<float, 3> key = {c[0], c[1], c[2]};
        auto it = ids.find(key);
        if (it != ids.end()) return it->second;
        const int id = (int)(raw->xyz.size() / 3);
        ids.emplace(key, id);
        for (int k = 0; k < 3; k++) raw->xyz.push_back((double)c[k]);
        return id;
    }
    void facet(const float c[9]) {
        const int a = vertex(c), b = vertex(c + 3), d = vertex(c + 6);
        if (a == b || b == d || a == d) return;           // degenerate after welding: skipped
        raw->tri.push_back(a); raw->tri.push_back(b); raw->tri.push_back(d);
    }
};

bool read_stl(const std::string& fname, Raw& raw) {
    std::ifstream in(fname, std::ios::binary);
    if (!in) return false;
    in.seekg(0, std::ios::end);
    const long long size = (long long)in.tellg();
    in.seekg(0);
    Welder w;
    w.raw = &raw;
    // binary iff the size matches the facet count in the header (OpenMesh's check_stl_type)
    bool binary = false;
    uint32_t nf = 0;
    if (size >= 84) {
        char head[80];
        in.read(head, 80);
        in.read(reinterpret_cast<char*>(&nf), 4);
        binary = size == 84 + 50LL * (long long)nf;
    }
    if (binary) {
        std::vector<char> buf(50 * (size_t)nf);
        in.read(buf.data(), (std::streamsize)buf.size());
        if ((size_t)in.gcount() != buf.size()) return false;
        for (uint32_t f = 0; f < nf; f++) {
            float c[9];
            std::memcpy(c, buf.data() + 50 * (size_t)f + 12, 36);     // after the 12-byte normal
            w.facet(c);
        }
        return true;
    }
    in.clear();
    in.seekg(0);
    std::string tok;
    float c[9];
    int nv = 0;
    bool any = false;
    while (in >> tok) {
        if (tok == "vertex") {
            if (nv >= 3) return false;
            if (!(in >> c[3 * nv] >> c[3 * nv + 1] >> c[3 * nv + 2])) return false;
            ++nv;
        } else if (tok == "endfacet") {
            if (nv != 3) return false;
            w.facet(c);
            nv = 0;
            any = true;
        }
    }
    return any;
}

void fan(const std::vector<int>& poly, Raw& raw) {
    for (size_t k = 1; k + 1 < poly.size(); k++) {
        raw.tri.push_back(poly[0]); raw.tri.push_back(poly[k]); raw.tri.push_back(poly[k + 1]);
    }
}

bool next_data_line(std::istream& in, std::string& line) {
    while (std::getline(in, line)) {
        const size_t h = line.find('#');
        if (h != std::string::npos) line.erase(h);
        if (line.find_first_not_of(" \t\r\n") != std::string::npos) return true;
    }
    return false;
}

bool read_off(const std::string& fname, Raw& raw) {
    std::ifstream in(fname);
    if (!in) return false;
    std::string line;
    if (!next_data_line(in, line)) return false;
    std::istringstream hs(line);
    std::string magic;
    hs >> magic;
    if (magic.find("OFF") == std::string::npos) return false;
    long long nv = -1, nf = -1, ne = 0;
    if (!(hs >> nv >> nf >> ne)) {
        if (!next_data_line(in, line)) return false;
        std::istringstream cs(line);
        if (!(cs >> nv >> nf)) return false;
    }
    if (nv < 0 || nf < 0) return false;
    for (long long v = 0; v < nv; v++) {
        if (!next_data_line(in, line)) return false;
        std::istringstream vs(line);
        double x, y, z;
        if (!(vs >> x >> y >> z)) return false;
        raw.xyz.push_back(x); raw.xyz.push_back(y); raw.xyz.push_back(z);
    }
    for (long long f = 0; f < nf; f++) {
        if (!next_data_line(in, line)) return false;
        std::istringstream fs(line);
        int n;
        if (!(fs >> n) || n < 3) return false;
        std::vector<int> poly(n);
        for (int k = 0; k < n; k++) if (!(fs >> poly[k]) || poly[k] < 0 || poly[k] >= nv) return false;
        fan(poly, raw);
    }
    return true;
}

bool read_obj(const std::string& fname, Raw& raw) {
    std::ifstream in(fname);
    if (!in) return false;
    std::string line;
    while (next_data_line(in, line)) {
        std::istringstream ls(line);
        std::string tag;
        ls >> tag;
        if (tag == "v") {
            double x, y, z;
            if (!(ls >> x >> y >> z)) return false;
            raw.xyz.push_back(x); raw.xyz.push_back(y); raw.xyz.push_back(z);
        } else if (tag == "f") {
            std::vector<int> poly;
            std::string corner;
            const int nv = (int)(raw.xyz.size() / 3);
            while (ls >> corner) {
                const int id = std::atoi(corner.substr(0, corner.find('/')).c_str());   // v, v/vt, v/vt/vn, v//vn
                const int v = id > 0 ? id - 1 : nv + id;                                // negative = relative
                if (id == 0 || v < 0 || v >= nv) return false;
                poly.push_back(v);
            }
            if (poly.size() < 3) return false;
            fan(poly, raw);
        }
    }
    return !raw.xyz.empty();
}

// PLY (ascii, binary_little_endian, binary_big_endian): any element / property layout; the vertex element's
// x, y, z (float32 widened to double, or float64 as is: OpenMesh's _PLYReader_ reads into the declared type) and the
// face element's list property (vertex_indices / vertex_index); every other property is parsed and skipped.
struct PlyProp {
    std::string name;
    int type = 0, count_type = 0;     // index into PLY_TYPES; count_type > 0 for list properties
    bool is_list = false;
};
struct PlyElem {
    std::string name;
    long long count = 0;
    std::vector<PlyProp> props;
};
const char* const PLY_TYPES[] = {"", "char", "uchar", "short", "ushort", "int", "uint", "float", "double"};
const int PLY_SIZE[] = {0, 1, 1, 2, 2, 4, 4, 4, 8};

int ply_type(const std::string& t) {
    static const std::map<std::string, int> alias = {
        {"char", 1}, {"int8", 1}, {"uchar", 2}, {"uint8", 2}, {"short", 3}, {"int16", 3}, {"ushort", 4}, {"uint16", 4},
        {"int", 5}, {"int32", 5}, {"uint", 6}, {"uint32", 6}, {"float", 7}, {"float32", 7}, {"double", 8}, {"float64", 8}};
    auto it = alias.find(t);
    return it == alias.end() ? 0 : it->second;
}

// one scalar of the given type as a double (exact for every PLY type), from a binary stream
bool ply_scalar_bin(std::istream& in, int type, bool swap, double& out) {
    unsigned char b[8];
    const int n = PLY_SIZE[type];
    in.read(reinterpret_cast<char*>(b), n);
    if (in.gcount() != n) return false;
    if (swap) for (int k = 0; k < n / 2; k++) std::swap(b[k], b[n - 1 - k]);
    switch (type) {
        case 1: { int8_t v; std::memcpy(&v, b, 1); out = v; break; }
        case 2: { uint8_t v; std::memcpy(&v, b, 1); out = v; break; }
        case 3: { int16_t v; std::memcpy(&v, b, 2); out = v; break; }
        case 4: { uint16_t v; std::memcpy(&v, b, 2); out = v; break; }
        case 5: { int32_t v; std::memcpy(&v, b, 4); out = v; break; }
        case 6: { uint32_t v; std::memcpy(&v, b, 4); out = v; break; }
        case 7: { float v; std::memcpy(&v, b, 4); out = (double)v; break; }
        case 8: { double v; std::memcpy(&v, b, 8); out = v; break; }
        default: return false;
    }
    return true;
}

bool ply_scalar_ascii(std::istream& in, int type, double& out) {
    if (type == 7) { float v; if (!(in >> v)) return false; out = (double)v; return true; }
    if (type == 8) return (bool)(in >> out);
    long long v;
    if (!(in >> v)) return false;
    out = (double)v;
    return true;
}

bool read_ply(const std::string& fname, Raw& raw) {
    std::ifstream in(fname, std::ios::binary);
    if (!in) return false;
    std::string line;
    if (!std::getline(in, line) || line.compare(0, 3, "ply") != 0) return false;
    int format = -1;                   // 0 ascii, 1 little endian, 2 big endian
    std::vector<PlyElem> elems;
    bool ended = false;
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        std::istringstream ls(line);
        std::string tag;
        ls >> tag;
        if (tag == "format") {
            std::string f;
            ls >> f;
            format = f == "ascii" ? 0 : f == "binary_little_endian" ? 1 : f == "binary_big_endian" ? 2 : -1;
        } else if (tag == "element") {
            PlyElem e;
            if (!(ls >> e.name >> e.count) || e.count < 0) return false;
            elems.push_back(e);
        } else if (tag == "property") {
            if (elems.empty()) return false;
            PlyProp p;
            std::string t;
            ls >> t;
            if (t == "list") {
                std::string ct, it;
                ls >> ct >> it >> p.name;
                p.is_list = true;
                p.count_type = ply_type(ct);
                p.type = ply_type(it);
                if (!p.count_type) return false;
            } else {
                p.type = ply_type(t);
                ls >> p.name;
            }
            if (!p.type || p.name.empty()) return false;
            elems.back().props.push_back(p);
        } else if (tag == "end_header") {
            ended = true;
            break;
        }                               // comment, obj_info: skipped
    }
    if (!ended || format < 0) return false;
    const uint16_t probe = 1;
    const bool host_little = *reinterpret_cast<const unsigned char*>(&probe) == 1;
    const bool swap = (format == 1 && !host_little) || (format == 2 && host_little);
    auto scalar = [&](int type, double& out) {
        return format == 0 ? ply_scalar_ascii(in, type, out) : ply_scalar_bin(in, type, swap, out);
    };
    long long nv = 0;
    std::vector<int> poly;
    for (const PlyElem& e : elems) {
        const bool is_vertex = e.name == "vertex", is_face = e.name == "face";
        for (long long r = 0; r < e.count; r++) {
            double xyz[3] = {0.0, 0.0, 0.0};
            for (const PlyProp& p : e.props) {
                if (p.is_list) {
                    double cnt;
                    if (!scalar(p.count_type, cnt) || cnt < 0) return false;
                    const bool take = is_face && (p.name == "vertex_indices" || p.name == "vertex_index");
                    if (take) poly.clear();
                    for (long long k = 0; k < (long long)cnt; k++) {
                        double v;
                        if (!scalar(p.type, v)) return false;
                        if (take) {
                            if (v < 0 || v >= (double)nv) return false;
                            poly.push_back((int)v);
                        }
                    }
                    if (take) {
                        if (poly.size() < 3) return false;
                        fan(poly, raw);
                    }
                } else {
                    double v;
                    if (!scalar(p.type, v)) return false;
                    if (is_vertex) {
                        if (p.name == "x") xyz[0] = v;
                        else if (p.name == "y") xyz[1] = v;
                        else if (p.name == "z") xyz[2] = v;
                    }
                }
            }
            if (is_vertex) { raw.xyz.push_back(xyz[0]); raw.xyz.push_back(xyz[1]); raw.xyz.push_back(xyz[2]); }
        }
        if (is_vertex) nv = e.count;
    }
    return nv > 0;
}

}  // namespace

// false on failure (the caller reports it); on success `mesh` holds the file's triangulation
bool read_mesh_file(const std::string& fname, HalfedgeMesh& mesh) {
    Raw raw;
    const std::string ext = lower_ext(fname);
    bool ok = false;
    if (ext == "stl" || ext == "stla" || ext == "stlb") ok = read_stl(fname, raw);
    else if (ext == "off") ok = read_off(fname, raw);
    else if (ext == "obj") ok = read_obj(fname, raw);
    else if (ext == "ply") ok = read_ply(fname, raw);
    if (!ok) return false;
    mesh.set_points(raw.xyz.data(), (int)(raw.xyz.size() / 3));
    for (size_t f = 0; f + 2 < raw.tri.size(); f += 3) mesh.add_face(raw.tri[f], raw.tri[f + 1], raw.tri[f + 2]);
    return true;
}

}  // namespace trimem_b200
