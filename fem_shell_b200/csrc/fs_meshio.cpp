// fs_meshio.cpp -- host-side input formats of the path: in-memory meshGen, XDA and load files.
//
// Reference anchors:
//   meshGen            src/meshgen/main_all.cpp:133-387
//   XDA reader         fs.cpp:37 (libMesh XdrIO, ASCII "libMesh-0.7.0+" layout written at main_all.cpp:233-339)
//   XDR / MSH readers  fs.cpp:37,45-48: mesh.read() picks the format from the extension; the reference ships no
//                      fixture of either, so these two follow the format descriptions only (see the comments there)
//   load file reader   fs.cpp:44-67
// The generator reproduces the text round trip the reference imposes on its own output: node
// coordinates and the load factor pass through operator<< of a default std::ostream (6 significant
// digits, "%g") before fem-shell reads them back (main_all.cpp:261,351,373).
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/femshell_b200.h"

namespace {

double g6(double v)
{
    char buf[64];
    snprintf(buf, sizeof buf, "%g", v);
    return strtod(buf, nullptr);
}

// ---------------------------------------------------------------------------------------------
// Whole-file text cursor.  The 96 M-DOF mesh of BASELINE config 3 is a 1.3 GB XDA file; getline + stringstream
// parse it at ~45 MB/s, this cursor at several hundred (one read, no per-line allocation, hand-rolled integers).
// ---------------------------------------------------------------------------------------------
class TextFile {
public:
    ~TextFile() { free(buf_); }
    bool load(const char *path)
    {
        FILE *f = fopen(path, "rb");
        if (!f) return false;
        bool ok = fseek(f, 0, SEEK_END) == 0;
        const long long size = ok ? ftell(f) : -1;
        ok = ok && size >= 0 && fseek(f, 0, SEEK_SET) == 0;
        if (ok) {
            buf_ = (char *)malloc((size_t)size + 1);
            ok = buf_ && fread(buf_, 1, (size_t)size, f) == (size_t)size;
        }
        fclose(f);
        if (!ok) return false;
        buf_[size] = '\0';  // strtod may look one character past the last token
        cur_ = buf_;
        end_ = buf_ + size;
        return true;
    }
    const char *begin() const { return buf_; }
    const char *end() const { return end_; }
    // the next line as it stands (std::getline)
    bool raw_line(const char *&b, const char *&e)
    {
        if (cur_ >= end_) return false;
        b = cur_;
        const char *nl = (const char *)memchr(cur_, '\n', (size_t)(end_ - cur_));
        e = nl ? nl : end_;
        cur_ = nl ? nl + 1 : end_;
        return true;
    }
    // the next line that holds something besides blanks once its "# comment" tail is cut off
    bool next_line(const char *&b, const char *&e)
    {
        while (raw_line(b, e)) {
            const char *h = (const char *)memchr(b, '#', (size_t)(e - b));
            if (h) e = h;
            for (const char *p = b; p < e; p++)
                if (*p != ' ' && *p != '\t' && *p != '\r') return true;
        }
        return false;
    }

private:
    char *buf_ = nullptr;
    const char *cur_ = nullptr, *end_ = nullptr;
};

inline bool is_blank(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\v' || c == '\f'; }

// next whitespace-separated token of [b, e) as an integer; advances b past it
bool parse_int(const char *&b, const char *e, long long &out)
{
    while (b < e && is_blank(*b)) b++;
    if (b >= e) return false;
    bool neg = false;
    if (*b == '-' || *b == '+') neg = *b++ == '-';
    if (b >= e || *b < '0' || *b > '9') return false;
    long long v = 0;
    while (b < e && *b >= '0' && *b <= '9') v = v * 10 + (*b++ - '0');
    out = neg ? -v : v;
    return true;
}

// next token as a double (strtod: the buffer is NUL-terminated and a number never spans a line break)
bool parse_double(const char *&b, const char *e, double &out)
{
    while (b < e && is_blank(*b)) b++;
    if (b >= e) return false;
    char *stop = nullptr;
    const double v = strtod(b, &stop);
    if (stop == b || stop > e) return false;
    b = stop;
    out = v;
    return true;
}


// ---------------------------------------------------------------------------------------------
// Gmsh MSH 2.x ASCII as libMesh's GmshIO reads it (fs.cpp:37 with a *.msh name).  Nodes are renumbered in file order
// (Gmsh ids are 1-based and may have gaps); the elements of the highest dimension present -- here 3-node triangles
// (Gmsh type 2) and 4-node quadrangles (type 3), node order unchanged -- become the mesh elements in file order;
// 2-node lines (type 1) are boundary descriptions: their first tag (the physical group) becomes the boundary id of
// every element side with the same two nodes.  Points (type 15) are skipped; anything else is refused like an
// unsupported element in an XDA file.  No reference fixture exists: pinned by tests/test_meshio_formats.py only.
// ---------------------------------------------------------------------------------------------
int read_msh(const char *path, int64_t *n_nodes_o, int64_t *n_elem_o, int64_t *n_enodes_o, int64_t *n_bc_o, double *xyz,
             int32_t *etype, int64_t *eptr, int32_t *enodes, int32_t *bc)
{
    TextFile f;
    if (!f.load(path)) return FS_ERR_IO;
    const char *b, *e;
    auto is_tag = [&](const char *tag) {
        const size_t n = strlen(tag);
        const char *q = b;
        while (q < e && is_blank(*q)) q++;
        return (size_t)(e - q) >= n && memcmp(q, tag, n) == 0;
    };
    std::unordered_map<long long, int32_t> node_of;     // Gmsh node id -> consecutive id
    std::vector<double> X;
    std::vector<int32_t> T, EN;                          // element types / nodes (consecutive ids)
    std::vector<int64_t> EP(1, 0);
    struct Line { int32_t a, b, id; };
    std::vector<Line> lines;
    bool have_nodes = false, have_elems = false;
    while (f.raw_line(b, e)) {
        if (is_tag("$MeshFormat")) {
            if (!f.raw_line(b, e)) return FS_ERR_IO;
            double ver = 0.0;
            long long ftype = 0;
            if (!parse_double(b, e, ver) || !parse_int(b, e, ftype)) return FS_ERR_IO;
            if (ver < 2.0 || ver >= 3.0 || ftype != 0) return FS_ERR_ARG;   // MSH 2.x ASCII only
        } else if (is_tag("$Nodes")) {
            long long n = 0;
            if (!f.raw_line(b, e) || !parse_int(b, e, n) || n <= 0) return FS_ERR_IO;
            X.resize(3 * (size_t)n);
            node_of.reserve((size_t)n * 2);
            for (long long i = 0; i < n; i++) {
                long long id;
                if (!f.raw_line(b, e) || !parse_int(b, e, id)) return FS_ERR_IO;
                for (int k = 0; k < 3; k++)
                    if (!parse_double(b, e, X[3 * (size_t)i + k])) return FS_ERR_IO;
                if (!node_of.emplace(id, (int32_t)i).second) return FS_ERR_IO;   // duplicate node id
            }
            have_nodes = true;
        } else if (is_tag("$Elements")) {
            if (!have_nodes) return FS_ERR_IO;
            long long n = 0;
            if (!f.raw_line(b, e) || !parse_int(b, e, n) || n <= 0) return FS_ERR_IO;
            for (long long i = 0; i < n; i++) {
                long long id, type, ntags, tag0 = 0, v;
                if (!f.raw_line(b, e) || !parse_int(b, e, id) || !parse_int(b, e, type) || !parse_int(b, e, ntags)) return FS_ERR_IO;
                for (long long k = 0; k < ntags; k++) {
                    if (!parse_int(b, e, v)) return FS_ERR_IO;
                    if (k == 0) tag0 = v;
                }
                const int nen = type == 1 ? 2 : (type == 2 ? 3 : (type == 3 ? 4 : (type == 15 ? 1 : 0)));
                if (!nen) return FS_ERR_ARG;   // only what fem-shell handles (fs.cpp:315,342)
                int32_t nd[4];
                for (int k = 0; k < nen; k++) {
                    if (!parse_int(b, e, v)) return FS_ERR_IO;
                    auto it = node_of.find(v);
                    if (it == node_of.end()) return FS_ERR_IO;
                    nd[k] = it->second;
                }
                if (type == 1) lines.push_back({nd[0], nd[1], (int32_t)tag0});
                else if (type == 2 || type == 3) {
                    T.push_back(type == 2 ? FS_TRI3 : FS_QUAD4);
                    EN.insert(EN.end(), nd, nd + nen);
                    EP.push_back((int64_t)EN.size());
                }
            }
            have_elems = true;
        }
    }
    if (!have_nodes || !have_elems || T.empty()) return FS_ERR_IO;
    // sides of the mesh elements by their (unordered) node pair; a line tags every side it coincides with
    auto key = [](int32_t a, int32_t c) { return ((long long)std::min(a, c) << 32) | (unsigned int)std::max(a, c); };
    std::unordered_multimap<long long, std::pair<int32_t, int32_t>> side_of;
    if (!lines.empty()) {
        side_of.reserve(EN.size());
        for (size_t el = 0; el < T.size(); el++) {
            const int nen = (int)(EP[el + 1] - EP[el]);
            for (int sd = 0; sd < nen; sd++) side_of.emplace(key(EN[EP[el] + sd], EN[EP[el] + (sd + 1) % nen]), std::make_pair((int32_t)el, (int32_t)sd));
        }
    }
    std::vector<int32_t> B;
    for (const Line &ln : lines) {
        auto range = side_of.equal_range(key(ln.a, ln.b));
        std::vector<std::pair<int32_t, int32_t>> hits;
        for (auto it = range.first; it != range.second; ++it) hits.push_back(it->second);
        std::sort(hits.begin(), hits.end());   // the multimap's order is unspecified
        for (const auto &h : hits) { B.push_back(h.first); B.push_back(h.second); B.push_back(ln.id); }
    }
    if (n_nodes_o) *n_nodes_o = (int64_t)(X.size() / 3);
    if (n_elem_o) *n_elem_o = (int64_t)T.size();
    if (n_enodes_o) *n_enodes_o = (int64_t)EN.size();
    if (n_bc_o) *n_bc_o = (int64_t)(B.size() / 3);
    if (xyz) memcpy(xyz, X.data(), sizeof(double) * X.size());
    if (etype) memcpy(etype, T.data(), sizeof(int32_t) * T.size());
    if (eptr) memcpy(eptr, EP.data(), sizeof(int64_t) * EP.size());
    if (enodes) memcpy(enodes, EN.data(), sizeof(int32_t) * EN.size());
    if (bc && !B.empty()) memcpy(bc, B.data(), sizeof(int32_t) * B.size());
    return FS_OK;
}

// ---------------------------------------------------------------------------------------------
// XDR: the binary twin of the XDA layout above -- the same fields in the same order through Sun XDR (RFC 4506) instead
// of text, no comments: strings as a 4-byte big-endian length + bytes padded to a multiple of four, ids and counts as
// 4-byte big-endian unsigned integers (libMesh's default 32-bit ids), coordinates as 8-byte big-endian IEEE doubles.
//   string "libMesh-0.7.0+" | n_elem | n_nodes | strings ".", "n/a", "n/a", "n/a" (bc / subdomain / processor / p-level
//   specification) | n_elem at level 0 | per element: type, nodes | per node: x y z | n_bc | per record: elem side id
// No reference fixture exists and libMesh is not in this image: the layout is the format description restated, pinned
// only by the round trip with fs_write_xdr (tests/test_meshio_formats.py).
// ---------------------------------------------------------------------------------------------
class XdrIn {
public:
    ~XdrIn() { free(buf_); }
    bool load(const char *path)
    {
        FILE *f = fopen(path, "rb");
        if (!f) return false;
        bool ok = fseek(f, 0, SEEK_END) == 0;
        const long long size = ok ? ftell(f) : -1;
        ok = ok && size >= 0 && fseek(f, 0, SEEK_SET) == 0;
        if (ok) {
            buf_ = (unsigned char *)malloc((size_t)size + 1);
            ok = buf_ && fread(buf_, 1, (size_t)size, f) == (size_t)size;
        }
        fclose(f);
        n_ = ok ? (size_t)size : 0;
        return ok;
    }
    bool u32(uint32_t &v)
    {
        if (at_ + 4 > n_) return false;
        const unsigned char *p = buf_ + at_;
        v = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
        at_ += 4;
        return true;
    }
    bool f64(double &v)
    {
        if (at_ + 8 > n_) return false;
        const unsigned char *p = buf_ + at_;
        uint64_t u = 0;
        for (int k = 0; k < 8; k++) u = (u << 8) | p[k];
        memcpy(&v, &u, 8);
        at_ += 8;
        return true;
    }
    bool str(std::string &out)
    {
        uint32_t len;
        if (!u32(len) || len > 4096 || at_ + ((len + 3u) & ~3u) > n_) return false;
        out.assign((const char *)buf_ + at_, len);
        at_ += (len + 3u) & ~3u;
        return true;
    }

private:
    unsigned char *buf_ = nullptr;
    size_t n_ = 0, at_ = 0;
};

int read_xdr(const char *path, int64_t *n_nodes_o, int64_t *n_elem_o, int64_t *n_enodes_o, int64_t *n_bc_o, double *xyz,
             int32_t *etype, int64_t *eptr, int32_t *enodes, int32_t *bc)
{
    XdrIn f;
    if (!f.load(path)) return FS_ERR_IO;
    std::string s;
    if (!f.str(s) || s.compare(0, 7, "libMesh") != 0) return FS_ERR_IO;
    uint32_t n_elem = 0, n_nodes = 0, lvl0 = 0, v = 0;
    if (!f.u32(n_elem) || !f.u32(n_nodes)) return FS_ERR_IO;
    for (int i = 0; i < 4; i++)
        if (!f.str(s)) return FS_ERR_IO;
    if (!f.u32(lvl0) || n_elem == 0 || n_nodes == 0 || lvl0 != n_elem) return FS_ERR_IO;
    int64_t n_en = 0;
    if (eptr) eptr[0] = 0;
    for (uint32_t el = 0; el < n_elem; el++) {
        uint32_t t;
        if (!f.u32(t)) return FS_ERR_IO;
        const int nen = (t == FS_TRI3) ? 3 : (t == FS_QUAD4 ? 4 : 0);
        if (!nen) return FS_ERR_ARG;
        for (int k = 0; k < nen; k++) {
            if (!f.u32(v) || v >= n_nodes) return FS_ERR_IO;
            if (enodes) enodes[n_en + k] = (int32_t)v;
        }
        if (etype) etype[el] = (int32_t)t;
        n_en += nen;
        if (eptr) eptr[el + 1] = n_en;
    }
    for (uint32_t i = 0; i < n_nodes; i++)
        for (int k = 0; k < 3; k++) {
            double c;
            if (!f.f64(c)) return FS_ERR_IO;
            if (xyz) xyz[3 * (size_t)i + k] = c;
        }
    uint32_t n_bc = 0;
    if (!f.u32(n_bc)) n_bc = 0;   // like the text reader: a file may end after the nodes
    for (uint32_t i = 0; i < n_bc; i++)
        for (int k = 0; k < 3; k++) {
            if (!f.u32(v)) return FS_ERR_IO;
            if (bc) bc[3 * (size_t)i + k] = (int32_t)v;
        }
    if (n_nodes_o) *n_nodes_o = n_nodes;
    if (n_elem_o) *n_elem_o = n_elem;
    if (n_enodes_o) *n_enodes_o = n_en;
    if (n_bc_o) *n_bc_o = n_bc;
    return FS_OK;
}

bool ends_with(const char *path, const char *ext)
{
    const size_t n = strlen(path), m = strlen(ext);
    return n >= m && strcmp(path + n - m, ext) == 0;
}

}  // namespace

extern "C" {

int fs_read_mesh(const char *path, int64_t *n_nodes, int64_t *n_elem, int64_t *n_enodes, int64_t *n_bc, double *xyz,
                 int32_t *etype, int64_t *eptr, int32_t *enodes, int32_t *bc)
{
    if (!path) return FS_ERR_ARG;
    if (ends_with(path, ".msh")) return read_msh(path, n_nodes, n_elem, n_enodes, n_bc, xyz, etype, eptr, enodes, bc);
    if (ends_with(path, ".xdr")) return read_xdr(path, n_nodes, n_elem, n_enodes, n_bc, xyz, etype, eptr, enodes, bc);
    return fs_read_xda(path, n_nodes, n_elem, n_enodes, n_bc, xyz, etype, eptr, enodes, bc);
}

int fs_write_xdr(const char *path, int64_t n_nodes, const double *xyz, int64_t n_elem, const int32_t *etype,
                 const int64_t *eptr, const int32_t *enodes, int64_t n_bc, const int32_t *bc)
{
    if (!path || !xyz || !etype || !eptr || !enodes || (n_bc > 0 && !bc)) return FS_ERR_ARG;
    FILE *f = fopen(path, "wb");
    if (!f) return FS_ERR_IO;
    std::vector<unsigned char> buf;
    buf.reserve(1 << 20);
    bool ok = true;
    auto flush = [&]() {
        ok = ok && fwrite(buf.data(), 1, buf.size(), f) == buf.size();
        buf.clear();
    };
    auto u32 = [&](uint32_t v) {
        if (buf.size() + 16 > buf.capacity()) flush();
        for (int k = 3; k >= 0; k--) buf.push_back((unsigned char)(v >> (8 * k)));
    };
    auto f64 = [&](double d) {
        if (buf.size() + 16 > buf.capacity()) flush();
        uint64_t u;
        memcpy(&u, &d, 8);
        for (int k = 7; k >= 0; k--) buf.push_back((unsigned char)(u >> (8 * k)));
    };
    auto str = [&](const char *s) {
        const uint32_t len = (uint32_t)strlen(s);
        u32(len);
        for (uint32_t k = 0; k < ((len + 3u) & ~3u); k++) buf.push_back(k < len ? (unsigned char)s[k] : 0);
    };
    str("libMesh-0.7.0+");
    u32((uint32_t)n_elem);
    u32((uint32_t)n_nodes);
    str(".");
    str("n/a");
    str("n/a");
    str("n/a");
    u32((uint32_t)n_elem);
    for (int64_t e = 0; e < n_elem; e++) {
        u32((uint32_t)etype[e]);
        for (int64_t k = eptr[e]; k < eptr[e + 1]; k++) u32((uint32_t)enodes[k]);
    }
    for (int64_t i = 0; i < 3 * n_nodes; i++) f64(xyz[i]);
    u32((uint32_t)n_bc);
    for (int64_t i = 0; i < 3 * n_bc; i++) u32((uint32_t)bc[i]);
    flush();
    ok = fclose(f) == 0 && ok;
    return ok ? FS_OK : FS_ERR_IO;
}

}  // extern "C"

extern "C" {

int fs_meshgen(char kind, int nx, int ny, double min_x, double min_y, double max_x, double max_y,
               const int bcids[4], double factor, int loading, int ul_lr, char dead_axis, int64_t *n_nodes_o,
               int64_t *n_elem_o, int64_t *n_bc_o, double *xyz, int32_t *etype, int64_t *eptr, int32_t *enodes,
               int32_t *bc, double *forces)
{
    if (kind == 'Q') kind = 'q';
    if (kind == 'T') kind = 't';
    if ((kind != 'q' && kind != 't') || nx <= 0 || ny <= 0 || !bcids) return FS_ERR_ARG;       // main_all.cpp:40-66
    if (dead_axis != 'x' && dead_axis != 'y' && dead_axis != 'z') return FS_ERR_ARG;           // main_all.cpp:127-131
    const int t_id = bcids[0], b_id = bcids[1], l_id = bcids[2], r_id = bcids[3];
    const int64_t n_nodes = (int64_t)(nx + 1) * (ny + 1);
    const int64_t n_elem = (int64_t)nx * ny * (kind == 't' ? 2 : 1);
    int64_t n_bc = 0;
    if (l_id >= 0) n_bc += ny;
    if (r_id >= 0) n_bc += ny;
    if (t_id >= 0) n_bc += nx;
    if (b_id >= 0) n_bc += nx;
    if (n_nodes_o) *n_nodes_o = n_nodes;
    if (n_elem_o) *n_elem_o = n_elem;
    if (n_bc_o) *n_bc_o = n_bc;
    const int nen = kind == 't' ? 3 : 4;
    const int64_t W = nx + 1;

    if (xyz) {  // main_all.cpp:141-160
        const double fracx = (max_x - min_x) / (double)nx, fracy = (max_y - min_y) / (double)ny;
        std::vector<double> xs(nx + 1), ys(ny + 1);
        for (int x = 0; x <= nx; x++) xs[x] = g6(min_x + x * fracx);
        for (int y = 0; y <= ny; y++) ys[y] = g6(min_y + y * fracy);
        const int pa = dead_axis == 'x' ? 1 : 0, sa = dead_axis == 'z' ? 1 : 2;
        for (int y = 0; y <= ny; y++)
            for (int x = 0; x <= nx; x++) {
                double *p = xyz + 3 * ((int64_t)y * W + x);
                p[0] = p[1] = p[2] = 0.0;
                p[sa] = ys[y];
                p[pa] = xs[x];
            }
    }
    if (etype && eptr && enodes) {  // main_all.cpp:163-224
        int64_t e = 0;
        eptr[0] = 0;
        for (int y = 0; y < ny; y++)
            for (int x = 0; x < nx; x++) {
                const int32_t n = (int32_t)(x + (int64_t)y * W);
                const int32_t A = n, B = n + 1, C = (int32_t)(n + W + 1), D = (int32_t)(n + W);
                if (kind == 'q') {
                    int32_t *q = enodes + 4 * e;
                    q[0] = A; q[1] = B; q[2] = C; q[3] = D;
                    etype[e] = FS_QUAD4;
                    eptr[e + 1] = 4 * (e + 1);
                    e++;
                } else {
                    int32_t *t1 = enodes + 3 * e, *t2 = t1 + 3;
                    if (ul_lr) {
                        t1[0] = A; t1[1] = B; t1[2] = D;
                        t2[0] = B; t2[1] = C; t2[2] = D;
                    } else {
                        t1[0] = A; t1[1] = C; t1[2] = B;
                        t2[0] = C; t2[1] = A; t2[2] = D;
                    }
                    etype[e] = etype[e + 1] = FS_TRI3;
                    eptr[e + 1] = 3 * (e + 1);
                    eptr[e + 2] = 3 * (e + 2);
                    e += 2;
                }
            }
    }
    if (bc) {  // main_all.cpp:284-338, same record order
        int64_t k = 0;
        auto put = [&](int64_t el, int side, int id) {
            bc[3 * k] = (int32_t)el; bc[3 * k + 1] = side; bc[3 * k + 2] = id;
            k++;
        };
        const int64_t nxy = (int64_t)nx * ny;
        for (int i = 0; i < nx; i++) {
            if (kind == 't') {
                if (b_id >= 0) put(2 * i, ul_lr ? 0 : 2, b_id);
                if (t_id >= 0) put(2 * nxy - 2 * i - 1, ul_lr ? 1 : 2, t_id);
            } else {
                if (b_id >= 0) put(i, 0, b_id);
                if (t_id >= 0) put(nxy - 1 - i, 2, t_id);
            }
        }
        for (int i = 0; i < ny; i++) {
            if (kind == 't') {
                if (ul_lr) {
                    if (l_id >= 0) put(2 * (int64_t)nx * i, 2, l_id);
                    if (r_id >= 0) put(2 * (int64_t)nx * (i + 1) - 1, 0, r_id);
                } else {
                    if (l_id >= 0) put(2 * (int64_t)nx * i + 1, 1, l_id);
                    if (r_id >= 0) put(2 * (int64_t)nx * (i + 1) - 2, 1, r_id);
                }
            } else {
                if (l_id >= 0) put((int64_t)nx * i, 3, l_id);
                if (r_id >= 0) put((int64_t)nx * (i + 1) - 1, 1, r_id);
            }
        }
    }
    if (forces) {  // main_all.cpp:343-387 followed by fs.cpp:52-66
        memset(forces, 0, sizeof(double) * 6 * n_nodes);
        const int comp = dead_axis == 'x' ? 0 : (dead_axis == 'y' ? 1 : 2);
        if (loading == 1) {
            const double f = g6(factor);
            if (n_nodes / 2 < n_nodes - 1) forces[6 * (n_nodes / 2) + comp] = 1.0 * f;
        } else if (loading == 2) {
            const double f = g6(factor * ((max_x - min_x) / (double)nx) * ((max_y - min_y) / (double)ny));
            for (int64_t i = 0; i < n_nodes - 1; i++) forces[6 * i + comp] = 1.0 * f;  // the last node gets no row
        }
    }
    (void)nen;
    return FS_OK;
}

int fs_read_xda(const char *path, int64_t *n_nodes_o, int64_t *n_elem_o, int64_t *n_enodes_o, int64_t *n_bc_o,
                double *xyz, int32_t *etype, int64_t *eptr, int32_t *enodes, int32_t *bc)
{
    if (!path) return FS_ERR_ARG;
    TextFile f;
    if (!f.load(path)) return FS_ERR_IO;
    const char *b, *e;
    if (!f.raw_line(b, e) || e - b < 7 || memcmp(b, "libMesh", 7) != 0) return FS_ERR_IO;
    long long n_elem = 0, n_nodes = 0, v = 0;
    if (!f.next_line(b, e) || !parse_int(b, e, n_elem)) return FS_ERR_IO;
    if (!f.next_line(b, e) || !parse_int(b, e, n_nodes)) return FS_ERR_IO;
    for (int i = 0; i < 4; i++)  // bc / subdomain / processor / p-level specification lines
        if (!f.raw_line(b, e)) return FS_ERR_IO;
    if (!f.next_line(b, e)) return FS_ERR_IO;  // n_elem at level 0
    if (n_elem <= 0 || n_nodes <= 0) return FS_ERR_IO;
    int64_t n_en = 0;
    if (eptr) eptr[0] = 0;
    for (int64_t el = 0; el < n_elem; el++) {
        if (!f.next_line(b, e)) return FS_ERR_IO;
        long long t;
        if (!parse_int(b, e, t)) return FS_ERR_IO;
        const int nen = (t == FS_TRI3) ? 3 : (t == FS_QUAD4 ? 4 : 0);
        if (!nen) return FS_ERR_ARG;  // only the two element types fem-shell handles (fs.cpp:315,342)
        for (int k = 0; k < nen; k++) {
            if (!parse_int(b, e, v)) return FS_ERR_IO;
            if (enodes) enodes[n_en + k] = (int32_t)v;
        }
        if (etype) etype[el] = (int32_t)t;
        n_en += nen;
        if (eptr) eptr[el + 1] = n_en;
    }
    for (int64_t i = 0; i < n_nodes; i++) {
        if (!f.next_line(b, e)) return FS_ERR_IO;
        double c[3];
        for (int k = 0; k < 3; k++)
            if (!parse_double(b, e, c[k])) return FS_ERR_IO;
        if (xyz) { xyz[3 * i] = c[0]; xyz[3 * i + 1] = c[1]; xyz[3 * i + 2] = c[2]; }
    }
    long long n_bc = 0;
    if (f.next_line(b, e)) parse_int(b, e, n_bc);
    for (int64_t i = 0; i < n_bc; i++) {
        if (!f.next_line(b, e)) return FS_ERR_IO;
        long long t[3];
        for (int k = 0; k < 3; k++)
            if (!parse_int(b, e, t[k])) return FS_ERR_IO;
        if (bc) { bc[3 * i] = (int32_t)t[0]; bc[3 * i + 1] = (int32_t)t[1]; bc[3 * i + 2] = (int32_t)t[2]; }
    }
    if (n_nodes_o) *n_nodes_o = n_nodes;
    if (n_elem_o) *n_elem_o = n_elem;
    if (n_enodes_o) *n_enodes_o = n_en;
    if (n_bc_o) *n_bc_o = n_bc;
    return FS_OK;
}

int fs_read_forces(const char *path, int64_t n_nodes, double *forces)
{
    if (!path || !forces || n_nodes <= 0) return FS_ERR_ARG;
    memset(forces, 0, sizeof(double) * 6 * n_nodes);
    TextFile f;
    if (!f.load(path)) return FS_ERR_IO;  // the reference silently runs without loads (fs.cpp:52); callers decide
    // fs.cpp:52-66 reads with operator>>: whitespace-separated tokens, no comment syntax; the first failed extraction
    // leaves this and every later zero-initialised DenseVector entry untouched
    const char *b = f.begin(), *e = f.end();
    long long n = 0;
    double factor = 1.0;
    if (!parse_int(b, e, n)) return FS_OK;
    if (!parse_double(b, e, factor)) return FS_OK;
    for (long long i = 0; i < n && i < n_nodes; i++)
        for (int j = 0; j < 6; j++) {
            double v;
            if (!parse_double(b, e, v)) return FS_OK;
            forces[6 * i + j] = v * factor;
        }
    return FS_OK;
}

int fs_write_xda(const char *path, int64_t n_nodes, const double *xyz, int64_t n_elem, const int32_t *etype,
                 const int64_t *eptr, const int32_t *enodes, int64_t n_bc, const int32_t *bc)
{
    if (!path || !xyz || !etype || !eptr || !enodes) return FS_ERR_ARG;
    FILE *f = fopen(path, "w");
    if (!f) return FS_ERR_IO;
    fprintf(f, "libMesh-0.7.0+\n%lld      # number of elements\n%lld      # number of nodes\n", (long long)n_elem, (long long)n_nodes);
    fprintf(f, ".        # boundary condition specification file\nn/a      # subdomain id specification file\n");
    fprintf(f, "n/a      # processor id specification file\nn/a      # p-level specification file\n");
    fprintf(f, "%lld      # n_elem at level 0, [ type (n0 ... nN-1) ]\n", (long long)n_elem);
    // element and boundary lines are integers only: formatted by hand into a block buffer (fprintf per value is the
    // bottleneck of a 32 M-element file); coordinates keep "%g", the reference generator's 6-digit text (main_all.cpp:261)
    std::vector<char> buf(1 << 20);
    size_t at = 0;
    bool ok = true;
    auto flush = [&]() {
        ok = ok && fwrite(buf.data(), 1, at, f) == at;
        at = 0;
    };
    auto put_int = [&](long long v) {
        char tmp[24];
        int n = 0;
        const bool neg = v < 0;
        unsigned long long u = neg ? 0ull - (unsigned long long)v : (unsigned long long)v;
        do tmp[n++] = (char)('0' + u % 10); while (u /= 10);
        if (neg) buf[at++] = '-';
        while (n) buf[at++] = tmp[--n];
    };
    for (int64_t e = 0; e < n_elem; e++) {
        if (at + 256 > buf.size()) flush();
        put_int(etype[e]);
        for (int64_t k = eptr[e]; k < eptr[e + 1]; k++) {
            if (at + 64 > buf.size()) flush();
            buf[at++] = ' ';
            put_int(enodes[k]);
        }
        buf[at++] = '\n';
    }
    for (int64_t i = 0; i < n_nodes; i++) {
        if (at + 256 > buf.size()) flush();
        at += (size_t)snprintf(buf.data() + at, 200, "%g %g %g\n", xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    }
    flush();
    fprintf(f, "%lld        # number of boundary conditions\n", (long long)n_bc);
    for (int64_t i = 0; i < n_bc; i++) {
        if (at + 256 > buf.size()) flush();
        for (int k = 0; k < 3; k++) {
            if (k) buf[at++] = ' ';
            put_int(bc[3 * i + k]);
        }
        buf[at++] = '\n';
    }
    flush();
    ok = fclose(f) == 0 && ok;
    return ok ? FS_OK : FS_ERR_IO;
}

}  // extern "C"
