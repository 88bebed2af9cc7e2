// ply_io.cpp — 3DGS `.ply` reader / writer (host side).
//
// Replaces import_ply / export_ply (src/gaussians.jl:157-247), which go through PlyIO.jl: scenes trained by the
// reference or by any other 3DGS implementation come in as the raw parameter arrays gsr_forward_raw consumes
// (pre-sigmoid opacity, log-scales, wxyz quaternion, SH dc | rest), in the reference's own memory layout:
//   points (3,N), features_dc (3,1,N), features_rest (3,R,N), opacities (1,N), scales (3,N), rotations (4,N).
// File layout (gaussians.jl:140-156): one `vertex` element, properties identified by NAME only (order and storage
// type are free): x y z, nx ny nz (ignored), f_dc_0..2, f_rest_0..3R-1 CHANNEL-major (all coefficients of R, then
// G, then B — the model stores the transpose), opacity, scale_0..2, rot_0..3.
// Formats: binary_little_endian, binary_big_endian and ascii; any PLY scalar type is converted to float32.
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/gsrast.h"

namespace {

struct Prop {
    std::string name;
    int type = 0;  // 0 char 1 uchar 2 short 3 ushort 4 int 5 uint 6 float 7 double
    size_t offset = 0;
};
const size_t kTypeSize[8] = {1, 1, 2, 2, 4, 4, 4, 8};

int parse_type(const std::string &t) {
    static const char *names[8][2] = {{"char", "int8"},   {"uchar", "uint8"}, {"short", "int16"},   {"ushort", "uint16"},
                                      {"int", "int32"},   {"uint", "uint32"}, {"float", "float32"}, {"double", "float64"}};
    for (int i = 0; i < 8; i++)
        if (t == names[i][0] || t == names[i][1]) return i;
    return -1;
}

struct Reader {
    FILE *f = nullptr;
    int format = 0;  // 0 ascii, 1 little endian, 2 big endian
    int64_t n = 0;
    std::vector<Prop> props;
    size_t row_bytes = 0;
    long data_start = 0;
    int n_rest = 0;  // number of f_rest_* properties
    std::string err;
};

thread_local std::string g_ply_error;

int fail(const std::string &m) {
    g_ply_error = m;
    return GSR_EINVAL;
}

double load_scalar(const unsigned char *p, int type, bool swap) {
    unsigned char b[8];
    const size_t sz = kTypeSize[type];
    if (swap)
        for (size_t i = 0; i < sz; i++) b[i] = p[sz - 1 - i];
    else
        memcpy(b, p, sz);
    switch (type) {
        case 0: { int8_t v; memcpy(&v, b, 1); return v; }
        case 1: { uint8_t v; memcpy(&v, b, 1); return v; }
        case 2: { int16_t v; memcpy(&v, b, 2); return v; }
        case 3: { uint16_t v; memcpy(&v, b, 2); return v; }
        case 4: { int32_t v; memcpy(&v, b, 4); return v; }
        case 5: { uint32_t v; memcpy(&v, b, 4); return v; }
        case 6: { float v; memcpy(&v, b, 4); return v; }
        default: { double v; memcpy(&v, b, 8); return v; }
    }
}

bool host_is_little_endian() {
    const uint32_t x = 1;
    unsigned char c;
    memcpy(&c, &x, 1);
    return c == 1;
}

int find(const Reader &r, const std::string &name) {
    for (size_t i = 0; i < r.props.size(); i++)
        if (r.props[i].name == name) return (int)i;
    return -1;
}

}  // namespace

extern "C" {

const char *gsr_ply_last_error(void) { return g_ply_error.c_str(); }

int gsr_ply_open(const char *path, int64_t *n_out, int32_t *n_rest_coeffs_out, void **reader_out) {
    if (!path || !n_out || !n_rest_coeffs_out || !reader_out) return fail("gsr_ply_open: null argument");
    *reader_out = nullptr;
    FILE *f = fopen(path, "rb");
    if (!f) return fail(std::string("gsr_ply_open: cannot open ") + path + ": " + strerror(errno));
    Reader *r = new Reader;
    r->f = f;
    char line[1024];
    bool in_vertex = false, have_vertex = false, got_format = false, ended = false;
    if (!fgets(line, sizeof line, f) || strncmp(line, "ply", 3) != 0) {
        fclose(f);
        delete r;
        return fail("gsr_ply_open: not a PLY file");
    }
    while (fgets(line, sizeof line, f)) {
        char a[64] = {0}, b[64] = {0}, c[256] = {0};
        const int k = sscanf(line, "%63s %63s %255s", a, b, c);
        if (k < 1) continue;
        const std::string key = a;
        if (key == "end_header") { ended = true; break; }
        if (key == "comment" || key == "obj_info") continue;
        if (key == "format" && k >= 2) {
            const std::string fmt = b;
            r->format = fmt == "ascii" ? 0 : (fmt == "binary_little_endian" ? 1 : (fmt == "binary_big_endian" ? 2 : -1));
            got_format = r->format >= 0;
        } else if (key == "element" && k >= 3) {
            if (have_vertex && in_vertex) in_vertex = false;  // a later element: vertex data must come first
            if (std::string(b) == "vertex") {
                if (have_vertex) { fclose(f); delete r; return fail("gsr_ply_open: two vertex elements"); }
                r->n = strtoll(c, nullptr, 10);
                in_vertex = have_vertex = true;
            } else if (!have_vertex) {
                fclose(f); delete r;
                return fail("gsr_ply_open: an element precedes `vertex` (unsupported)");
            }
        } else if (key == "property" && in_vertex) {
            if (std::string(b) == "list") { fclose(f); delete r; return fail("gsr_ply_open: list property in vertex element"); }
            Prop p;
            p.type = parse_type(b);
            if (p.type < 0 || k < 3) { fclose(f); delete r; return fail(std::string("gsr_ply_open: bad property line: ") + line); }
            p.name = c;
            p.offset = r->row_bytes;
            r->row_bytes += kTypeSize[p.type];
            r->props.push_back(p);
        }
    }
    if (!ended || !got_format || !have_vertex || r->n < 0) {
        fclose(f);
        delete r;
        return fail("gsr_ply_open: malformed header");
    }
    r->data_start = ftell(f);
    for (const Prop &p : r->props)
        if (p.name.rfind("f_rest_", 0) == 0) r->n_rest++;
    if (r->n_rest % 3 != 0) {  // gaussians.jl:213-216
        const int nr = r->n_rest;
        fclose(f);
        delete r;
        return fail("gsr_ply_open: " + std::to_string(nr) + " f_rest_* properties is not a whole number of SH coefficients per colour channel");
    }
    static const char *required[] = {"x", "y", "z", "f_dc_0", "f_dc_1", "f_dc_2", "opacity", "scale_0", "scale_1", "scale_2",
                                     "rot_0", "rot_1", "rot_2", "rot_3"};
    for (const char *name : required)
        if (find(*r, name) < 0) {
            fclose(f);
            delete r;
            return fail(std::string("gsr_ply_open: missing property ") + name);
        }
    for (int j = 0; j < r->n_rest; j++)
        if (find(*r, "f_rest_" + std::to_string(j)) < 0) {
            const int last = r->n_rest - 1;
            fclose(f);
            delete r;
            return fail("gsr_ply_open: f_rest_* properties are not numbered 0.." + std::to_string(last));
        }
    *n_out = r->n;
    *n_rest_coeffs_out = r->n_rest / 3;
    *reader_out = r;
    return GSR_OK;
}

int gsr_ply_read(void *reader, float *points, float *features_dc, float *features_rest, float *opacities, float *scales,
                 float *rotations) {
    Reader *r = static_cast<Reader *>(reader);
    if (!r) return fail("gsr_ply_read: null reader");
    const int64_t n = r->n;
    const int R = r->n_rest / 3;
    if (n > 0 && (!points || !features_dc || !opacities || !scales || !rotations || (R > 0 && !features_rest)))
        return fail("gsr_ply_read: null output");
    // destination of every property: (array, stride per vertex, index inside the vertex) or ignored
    struct Dst { float *base; int stride, idx; };
    std::vector<Dst> dst(r->props.size(), Dst{nullptr, 0, 0});
    auto bind = [&](const std::string &name, float *base, int stride, int idx) { dst[find(*r, name)] = Dst{base, stride, idx}; };
    bind("x", points, 3, 0); bind("y", points, 3, 1); bind("z", points, 3, 2);
    for (int c = 0; c < 3; c++) bind("f_dc_" + std::to_string(c), features_dc, 3, c);
    for (int c = 0; c < 3; c++)  // file: channel-major c*R + k ; model (3,R,N): c + 3k  (gaussians.jl:228-233)
        for (int k = 0; k < R; k++) bind("f_rest_" + std::to_string(c * R + k), features_rest, 3 * R, c + 3 * k);
    bind("opacity", opacities, 1, 0);
    for (int k = 0; k < 3; k++) bind("scale_" + std::to_string(k), scales, 3, k);
    for (int k = 0; k < 4; k++) bind("rot_" + std::to_string(k), rotations, 4, k);

    if (fseek(r->f, r->data_start, SEEK_SET) != 0) return fail("gsr_ply_read: seek failed");
    if (r->format == 0) {  // ascii
        for (int64_t i = 0; i < n; i++)
            for (size_t p = 0; p < r->props.size(); p++) {
                double v;
                if (fscanf(r->f, "%lf", &v) != 1) return fail("gsr_ply_read: truncated ascii data at vertex " + std::to_string(i));
                if (dst[p].base) dst[p].base[i * dst[p].stride + dst[p].idx] = (float)v;
            }
        return GSR_OK;
    }
    const bool swap = (r->format == 1) != host_is_little_endian();
    const int64_t chunk = 4096;
    std::vector<unsigned char> buf(r->row_bytes * chunk);
    for (int64_t i0 = 0; i0 < n; i0 += chunk) {
        const int64_t m = (n - i0) < chunk ? (n - i0) : chunk;
        if (fread(buf.data(), r->row_bytes, (size_t)m, r->f) != (size_t)m)
            return fail("gsr_ply_read: truncated binary data at vertex " + std::to_string(i0));
        for (size_t p = 0; p < r->props.size(); p++) {
            if (!dst[p].base) continue;
            const Prop &pr = r->props[p];
            float *out = dst[p].base + i0 * dst[p].stride + dst[p].idx;
            const unsigned char *src = buf.data() + pr.offset;
            if (pr.type == 6 && !swap) {
                for (int64_t i = 0; i < m; i++) memcpy(out + i * dst[p].stride, src + i * r->row_bytes, 4);
            } else {
                for (int64_t i = 0; i < m; i++) out[i * dst[p].stride] = (float)load_scalar(src + i * r->row_bytes, pr.type, swap);
            }
        }
    }
    return GSR_OK;
}

void gsr_ply_close(void *reader) {
    Reader *r = static_cast<Reader *>(reader);
    if (!r) return;
    if (r->f) fclose(r->f);
    delete r;
}

int gsr_ply_write(const char *path, int64_t n, int32_t n_rest_coeffs, const float *points, const float *features_dc,
                  const float *features_rest, const float *opacities, const float *scales, const float *rotations) {
    return gsr_ply_write_scales(path, n, n_rest_coeffs, 3, points, features_dc, features_rest, opacities, scales, rotations);
}

int gsr_ply_write_scales(const char *path, int64_t n, int32_t n_rest_coeffs, int32_t n_scale_rows, const float *points,
                         const float *features_dc, const float *features_rest, const float *opacities,
                         const float *scales, const float *rotations) {
    if (!path || n < 0 || n_rest_coeffs < 0) return fail("gsr_ply_write: bad argument");
    // export_ply writes one scale_i per row of `scales` (gaussians.jl:176): (3,N) or the isotropic (1,N)
    if (n_scale_rows != 1 && n_scale_rows != 3) return fail("gsr_ply_write: scales must be (3,N) or isotropic (1,N)");
    const int R = n_rest_coeffs, S = n_scale_rows;
    if (n > 0 && (!points || !features_dc || !opacities || !scales || !rotations || (R > 0 && !features_rest)))
        return fail("gsr_ply_write: null input");
    FILE *f = fopen(path, "wb");
    if (!f) return fail(std::string("gsr_ply_write: cannot open ") + path + ": " + strerror(errno));
    // header exactly as export_ply writes it (gaussians.jl:171-197)
    fprintf(f, "ply\nformat %s 1.0\nelement vertex %lld\n", host_is_little_endian() ? "binary_little_endian" : "binary_big_endian",
            (long long)n);
    for (const char *nm : {"x", "y", "z", "nx", "ny", "nz"}) fprintf(f, "property float %s\n", nm);
    for (int i = 0; i < 3; i++) fprintf(f, "property float f_dc_%d\n", i);
    for (int i = 0; i < 3 * R; i++) fprintf(f, "property float f_rest_%d\n", i);
    fprintf(f, "property float opacity\n");
    for (int i = 0; i < S; i++) fprintf(f, "property float scale_%d\n", i);
    for (int i = 0; i < 4; i++) fprintf(f, "property float rot_%d\n", i);
    fprintf(f, "end_header\n");
    const int row = 3 + 3 + 3 + 3 * R + 1 + S + 4;
    const int64_t chunk = 4096;
    std::vector<float> buf((size_t)row * chunk);
    for (int64_t i0 = 0; i0 < n; i0 += chunk) {
        const int64_t m = (n - i0) < chunk ? (n - i0) : chunk;
        for (int64_t j = 0; j < m; j++) {
            const int64_t i = i0 + j;
            float *o = buf.data() + j * row;
            o[0] = points[3 * i]; o[1] = points[3 * i + 1]; o[2] = points[3 * i + 2];
            o[3] = o[4] = o[5] = 0.f;  // normals: zeros, as the reference writes them
            for (int c = 0; c < 3; c++) o[6 + c] = features_dc[3 * i + c];
            for (int c = 0; c < 3; c++)
                for (int k = 0; k < R; k++) o[9 + c * R + k] = features_rest[(size_t)i * 3 * R + c + 3 * k];
            o[9 + 3 * R] = opacities[i];
            for (int k = 0; k < S; k++) o[10 + 3 * R + k] = scales[S * i + k];
            for (int k = 0; k < 4; k++) o[10 + S + 3 * R + k] = rotations[4 * i + k];
        }
        if (fwrite(buf.data(), sizeof(float) * row, (size_t)m, f) != (size_t)m) {
            fclose(f);
            return fail("gsr_ply_write: short write");
        }
    }
    if (fclose(f) != 0) return fail("gsr_ply_write: close failed");
    return GSR_OK;
}

}  // extern "C"
