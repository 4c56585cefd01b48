// Asset ingestion in front of the hot path: what the reference's example programs do with
// third-party crates before they call Renderer::render.
//
//   Wavefront OBJ  -> indexed triangle models       tobj 1.0.0 `load_obj`  (examples/suzanne.rs:15-51, teapot.rs:17-64)
//   Radiance .hdr  -> fp32 RGB texels               image 0.23.9 `HdrDecoder::read_image_hdr` (examples/hdri_test.rs:45-67)
//
// Neither crate is vendored under the reference; their behaviour is restated from the published formats and pinned
// by the reference's own artefacts: loading suzanne.obj / teapot.obj must reproduce the vertex and index arrays of
// the committed scenes/suzanne.yml / scenes/teapot.yml bit for bit (tests/test_host.py).
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/firework_b200.h"
#include "scene_host.h"

namespace {

struct ObjModel {
    std::string name;
    std::vector<float> positions, normals, texcoords;   // 3 / 3 / 2 floats per vertex
    std::vector<uint32_t> indices;                      // 3 per triangle
};

// One corner of a face: 0-based indices into the file's v / vt / vn lists, -1 = absent.
struct Corner {
    long v, vt, vn;
    bool operator<(const Corner& o) const { return std::tie(v, vt, vn) < std::tie(o.v, o.vt, o.vn); }
};

bool parse_index(const std::string& tok, size_t count, long& out) {
    if (tok.empty()) { out = -1; return true; }
    char* end = nullptr;
    long i = strtol(tok.c_str(), &end, 10);
    if (end == tok.c_str() || *end != 0 || i == 0) return false;
    out = i > 0 ? i - 1 : (long)count + i;   // negative indices count back from the current end of the list
    return out >= 0;
}

}  // namespace

struct fw_obj {
    std::vector<ObjModel> models;
};

extern "C" {

// tobj::load_obj: every `o` / `g` statement (and a material change) closes the model gathered so far; faces are
// fan-triangulated ((0,1,2), (0,2,3), ...); vertices are re-indexed per model so that each distinct v/vt/vn triple
// becomes one vertex, in order of first use.
static int obj_load_impl(const char* path, fw_obj** out);
int fw_obj_load(const char* path, fw_obj** out) {
    if (!path || !out) return fw::set_last_error(FW_ERR_ARG, "null argument");
    try {   // nothing may propagate through the C ABI
        return obj_load_impl(path, out);
    } catch (const std::exception& e) {
        return fw::set_last_error(FW_ERR_PARSE, std::string(path) + ": " + e.what());
    } catch (...) {
        return fw::set_last_error(FW_ERR_PARSE, std::string(path) + ": internal error");
    }
}
static int obj_load_impl(const char* path, fw_obj** out) {
    std::ifstream f(path);
    if (!f) return fw::set_last_error(FW_ERR_PARSE, std::string("cannot open ") + path);
    std::vector<float> pos, nrm, tex;
    std::vector<std::vector<Corner>> faces;   // faces of the model being gathered
    std::string name = "unnamed_object", material;
    auto obj = new fw_obj();
    auto flush = [&]() {
        if (faces.empty()) return;
        ObjModel m;
        m.name = name;
        std::map<Corner, uint32_t> seen;
        auto add = [&](const Corner& c) {
            auto it = seen.find(c);
            if (it != seen.end()) { m.indices.push_back(it->second); return; }
            uint32_t idx = (uint32_t)(m.positions.size() / 3);
            for (int k = 0; k < 3; ++k) m.positions.push_back(pos[3 * c.v + k]);
            if (c.vt >= 0) for (int k = 0; k < 2; ++k) m.texcoords.push_back(tex[2 * c.vt + k]);
            if (c.vn >= 0) for (int k = 0; k < 3; ++k) m.normals.push_back(nrm[3 * c.vn + k]);
            seen.emplace(c, idx);
            m.indices.push_back(idx);
        };
        for (const auto& face : faces)
            for (size_t i = 1; i + 1 < face.size(); ++i) { add(face[0]); add(face[i]); add(face[i + 1]); }
        obj->models.push_back(std::move(m));
        faces.clear();
    };
    std::string line;
    size_t lineno = 0;
    auto fail = [&](const std::string& what) {
        delete obj;
        return fw::set_last_error(FW_ERR_PARSE, std::string(path) + ":" + std::to_string(lineno) + ": " + what);
    };
    while (std::getline(f, line)) {
        ++lineno;
        while (!line.empty() && (line.back() == '\r' || line.back() == ' ' || line.back() == '\t')) line.pop_back();
        std::istringstream ss(line);
        std::string kw;
        if (!(ss >> kw) || kw[0] == '#') continue;
        if (kw == "v" || kw == "vn") {
            float x, y, z;
            if (!(ss >> x >> y >> z)) return fail("expected three numbers after `" + kw + "`");
            auto& dst = kw == "v" ? pos : nrm;
            dst.push_back(x); dst.push_back(y); dst.push_back(z);
        } else if (kw == "vt") {
            float u, v = 0.0f;
            if (!(ss >> u)) return fail("expected numbers after `vt`");
            ss >> v;
            tex.push_back(u); tex.push_back(v);
        } else if (kw == "f") {
            std::vector<Corner> face;
            std::string tok;
            while (ss >> tok) {
                std::string part[3];
                int n = 0;
                size_t b = 0;
                for (size_t i = 0; i <= tok.size() && n < 3; ++i)
                    if (i == tok.size() || tok[i] == '/') { part[n++] = tok.substr(b, i - b); b = i + 1; }
                Corner c{-1, -1, -1};
                if (!parse_index(part[0], pos.size() / 3, c.v) || c.v < 0 || (size_t)c.v >= pos.size() / 3) return fail("bad vertex index `" + tok + "`");
                if (!parse_index(part[1], tex.size() / 2, c.vt) || (c.vt >= 0 && (size_t)c.vt >= tex.size() / 2)) return fail("bad texcoord index `" + tok + "`");
                if (!parse_index(part[2], nrm.size() / 3, c.vn) || (c.vn >= 0 && (size_t)c.vn >= nrm.size() / 3)) return fail("bad normal index `" + tok + "`");
                face.push_back(c);
            }
            if (face.size() >= 3) faces.push_back(std::move(face));   // points and lines carry no surface
        } else if (kw == "o" || kw == "g") {
            flush();
            std::string rest;
            std::getline(ss, rest);
            size_t a = rest.find_first_not_of(" \t");
            name = a == std::string::npos ? "unnamed_object" : rest.substr(a);
        } else if (kw == "usemtl") {
            std::string m;
            ss >> m;
            if (!faces.empty() && m != material) flush();
            material = m;
        }
        // mtllib, s, and anything else: not geometry
    }
    flush();
    // a model mixes corners with and without vt / vn only in malformed files: keep the arrays all-or-nothing
    for (ObjModel& m : obj->models) {
        size_t nv = m.positions.size() / 3;
        if (m.normals.size() != 3 * nv) m.normals.clear();
        if (m.texcoords.size() != 2 * nv) m.texcoords.clear();
    }
    *out = obj;
    return FW_OK;
}
int fw_obj_num_models(const fw_obj* o) { return o ? (int)o->models.size() : 0; }
const char* fw_obj_model_name(const fw_obj* o, int m) {
    if (!o || m < 0 || m >= (int)o->models.size()) return nullptr;
    return o->models[m].name.c_str();
}
int fw_obj_model_sizes(const fw_obj* o, int m, uint32_t sizes[4]) {
    if (!o || !sizes || m < 0 || m >= (int)o->models.size()) return fw::set_last_error(FW_ERR_ARG, "bad model index");
    const ObjModel& q = o->models[m];
    sizes[0] = (uint32_t)q.positions.size(); sizes[1] = (uint32_t)q.normals.size();
    sizes[2] = (uint32_t)q.texcoords.size(); sizes[3] = (uint32_t)q.indices.size();
    return FW_OK;
}
int fw_obj_model_copy(const fw_obj* o, int m, float* positions, float* normals, float* texcoords, uint32_t* indices) {
    if (!o || m < 0 || m >= (int)o->models.size()) return fw::set_last_error(FW_ERR_ARG, "bad model index");
    const ObjModel& q = o->models[m];
    if (positions) memcpy(positions, q.positions.data(), q.positions.size() * 4);
    if (normals) memcpy(normals, q.normals.data(), q.normals.size() * 4);
    if (texcoords) memcpy(texcoords, q.texcoords.data(), q.texcoords.size() * 4);
    if (indices) memcpy(indices, q.indices.data(), q.indices.size() * 4);
    return FW_OK;
}
void fw_obj_destroy(fw_obj* o) { delete o; }

// Radiance RGBE (.hdr / .pic): text header ("#?RADIANCE", FORMAT=32-bit_rle_rgbe, blank line, "-Y h +X w"), then
// scanlines either flat (4 bytes per pixel) or new-style RLE (2 2 hi lo, then each of the four channels run-length
// coded).  Texel -> float as image 0.23.9's Rgbe8Pixel::to_hdr: (0,0,0) if e == 0, else c * 2^(e - 136).
int fw_hdr_load(const char* path, uint32_t* w_out, uint32_t* h_out, float** rgb_out) {
    if (!path || !w_out || !h_out || !rgb_out) return fw::set_last_error(FW_ERR_ARG, "null argument");
    FILE* f = fopen(path, "rb");
    if (!f) return fw::set_last_error(FW_ERR_PARSE, std::string("cannot open ") + path);
    auto fail = [&](const std::string& what) {
        fclose(f);
        return fw::set_last_error(FW_ERR_PARSE, std::string(path) + ": " + what);
    };
    auto read_line = [&](std::string& s) {
        s.clear();
        int c;
        while ((c = fgetc(f)) != EOF && c != '\n') s.push_back((char)c);
        return c != EOF || !s.empty();
    };
    std::string line;
    if (!read_line(line) || (line.rfind("#?RADIANCE", 0) != 0 && line.rfind("#?RGBE", 0) != 0)) return fail("not a Radiance HDR file");
    bool format_ok = false;
    for (;;) {
        if (!read_line(line)) return fail("truncated header");
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.empty()) break;
        if (line.rfind("FORMAT=", 0) == 0) format_ok = line == "FORMAT=32-bit_rle_rgbe";
    }
    if (!format_ok) return fail("unsupported FORMAT (only 32-bit_rle_rgbe)");
    if (!read_line(line)) return fail("missing resolution line");
    long hh = 0, ww = 0;
    if (sscanf(line.c_str(), "-Y %ld +X %ld", &hh, &ww) != 2 || hh <= 0 || ww <= 0) return fail("unsupported orientation `" + line + "`");
    // a header is untrusted input: bound the size before any allocation (65536 x 65536 x 12 B would already be 48 GiB)
    if (ww > 65536 || hh > 65536 || (unsigned long long)ww * (unsigned long long)hh > (1ull << 28))
        return fail("unreasonable image size " + std::to_string(ww) + " x " + std::to_string(hh));
    const size_t w = (size_t)ww, h = (size_t)hh;
    float* rgb = (float*)malloc(w * h * 3 * sizeof(float));
    if (!rgb) return fail("out of memory");
    std::vector<unsigned char> scan;
    try {
        scan.resize(w * 4);
    } catch (const std::exception&) {   // nothing may propagate through the C ABI
        free(rgb);
        return fail("out of memory");
    }
    for (size_t y = 0; y < h; ++y) {
        unsigned char head[4];
        if (fread(head, 1, 4, f) != 4) { free(rgb); return fail("truncated pixel data"); }
        if (w >= 8 && w < 32768 && head[0] == 2 && head[1] == 2 && (((size_t)head[2] << 8) | head[3]) == w) {
            for (int ch = 0; ch < 4; ++ch) {   // each channel of the scanline separately
                size_t x = 0;
                while (x < w) {
                    int n = fgetc(f);
                    if (n == EOF) { free(rgb); return fail("truncated run"); }
                    if (n > 128) {   // run of one value
                        n -= 128;
                        int v = fgetc(f);
                        if (v == EOF || x + (size_t)n > w) { free(rgb); return fail("bad run"); }
                        for (int k = 0; k < n; ++k) scan[4 * (x++) + ch] = (unsigned char)v;
                    } else {         // literal values
                        if (n == 0 || x + (size_t)n > w) { free(rgb); return fail("bad literal run"); }
                        for (int k = 0; k < n; ++k) {
                            int v = fgetc(f);
                            if (v == EOF) { free(rgb); return fail("truncated literal run"); }
                            scan[4 * (x++) + ch] = (unsigned char)v;
                        }
                    }
                }
            }
        } else {   // flat scanline: the four bytes already read are the first pixel
            memcpy(scan.data(), head, 4);
            if (w > 1 && fread(scan.data() + 4, 1, (w - 1) * 4, f) != (w - 1) * 4) { free(rgb); return fail("truncated pixel data"); }
        }
        for (size_t x = 0; x < w; ++x) {
            const unsigned char* p = &scan[4 * x];
            float* o = &rgb[3 * (y * w + x)];
            if (p[3] == 0) { o[0] = o[1] = o[2] = 0.0f; continue; }
            float scale = exp2f((float)p[3] - (128.0f + 8.0f));
            o[0] = scale * (float)p[0]; o[1] = scale * (float)p[1]; o[2] = scale * (float)p[2];
        }
    }
    fclose(f);
    *w_out = (uint32_t)w; *h_out = (uint32_t)h; *rgb_out = rgb;
    return FW_OK;
}
void fw_hdr_free(float* rgb) { free(rgb); }

}  // extern "C"
