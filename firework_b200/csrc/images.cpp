// Image files in front of and behind the hot path: what the reference does with the `image`
// crate around Renderer::render.
//
//   PNG / baseline JPEG -> RGBA8 texels    image 0.23.9 `image::open(path)…to_rgba()`   (src/texture.rs:285-292, 304)
//   RGB8 pixels -> PNG file                image::save_buffer(…, ColorType::Rgb8)        (src/main.rs:44-60)
//
// The crates (png 0.16.7, jpeg-decoder 0.1.20) are not vendored under the reference; the formats are restated from their
// published specifications.  PNG decoding is exact by definition.  JPEG decoding is only defined up to the IDCT and the
// chroma upsampling a decoder chooses: this one uses an accurate IDCT, the triangle ("fancy") upsampling and the
// fixed-point YCbCr conversion that libjpeg and jpeg-decoder both use, and is compared with PIL in tests/test_host.py.
// Inflate / deflate / crc32 come from zlib.
#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/firework_b200.h"
#include "scene_host.h"

namespace {

bool read_file(const char* path, std::vector<uint8_t>& out, std::string& err) {
    FILE* f = fopen(path, "rb");
    if (!f) { err = std::string("cannot open `") + path + "`"; return false; }
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (n < 0 || n > (1l << 30)) { fclose(f); err = "image file too large"; return false; }
    out.resize((size_t)n);
    size_t got = n ? fread(out.data(), 1, (size_t)n, f) : 0;
    fclose(f);
    if (got != (size_t)n) { err = "short read"; return false; }
    return true;
}

uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
constexpr uint64_t kMaxTexels = 1ull << 26;   // 8192 x 8192: a hostile header must not be able to ask for gigabytes

// ---- PNG ---------------------------------------------------------------------------------------------------
int paeth(int a, int b, int c) {
    int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

bool decode_png(const std::vector<uint8_t>& d, uint32_t& w, uint32_t& h, std::vector<uint8_t>& rgba, std::string& err) {
    size_t pos = 8;
    bool have_ihdr = false;
    int depth = 0, ctype = 0;
    std::vector<uint8_t> idat, plte, trns;
    while (pos + 12 <= d.size()) {
        uint32_t len = be32(&d[pos]);
        if (len > d.size() - pos - 12) { err = "PNG: truncated chunk"; return false; }
        const uint8_t* type = &d[pos + 4];
        const uint8_t* body = &d[pos + 8];
        if (be32(body + len) != (uint32_t)crc32(crc32(0, type, 4), body, len)) { err = "PNG: chunk CRC mismatch"; return false; }
        if (!memcmp(type, "IHDR", 4)) {
            if (len != 13) { err = "PNG: bad IHDR"; return false; }
            w = be32(body); h = be32(body + 4);
            depth = body[8]; ctype = body[9];
            if (body[10] != 0 || body[11] != 0) { err = "PNG: unknown compression / filter method"; return false; }
            if (body[12] != 0) { err = "PNG: interlaced files are not supported"; return false; }
            have_ihdr = true;
        } else if (!memcmp(type, "PLTE", 4)) {
            plte.assign(body, body + len);
        } else if (!memcmp(type, "tRNS", 4)) {
            trns.assign(body, body + len);
        } else if (!memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), body, body + len);
        } else if (!memcmp(type, "IEND", 4)) {
            break;
        }
        pos += 12 + (size_t)len;
    }
    if (!have_ihdr || w == 0 || h == 0 || (uint64_t)w * h > kMaxTexels) { err = "PNG: bad dimensions"; return false; }
    int channels;
    switch (ctype) {
        case 0: channels = 1; break;
        case 2: channels = 3; break;
        case 3: channels = 1; break;
        case 4: channels = 2; break;
        case 6: channels = 4; break;
        default: err = "PNG: bad colour type"; return false;
    }
    const bool depth_ok = ctype == 3 ? (depth == 1 || depth == 2 || depth == 4 || depth == 8)
                                     : ctype == 0 ? (depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16) : (depth == 8 || depth == 16);
    if (!depth_ok) { err = "PNG: bad bit depth"; return false; }
    if (ctype == 3 && plte.size() < 3) { err = "PNG: palette image without PLTE"; return false; }
    const size_t bits_pp = (size_t)channels * depth;
    const size_t stride = ((size_t)w * bits_pp + 7) / 8;
    const size_t bpp = std::max<size_t>(1, bits_pp / 8);   // filter unit
    std::vector<uint8_t> raw((stride + 1) * (size_t)h);
    uLongf raw_len = (uLongf)raw.size();
    int zrc = uncompress(raw.data(), &raw_len, idat.data(), (uLong)idat.size());
    if (zrc != Z_OK || raw_len != raw.size()) { err = "PNG: inflate failed"; return false; }
    std::vector<uint8_t> prev(stride, 0);
    rgba.assign((size_t)w * h * 4, 255);
    for (uint32_t y = 0; y < h; ++y) {
        uint8_t* row = &raw[(stride + 1) * (size_t)y];
        const int filter = row[0];
        uint8_t* cur = row + 1;
        if (filter > 4) { err = "PNG: bad filter type"; return false; }
        for (size_t i = 0; i < stride; ++i) {
            int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
            int pred = filter == 0 ? 0 : filter == 1 ? a : filter == 2 ? b : filter == 3 ? (a + b) / 2 : paeth(a, b, c);
            cur[i] = (uint8_t)(cur[i] + pred);
        }
        memcpy(prev.data(), cur, stride);
        uint8_t* o = &rgba[(size_t)y * w * 4];
        auto sample = [&](size_t idx) -> int {   // idx-th sample of the row
            if (depth == 8) return cur[idx];
            if (depth == 16) return cur[2 * idx];   // high byte
            size_t bit = idx * depth;
            return (cur[bit >> 3] >> (8 - depth - (bit & 7))) & ((1 << depth) - 1);
        };
        for (uint32_t x = 0; x < w; ++x, o += 4) {
            switch (ctype) {
                case 0: {
                    int v = sample(x);
                    if (depth < 8) v = v * 255 / ((1 << depth) - 1);
                    o[0] = o[1] = o[2] = (uint8_t)v;
                    break;
                }
                case 2: o[0] = (uint8_t)sample(3 * x); o[1] = (uint8_t)sample(3 * x + 1); o[2] = (uint8_t)sample(3 * x + 2); break;
                case 3: {
                    size_t k = (size_t)sample(x);
                    if (3 * k + 2 >= plte.size()) { err = "PNG: palette index out of range"; return false; }
                    o[0] = plte[3 * k]; o[1] = plte[3 * k + 1]; o[2] = plte[3 * k + 2];
                    if (k < trns.size()) o[3] = trns[k];
                    break;
                }
                case 4: o[0] = o[1] = o[2] = (uint8_t)sample(2 * x); o[3] = (uint8_t)sample(2 * x + 1); break;
                default: o[0] = (uint8_t)sample(4 * x); o[1] = (uint8_t)sample(4 * x + 1); o[2] = (uint8_t)sample(4 * x + 2); o[3] = (uint8_t)sample(4 * x + 3); break;
            }
        }
    }
    return true;
}

void put_chunk(std::vector<uint8_t>& out, const char* type, const uint8_t* body, size_t len) {
    auto be = [&](uint32_t v) { out.push_back((uint8_t)(v >> 24)); out.push_back((uint8_t)(v >> 16)); out.push_back((uint8_t)(v >> 8)); out.push_back((uint8_t)v); };
    be((uint32_t)len);
    size_t at = out.size();
    out.insert(out.end(), type, type + 4);
    if (len) out.insert(out.end(), body, body + len);
    be((uint32_t)crc32(0, &out[at], (uInt)(len + 4)));
}

bool encode_png_rgb8(uint32_t w, uint32_t h, const uint8_t* rgb, std::vector<uint8_t>& out, std::string& err) {
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    out.assign(sig, sig + 8);
    uint8_t ihdr[13] = {(uint8_t)(w >> 24), (uint8_t)(w >> 16), (uint8_t)(w >> 8), (uint8_t)w,
                        (uint8_t)(h >> 24), (uint8_t)(h >> 16), (uint8_t)(h >> 8), (uint8_t)h, 8, 2, 0, 0, 0};
    put_chunk(out, "IHDR", ihdr, 13);
    const size_t stride = (size_t)w * 3;
    std::vector<uint8_t> raw((stride + 1) * (size_t)h);
    for (uint32_t y = 0; y < h; ++y) {   // filter 1 (Sub) compresses renders much better than None at no decoding risk
        uint8_t* r = &raw[(stride + 1) * (size_t)y];
        const uint8_t* s = rgb + stride * y;
        r[0] = 1;
        for (size_t i = 0; i < stride; ++i) r[1 + i] = (uint8_t)(s[i] - (i >= 3 ? s[i - 3] : 0));
    }
    uLongf cap = compressBound((uLong)raw.size());
    std::vector<uint8_t> z(cap);
    if (compress2(z.data(), &cap, raw.data(), (uLong)raw.size(), 6) != Z_OK) { err = "PNG: deflate failed"; return false; }
    put_chunk(out, "IDAT", z.data(), cap);
    put_chunk(out, "IEND", nullptr, 0);
    return true;
}

// ---- baseline JPEG -----------------------------------------------------------------------------------------
struct Huff {
    // canonical code tables: for each length 1..16 the smallest code, the largest code (+1) and the index of its first symbol
    int mincode[17], maxcode[18], valptr[17];
    uint8_t vals[256];
    bool present = false;
};
struct Comp {
    int id = 0, hs = 1, vs = 1, tq = 0, td = 0, ta = 0;
    int bw = 0, bh = 0;             // blocks per row / column (padded to whole MCUs)
    int w = 0, h = 0;               // real downsampled size
    int pred = 0;
    std::vector<uint8_t> px;        // bw*8 x bh*8 samples
};
struct BitReader {
    const uint8_t* p;
    const uint8_t* e;
    uint32_t acc = 0;
    int n = 0;
    bool hit_marker = false;
    void fill() {
        while (n <= 24) {
            int b = 0;
            if (!hit_marker && p < e) {
                b = *p;
                if (b == 0xff) {
                    if (p + 1 < e && p[1] == 0) { p += 2; }
                    else { hit_marker = true; b = 0; }   // a marker: feed zeros until the caller deals with it
                } else {
                    ++p;
                }
            }
            acc |= (uint32_t)b << (24 - n);
            n += 8;
        }
    }
    int bits(int k) {
        if (k == 0) return 0;
        if (n < k) fill();
        int v = (int)(acc >> (32 - k));
        acc <<= k;
        n -= k;
        return v;
    }
    void reset() { acc = 0; n = 0; hit_marker = false; }
};
const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

bool build_huff(Huff& h, const uint8_t counts[16], const uint8_t* vals, int nvals) {
    int code = 0, k = 0;
    for (int len = 1; len <= 16; ++len) {
        h.valptr[len] = k;
        h.mincode[len] = code;
        code += counts[len - 1];
        k += counts[len - 1];
        h.maxcode[len] = code;   // exclusive
        if (code > (1 << len)) return false;
        code <<= 1;
    }
    h.maxcode[17] = 0x7fffffff;
    if (k != nvals || k > 256) return false;
    memcpy(h.vals, vals, (size_t)nvals);
    h.present = true;
    return true;
}
int decode_sym(BitReader& br, const Huff& h) {
    int code = 0;
    for (int len = 1; len <= 16; ++len) {
        code = (code << 1) | br.bits(1);
        if (code < h.maxcode[len] && code >= h.mincode[len]) return h.vals[h.valptr[len] + code - h.mincode[len]];
    }
    return -1;
}
int extend(int v, int t) { return v < (1 << (t - 1)) ? v - (1 << t) + 1 : v; }

// Separable 8x8 inverse DCT in double precision (the definition), level shift, clamp.
void idct_block(const int coef[64], const uint16_t q[64], uint8_t* out, int stride) {
    static double c[8][8];
    static bool init = false;
    if (!init) {
        for (int x = 0; x < 8; ++x)
            for (int u = 0; u < 8; ++u) c[x][u] = (u == 0 ? std::sqrt(0.125) : 0.5) * std::cos((2 * x + 1) * u * M_PI / 16.0);
        init = true;
    }
    double tmp[64], f[64];
    for (int i = 0; i < 64; ++i) f[i] = (double)coef[i] * q[i];
    for (int v = 0; v < 8; ++v)          // rows: over u
        for (int x = 0; x < 8; ++x) {
            double s = 0;
            for (int u = 0; u < 8; ++u) s += c[x][u] * f[v * 8 + u];
            tmp[v * 8 + x] = s;
        }
    for (int x = 0; x < 8; ++x)          // columns: over v
        for (int y = 0; y < 8; ++y) {
            double s = 0;
            for (int v = 0; v < 8; ++v) s += c[y][v] * tmp[v * 8 + x];
            long r = std::lround(s) + 128;
            out[y * stride + x] = (uint8_t)(r < 0 ? 0 : r > 255 ? 255 : r);
        }
}

bool decode_jpeg(const std::vector<uint8_t>& d, uint32_t& W, uint32_t& H, std::vector<uint8_t>& rgba, std::string& err) {
    uint16_t qt[4][64];
    bool qt_ok[4] = {false, false, false, false};
    Huff hdc[4], hac[4];
    std::vector<Comp> comps;
    int restart = 0, hmax = 1, vmax = 1;
    int adobe_transform = -1;
    bool have_sof = false, done = false;
    size_t pos = 2;
    while (!done && pos + 4 <= d.size()) {
        if (d[pos] != 0xff) { ++pos; continue; }
        int m = d[pos + 1];
        if (m == 0xff) { ++pos; continue; }
        if (m == 0xd8 || (m >= 0xd0 && m <= 0xd7) || m == 0x01) { pos += 2; continue; }
        if (m == 0xd9) break;
        size_t len = ((size_t)d[pos + 2] << 8) | d[pos + 3];
        if (len < 2 || pos + 2 + len > d.size()) { err = "JPEG: truncated segment"; return false; }
        const uint8_t* s = &d[pos + 4];
        size_t n = len - 2;
        if (m == 0xdb) {                                   // DQT
            size_t i = 0;
            while (i < n) {
                int pq = s[i] >> 4, tq = s[i] & 15;
                ++i;
                if (tq > 3 || i + (pq ? 128u : 64u) > n) { err = "JPEG: bad DQT"; return false; }
                for (int k = 0; k < 64; ++k) {
                    qt[tq][kZigzag[k]] = pq ? (uint16_t)((s[i] << 8) | s[i + 1]) : s[i];
                    i += pq ? 2 : 1;
                }
                qt_ok[tq] = true;
            }
        } else if (m == 0xc4) {                            // DHT
            size_t i = 0;
            while (i + 17 <= n) {
                int tc = s[i] >> 4, th = s[i] & 15;
                int total = 0;
                for (int k = 0; k < 16; ++k) total += s[i + 1 + k];
                if (tc > 1 || th > 3 || i + 17 + (size_t)total > n) { err = "JPEG: bad DHT"; return false; }
                if (!build_huff(tc ? hac[th] : hdc[th], &s[i + 1], &s[i + 17], total)) { err = "JPEG: bad Huffman table"; return false; }
                i += 17 + (size_t)total;
            }
        } else if (m == 0xc0 || m == 0xc1) {               // SOF0 / SOF1: sequential, Huffman
            if (n < 6 || s[0] != 8) { err = "JPEG: only 8-bit samples are supported"; return false; }
            H = ((uint32_t)s[1] << 8) | s[2];
            W = ((uint32_t)s[3] << 8) | s[4];
            int nc = s[5];
            if (W == 0 || H == 0 || (uint64_t)W * H > kMaxTexels || (nc != 1 && nc != 3) || n < 6 + 3 * (size_t)nc) { err = "JPEG: unsupported frame header"; return false; }
            comps.assign((size_t)nc, Comp());
            for (int c = 0; c < nc; ++c) {
                comps[c].id = s[6 + 3 * c];
                comps[c].hs = s[7 + 3 * c] >> 4;
                comps[c].vs = s[7 + 3 * c] & 15;
                comps[c].tq = s[8 + 3 * c];
                if (comps[c].hs < 1 || comps[c].hs > 4 || comps[c].vs < 1 || comps[c].vs > 4 || comps[c].tq > 3) { err = "JPEG: bad component"; return false; }
                hmax = std::max(hmax, comps[c].hs);
                vmax = std::max(vmax, comps[c].vs);
            }
            have_sof = true;
        } else if (m == 0xc2 || (m >= 0xc3 && m <= 0xcf && m != 0xc4 && m != 0xc8 && m != 0xcc)) {
            err = "JPEG: only baseline / sequential Huffman files are supported (this one is progressive, lossless or arithmetic)";
            return false;
        } else if (m == 0xdd) {                            // DRI
            if (n >= 2) restart = (s[0] << 8) | s[1];
        } else if (m == 0xee) {                            // APP14 "Adobe"
            if (n >= 12 && !memcmp(s, "Adobe", 5)) adobe_transform = s[11];
        } else if (m == 0xda) {                            // SOS: the (single) scan follows
            if (!have_sof) { err = "JPEG: scan before frame header"; return false; }
            int ns = s[0];
            if (ns != (int)comps.size() || n < 1 + 2 * (size_t)ns + 3) { err = "JPEG: multi-scan sequential files are not supported"; return false; }
            for (int k = 0; k < ns; ++k) {
                int id = s[1 + 2 * k];
                bool found = false;
                for (Comp& c : comps)
                    if (c.id == id) { c.td = s[2 + 2 * k] >> 4; c.ta = s[2 + 2 * k] & 15; found = true; }
                if (!found) { err = "JPEG: scan names an unknown component"; return false; }
            }
            const int mcux = (int)((W + 8 * hmax - 1) / (8 * hmax)), mcuy = (int)((H + 8 * vmax - 1) / (8 * vmax));
            for (Comp& c : comps) {
                if (!qt_ok[c.tq] || c.td > 3 || c.ta > 3 || !hdc[c.td].present || !hac[c.ta].present) { err = "JPEG: missing table"; return false; }
                c.bw = mcux * c.hs; c.bh = mcuy * c.vs;
                c.w = (int)(((uint64_t)W * c.hs + hmax - 1) / hmax);
                c.h = (int)(((uint64_t)H * c.vs + vmax - 1) / vmax);
                c.px.assign((size_t)c.bw * 8 * c.bh * 8, 0);
                c.pred = 0;
            }
            BitReader br{&d[pos + 2 + len], d.data() + d.size()};
            int until_restart = restart;
            for (int my = 0; my < mcuy; ++my)
                for (int mx = 0; mx < mcux; ++mx) {
                    if (restart && until_restart == 0) {
                        // byte-align, expect RSTn
                        br.reset();
                        while (br.p + 1 < br.e && !(br.p[0] == 0xff && br.p[1] >= 0xd0 && br.p[1] <= 0xd7)) ++br.p;
                        if (br.p + 1 < br.e) br.p += 2;
                        for (Comp& c : comps) c.pred = 0;
                        until_restart = restart;
                    }
                    for (Comp& c : comps)
                        for (int by = 0; by < c.vs; ++by)
                            for (int bx = 0; bx < c.hs; ++bx) {
                                int coef[64] = {0};
                                int t = decode_sym(br, hdc[c.td]);
                                if (t < 0 || t > 11) { err = "JPEG: corrupt DC code"; return false; }
                                int diff = t ? extend(br.bits(t), t) : 0;
                                c.pred += diff;
                                coef[0] = c.pred;
                                for (int k = 1; k < 64;) {
                                    int rs = decode_sym(br, hac[c.ta]);
                                    if (rs < 0) { err = "JPEG: corrupt AC code"; return false; }
                                    int r = rs >> 4, sz = rs & 15;
                                    if (sz == 0) {
                                        if (r == 15) { k += 16; continue; }
                                        break;   // EOB
                                    }
                                    k += r;
                                    if (k > 63) { err = "JPEG: corrupt block"; return false; }
                                    coef[kZigzag[k]] = extend(br.bits(sz), sz);
                                    ++k;
                                }
                                const int X = (mx * c.hs + bx) * 8, Y = (my * c.vs + by) * 8;
                                idct_block(coef, qt[c.tq], &c.px[(size_t)Y * c.bw * 8 + X], c.bw * 8);
                            }
                    if (restart) --until_restart;
                }
            done = true;
        }
        pos += 2 + len;
    }
    if (!done) { err = "JPEG: no image data"; return false; }

    // upsample every component to W x H (triangle filter for 2x, replication otherwise), then convert
    std::vector<std::vector<uint8_t>> full(comps.size());
    for (size_t ci = 0; ci < comps.size(); ++ci) {
        Comp& c = comps[ci];
        const int fx = hmax / c.hs, fy = vmax / c.vs;
        const int sw = c.bw * 8;
        auto at = [&](int x, int y) -> int {
            x = std::min(std::max(x, 0), c.w - 1);
            y = std::min(std::max(y, 0), c.h - 1);
            return c.px[(size_t)y * sw + x];
        };
        std::vector<uint8_t>& o = full[ci];
        o.resize((size_t)W * H);
        if (hmax % c.hs || vmax % c.vs) { err = "JPEG: fractional sampling ratios are not supported"; return false; }
        for (uint32_t y = 0; y < H; ++y)
            for (uint32_t x = 0; x < W; ++x) {
                int v;
                if (fx == 1 && fy == 1) {
                    v = at((int)x, (int)y);
                } else if (fx == 2 && fy == 1) {           // h2v1 fancy: 3/4 nearer + 1/4 farther
                    int sx = (int)x >> 1;
                    if ((x & 1) == 0) v = sx == 0 ? at(0, (int)y) : (3 * at(sx, (int)y) + at(sx - 1, (int)y) + 1) >> 2;
                    else v = sx == c.w - 1 ? at(sx, (int)y) : (3 * at(sx, (int)y) + at(sx + 1, (int)y) + 2) >> 2;
                } else if (fx == 2 && fy == 2) {           // h2v2 fancy: 9/16, 3/16, 3/16, 1/16
                    int sx = (int)x >> 1, sy = (int)y >> 1;
                    int ny = (y & 1) == 0 ? sy - 1 : sy + 1;          // the farther row is the one on this side
                    auto colsum = [&](int xx) { return 3 * at(xx, sy) + at(xx, ny); };
                    int cur = colsum(sx);
                    if ((x & 1) == 0) v = sx == 0 ? (cur * 4 + 8) >> 4 : (cur * 3 + colsum(sx - 1) + 8) >> 4;
                    else v = sx == c.w - 1 ? (cur * 4 + 7) >> 4 : (cur * 3 + colsum(sx + 1) + 7) >> 4;
                } else {
                    v = at((int)x / fx, (int)y / fy);
                }
                o[(size_t)y * W + x] = (uint8_t)v;
            }
    }
    rgba.assign((size_t)W * H * 4, 255);
    const bool ycc = comps.size() == 3 && adobe_transform != 0;
    auto clamp8 = [](int v) { return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); };
    for (size_t i = 0; i < (size_t)W * H; ++i) {
        uint8_t* o = &rgba[4 * i];
        if (comps.size() == 1) {
            o[0] = o[1] = o[2] = full[0][i];
        } else if (!ycc) {
            o[0] = full[0][i]; o[1] = full[1][i]; o[2] = full[2][i];
        } else {                                            // JFIF YCbCr -> RGB, 16-bit fixed point as in libjpeg's jdcolor.c
            const int y = full[0][i], cb = full[1][i] - 128, cr = full[2][i] - 128;
            const int half = 1 << 15;
            o[0] = clamp8(y + ((91881 * cr + half) >> 16));
            o[1] = clamp8(y + ((-22554 * cb - 46802 * cr + half) >> 16));
            o[2] = clamp8(y + ((116130 * cb + half) >> 16));
        }
    }
    return true;
}

}  // namespace

extern "C" {

int fw_image_load(const char* path, uint32_t* width, uint32_t* height, uint8_t** rgba) {
    if (!path || !width || !height || !rgba) return fw::set_last_error(FW_ERR_ARG, "fw_image_load: null argument");
    try {
        std::vector<uint8_t> d, px;
        std::string err;
        if (!read_file(path, d, err)) return fw::set_last_error(FW_ERR_ASSET, err);
        static const uint8_t png_sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
        uint32_t w = 0, h = 0;
        bool ok;
        if (d.size() >= 8 && !memcmp(d.data(), png_sig, 8)) ok = decode_png(d, w, h, px, err);
        else if (d.size() >= 4 && d[0] == 0xff && d[1] == 0xd8) ok = decode_jpeg(d, w, h, px, err);
        else { ok = false; err = "not a PNG or JPEG file"; }
        if (!ok) return fw::set_last_error(FW_ERR_ASSET, std::string(path) + ": " + err);
        uint8_t* out = static_cast<uint8_t*>(malloc(px.size()));
        if (!out) return fw::set_last_error(FW_ERR_ASSET, "out of host memory");
        memcpy(out, px.data(), px.size());
        *width = w; *height = h; *rgba = out;
        return FW_OK;
    } catch (const std::exception& e) {
        return fw::set_last_error(FW_ERR_ASSET, std::string("fw_image_load: ") + e.what());
    }
}
void fw_image_free(uint8_t* rgba) { free(rgba); }

int fw_png_write(const char* path, uint32_t width, uint32_t height, const uint8_t* rgb) {
    if (!path || !rgb || width == 0 || height == 0) return fw::set_last_error(FW_ERR_ARG, "fw_png_write: bad argument");
    try {
        std::vector<uint8_t> out;
        std::string err;
        if (!encode_png_rgb8(width, height, rgb, out, err)) return fw::set_last_error(FW_ERR_ASSET, err);
        FILE* f = fopen(path, "wb");
        if (!f) return fw::set_last_error(FW_ERR_ASSET, std::string("cannot write `") + path + "`");
        size_t put = fwrite(out.data(), 1, out.size(), f);
        int crc = fclose(f);
        if (put != out.size() || crc != 0) return fw::set_last_error(FW_ERR_ASSET, std::string("short write to `") + path + "`");
        return FW_OK;
    } catch (const std::exception& e) {
        return fw::set_last_error(FW_ERR_ASSET, std::string("fw_png_write: ") + e.what());
    }
}

}  // extern "C"
