// image.cpp — texture decode to RGBA8, standing in for
// image::ImageReader::open(p).decode().to_rgba8() (src/primitives.rs:391-404, src/texture.rs:92).
//   1. "<file>.rgba8" sidecar (u32 width, u32 height, RGBA8 rows) if present — lets a host
//      harness hand in pixels decoded elsewhere;
//   2. PNG: built-in decoder over zlib (non-interlaced, 1..8-bit grey / RGB / palette / +alpha, 16-bit truncated);
//   3. JPEG: built-in baseline + progressive Huffman decoder (jpeg.cpp).
// A file that cannot be decoded yields nullptr: the caller warns and falls back to the
// 1x1 empty texture exactly as the reference does (src/renderer.rs:424-430).
#include "scene.h"

#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>

namespace rc {

bool decode_jpeg(const std::vector<uint8_t>& file, Image& out, std::string* why);  // jpeg.cpp

static bool read_file(const std::string& path, std::vector<uint8_t>& out)
{
    std::ifstream in(path, std::ios::binary);
    if (!in) return false;
    in.seekg(0, std::ios::end);
    std::streamoff n = in.tellg();
    in.seekg(0);
    out.resize((size_t)n);
    if (n) in.read((char*)out.data(), n);
    return (bool)in;
}

static inline uint32_t be32(const uint8_t* p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }

static int paeth(int a, int b, int c)
{
    int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    if (pa <= pb && pa <= pc) return a;
    return pb <= pc ? b : c;
}

static bool decode_png(const std::vector<uint8_t>& f, Image& img, std::string* why)
{
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (f.size() < 33 || memcmp(f.data(), sig, 8) != 0) { if (why) *why = "not a PNG"; return false; }
    uint32_t w = 0, h = 0;
    int depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, plte, trns;
    size_t p = 8;
    while (p + 12 <= f.size()) {
        uint32_t len = be32(&f[p]);
        const uint8_t* type = &f[p + 4];
        const uint8_t* data = &f[p + 8];
        if (p + 12 + len > f.size()) break;
        if (!memcmp(type, "IHDR", 4)) {
            w = be32(data); h = be32(data + 4); depth = data[8]; ctype = data[9]; interlace = data[12];
        } else if (!memcmp(type, "PLTE", 4)) plte.assign(data, data + len);
        else if (!memcmp(type, "tRNS", 4)) trns.assign(data, data + len);
        else if (!memcmp(type, "IDAT", 4)) idat.insert(idat.end(), data, data + len);
        else if (!memcmp(type, "IEND", 4)) break;
        p += 12 + (size_t)len;
    }
    if (!w || !h || interlace) { if (why) *why = "unsupported PNG (interlaced or empty)"; return false; }
    int channels = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    if (!channels || (depth != 1 && depth != 2 && depth != 4 && depth != 8 && depth != 16)) {
        if (why) *why = "unsupported PNG colour type";
        return false;
    }
    size_t bpp_bits = (size_t)channels * depth;
    size_t stride = (w * bpp_bits + 7) / 8;
    size_t bpp = (bpp_bits + 7) / 8;
    std::vector<uint8_t> raw((stride + 1) * h);
    uLongf dlen = (uLongf)raw.size();
    if (uncompress(raw.data(), &dlen, idat.data(), (uLong)idat.size()) != Z_OK || dlen != raw.size()) {
        if (why) *why = "PNG inflate failed";
        return false;
    }
    std::vector<uint8_t> cur(stride), prev(stride, 0);
    img.width = w; img.height = h;
    img.rgba.assign((size_t)w * h * 4, 255);
    for (uint32_t y = 0; y < h; y++) {
        const uint8_t* row = &raw[(stride + 1) * y];
        int ft = row[0];
        for (size_t i = 0; i < stride; i++) {
            int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0, x = row[1 + i];
            switch (ft) {
            case 0: break;
            case 1: x += a; break;
            case 2: x += b; break;
            case 3: x += (a + b) >> 1; break;
            case 4: x += paeth(a, b, c); break;
            default: if (why) *why = "bad PNG filter"; return false;
            }
            cur[i] = (uint8_t)x;
        }
        auto sample = [&](uint32_t x, int ch) -> uint32_t {  // 8-bit value of channel ch at pixel x
            if (depth == 8) return cur[(size_t)x * channels + ch];
            if (depth == 16) return cur[((size_t)x * channels + ch) * 2];
            size_t bit = (size_t)x * depth;
            uint32_t v = (cur[bit >> 3] >> (8 - depth - (bit & 7))) & ((1u << depth) - 1);
            return v;
        };
        uint8_t* o = &img.rgba[(size_t)y * w * 4];
        for (uint32_t x = 0; x < w; x++, o += 4) {
            if (ctype == 3) {
                uint32_t idx = sample(x, 0);
                if (3 * idx + 2 < plte.size()) { o[0] = plte[3 * idx]; o[1] = plte[3 * idx + 1]; o[2] = plte[3 * idx + 2]; }
                else { o[0] = o[1] = o[2] = 0; }
                o[3] = idx < trns.size() ? trns[idx] : 255;
            } else if (ctype == 0 || ctype == 4) {
                uint32_t v = sample(x, 0);
                if (depth < 8) v = v * 255u / ((1u << depth) - 1);
                o[0] = o[1] = o[2] = (uint8_t)v;
                o[3] = ctype == 4 ? (uint8_t)sample(x, 1) : 255;
            } else {
                o[0] = (uint8_t)sample(x, 0); o[1] = (uint8_t)sample(x, 1); o[2] = (uint8_t)sample(x, 2);
                o[3] = ctype == 6 ? (uint8_t)sample(x, 3) : 255;
            }
        }
        prev.swap(cur);
    }
    return true;
}

std::shared_ptr<Image> load_image_rgba8(const std::string& path, std::string* why)
{
    auto img = std::make_shared<Image>();
    std::vector<uint8_t> f;
    if (read_file(path + ".rgba8", f) && f.size() >= 8) {
        uint32_t w, h;
        memcpy(&w, f.data(), 4);
        memcpy(&h, f.data() + 4, 4);
        if ((size_t)w * h * 4 + 8 == f.size() && w && h) {
            img->width = w; img->height = h;
            img->rgba.assign(f.begin() + 8, f.end());
            return img;
        }
    }
    if (!read_file(path, f)) { if (why) *why = path + ": cannot open"; return nullptr; }
    std::string reason;
    if (f.size() > 8 && f[0] == 0x89 && f[1] == 'P') {
        if (decode_png(f, *img, &reason)) return img;
    } else if (f.size() > 3 && f[0] == 0xff && f[1] == 0xd8) {
        if (decode_jpeg(f, *img, &reason)) return img;
    } else {
        reason = "unknown image format";
    }
    if (why) *why = path + ": " + reason;
    return nullptr;
}

}  // namespace rc
