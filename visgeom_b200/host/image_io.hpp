// image_io.hpp -- reading calibration images as 8-bit grey (the reference calls cv::imread(fileName, 0),
// unified_calibration.cpp:1025; OpenCV is not part of this engine).  Formats: binary PGM (P5, maxval <= 255) and
// non-interlaced 8- or 16-bit PNG (grey, grey + alpha, RGB, RGBA, 8-bit palette; inflate through zlib).  Colour PNGs are converted with
// OpenCV's fixed-point BGR2GRAY weights (R 4899, G 9617, B 1868, >> 14); imread's own PNG path lets libpng do that
// conversion, which can differ by one grey level -- calibration images are grey in practice.
#pragma once

#include <zlib.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
#include <string>
#include <vector>

#include "visgeom_b200/corner_detector.hpp"

namespace visgeom_b200 {
namespace image_io {

inline bool read_file(const std::string &name, std::vector<uint8_t> &bytes)
{
    std::ifstream f(name, std::ios::binary);
    if (!f) return false;
    bytes.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
    return true;
}

inline bool decode_pgm(const std::vector<uint8_t> &b, Mat8u &img)
{
    size_t p = 2;
    auto next_int = [&](int &v) {
        for (;;) {
            while (p < b.size() && (b[p] == ' ' || b[p] == '\n' || b[p] == '\r' || b[p] == '\t')) p++;
            if (p < b.size() && b[p] == '#') { while (p < b.size() && b[p] != '\n') p++; continue; }
            break;
        }
        if (p >= b.size() || b[p] < '0' || b[p] > '9') return false;
        v = 0;
        while (p < b.size() && b[p] >= '0' && b[p] <= '9') v = v * 10 + (b[p++] - '0');
        return true;
    };
    int w, h, maxval;
    if (!next_int(w) || !next_int(h) || !next_int(maxval) || maxval < 1 || maxval > 255 || w < 1 || h < 1) return false;
    p++;                                                // the single whitespace after maxval
    if (b.size() < p + (size_t)w * h) return false;
    img.cols = w; img.rows = h;
    img.data.assign(b.begin() + p, b.begin() + p + (size_t)w * h);
    return true;
}

inline uint32_t be32(const uint8_t *p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }

inline bool decode_png(const std::vector<uint8_t> &b, Mat8u &img)
{
    static const uint8_t sig[8] = {137, 80, 78, 71, 13, 10, 26, 10};
    if (b.size() < 8 || std::memcmp(b.data(), sig, 8) != 0) return false;
    size_t p = 8;
    uint32_t w = 0, h = 0;
    int depth = 0, colour = 0, interlace = 0;
    std::vector<uint8_t> z, palette;
    while (p + 12 <= b.size()) {
        const uint32_t len = be32(&b[p]);
        const char *type = reinterpret_cast<const char *>(&b[p + 4]);
        if (p + 12 + len > b.size()) return false;
        const uint8_t *d = &b[p + 8];
        if (!std::memcmp(type, "IHDR", 4) && len >= 13) { w = be32(d); h = be32(d + 4); depth = d[8]; colour = d[9]; interlace = d[12]; }
        else if (!std::memcmp(type, "PLTE", 4)) palette.assign(d, d + len);
        else if (!std::memcmp(type, "IDAT", 4)) z.insert(z.end(), d, d + len);
        else if (!std::memcmp(type, "IEND", 4)) break;
        p += 12 + len;
    }
    if (w == 0 || h == 0 || (depth != 8 && depth != 16) || interlace != 0) return false;
    if (colour == 3 && (depth != 8 || palette.size() < 3)) return false;
    const int samples = colour == 0 || colour == 3 ? 1 : colour == 4 ? 2 : colour == 2 ? 3 : colour == 6 ? 4 : 0;
    if (samples == 0) return false;
    const int bps = depth / 8, ch = samples * bps;     // bytes per sample (16-bit: big endian, the high byte is kept -- what
                                                       // libpng's strip_16 leaves cv::imread), bytes per pixel
    const size_t stride = (size_t)w * ch;
    std::vector<uint8_t> raw((stride + 1) * h);
    uLongf out_len = (uLongf)raw.size();
    if (uncompress(raw.data(), &out_len, z.data(), (uLong)z.size()) != Z_OK || out_len != raw.size()) return false;
    std::vector<uint8_t> prev(stride, 0), cur(stride);
    img.cols = (int)w; img.rows = (int)h;
    img.data.resize((size_t)w * h);
    for (uint32_t y = 0; y < h; y++) {
        const uint8_t *line = &raw[(stride + 1) * y];
        const int filter = line[0];
        for (size_t i = 0; i < stride; i++) {
            const int a = i >= (size_t)ch ? cur[i - ch] : 0, up = prev[i], c = i >= (size_t)ch ? prev[i - ch] : 0;
            int pred = 0;
            switch (filter) {
            case 0: pred = 0; break;
            case 1: pred = a; break;
            case 2: pred = up; break;
            case 3: pred = (a + up) / 2; break;
            case 4: {
                const int pa = std::abs(up - c), pb = std::abs(a - c), pc = std::abs(a + up - 2 * c);
                pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? up : c);
                break;
            }
            default: return false;
            }
            cur[i] = (uint8_t)(line[1 + i] + pred);
        }
        uint8_t *dst = &img.data[(size_t)w * y];
        for (uint32_t x = 0; x < w; x++) {
            const uint8_t *px = &cur[(size_t)x * ch];
            if (colour == 3) {
                const size_t e = (size_t)px[0] * 3;
                if (e + 2 >= palette.size()) return false;
                dst[x] = (uint8_t)((palette[e] * 4899 + palette[e + 1] * 9617 + palette[e + 2] * 1868 + 8192) >> 14);
            } else if (samples <= 2) dst[x] = px[0];
            else dst[x] = (uint8_t)((px[0] * 4899 + px[bps] * 9617 + px[2 * bps] * 1868 + 8192) >> 14);
        }
        prev.swap(cur);
    }
    return true;
}

// cv::imread(name, 0): an empty image when the file cannot be read or decoded
inline Mat8u imread_grey(const std::string &name)
{
    Mat8u img;
    std::vector<uint8_t> bytes;
    if (!read_file(name, bytes) || bytes.size() < 8) return img;
    bool ok = false;
    if (bytes[0] == 'P' && bytes[1] == '5') ok = decode_pgm(bytes, img);
    else ok = decode_png(bytes, img);
    if (!ok) img = Mat8u();
    return img;
}

}  // namespace image_io
}  // namespace visgeom_b200
