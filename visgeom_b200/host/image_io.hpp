// image_io.hpp -- reading calibration images as 8-bit grey (the reference calls cv::imread(fileName, 0),
// unified_calibration.cpp:1025; OpenCV is not part of this engine).  Formats: baseline JPEG (below), binary PGM (P5, maxval <= 255) and
// non-interlaced 8- or 16-bit PNG (grey, grey + alpha, RGB, RGBA, 8-bit palette; inflate through zlib).  Colour PNGs are converted with
// OpenCV's fixed-point BGR2GRAY weights (R 4899, G 9617, B 1868, >> 14); imread's own PNG path lets libpng do that
// conversion, which can differ by one grey level -- calibration images are grey in practice.
#pragma once

#include <zlib.h>

#include <algorithm>
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

// ---- baseline JPEG, luminance only -----------------------------------------------------------------------------------
// cv::imread(name, 0) asks libjpeg for JCS_GRAYSCALE: of a YCbCr or grey file only the Y component is reconstructed
// (the others are entropy-decoded and dropped), so reading a calibration picture as grey needs no colour conversion and
// no chroma upsampling.  Sequential Huffman files (SOF0 / SOF1, 8-bit samples; restart intervals; any sampling factors as
// long as Y is sampled at full resolution); the inverse DCT is libjpeg's default integer one (jidctint.c, "islow": 13-bit
// constants, two passes with an intermediate scaling of 2 bits), so the pixels are the ones OpenCV returns.  Progressive
// files are not decoded (an empty image, as for any unreadable file).
struct JpegHuff {
    int mincode[17], maxcode[18], valptr[17];
    uint8_t vals[256];
    bool set = false;
    void build(const uint8_t *bits, const uint8_t *v, int n)
    {
        std::memcpy(vals, v, (size_t)n);
        int code = 0, k = 0;
        for (int l = 1; l <= 16; l++) {
            valptr[l] = k;
            mincode[l] = code;
            code += bits[l - 1];
            k += bits[l - 1];
            maxcode[l] = bits[l - 1] ? code - 1 : -1;
            code <<= 1;
        }
        maxcode[17] = 0x7fffffff;
        set = true;
    }
};

struct JpegBits {
    const uint8_t *p, *end;
    uint32_t acc = 0;
    int n = 0;
    bool bad = false;
    int bit()
    {
        if (n == 0) {
            if (p >= end) { bad = true; return 0; }
            uint8_t b = *p++;
            if (b == 0xFF) {
                if (p < end && *p == 0x00) p++;              // stuffed zero
                else { p--; bad = true; return 0; }          // a marker inside the entropy-coded segment
            }
            acc = b; n = 8;
        }
        n--;
        return (acc >> n) & 1;
    }
    int receive(int s) { int v = 0; for (int i = 0; i < s; i++) v = (v << 1) | bit(); return v; }
    int decode(const JpegHuff &h)
    {
        int code = 0;
        for (int l = 1; l <= 16; l++) {
            code = (code << 1) | bit();
            if (h.maxcode[l] >= 0 && code <= h.maxcode[l] && code >= h.mincode[l]) return h.vals[h.valptr[l] + code - h.mincode[l]];
        }
        bad = true;
        return 0;
    }
    void reset() { n = 0; acc = 0; }
};

inline int jpeg_extend(int v, int s) { return s == 0 ? 0 : (v < (1 << (s - 1)) ? v - (1 << s) + 1 : v); }

// jidctint.c jpeg_idct_islow on dequantised coefficients (natural order) -> 8 x 8 samples
inline void jpeg_idct_islow(const int *in, uint8_t *out, size_t out_stride)
{
    const int CB = 13, P1 = 2;
    const long F0298 = 2446, F0390 = 3196, F0541 = 4433, F0765 = 6270, F0899 = 7373, F1175 = 9633, F1501 = 12299, F1847 = 15137,
               F1961 = 16069, F2053 = 16819, F2562 = 20995, F3072 = 25172;
    auto descale = [](long x, int n) { return (x + (1L << (n - 1))) >> n; };
    long ws[64];
    for (int c = 0; c < 8; c++) {
        const int *i = in + c;
        if (!(i[8] | i[16] | i[24] | i[32] | i[40] | i[48] | i[56])) {
            const long dc = (long)i[0] * (1 << P1);
            for (int r = 0; r < 8; r++) ws[r * 8 + c] = dc;
            continue;
        }
        long z2 = i[16], z3 = i[48];
        long z1 = (z2 + z3) * F0541;
        long tmp2 = z1 + z3 * (-F1847), tmp3 = z1 + z2 * F0765;
        z2 = i[0]; z3 = i[32];
        long tmp0 = (z2 + z3) * (1L << CB), tmp1 = (z2 - z3) * (1L << CB);
        const long tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
        tmp0 = i[56]; tmp1 = i[40]; tmp2 = i[24]; tmp3 = i[8];
        z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
        long z4 = tmp1 + tmp3;
        const long z5 = (z3 + z4) * F1175;
        tmp0 *= F0298; tmp1 *= F2053; tmp2 *= F3072; tmp3 *= F1501;
        z1 *= -F0899; z2 *= -F2562; z3 *= -F1961; z4 *= -F0390;
        z3 += z5; z4 += z5;
        tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
        ws[0 * 8 + c] = descale(tmp10 + tmp3, CB - P1); ws[7 * 8 + c] = descale(tmp10 - tmp3, CB - P1);
        ws[1 * 8 + c] = descale(tmp11 + tmp2, CB - P1); ws[6 * 8 + c] = descale(tmp11 - tmp2, CB - P1);
        ws[2 * 8 + c] = descale(tmp12 + tmp1, CB - P1); ws[5 * 8 + c] = descale(tmp12 - tmp1, CB - P1);
        ws[3 * 8 + c] = descale(tmp13 + tmp0, CB - P1); ws[4 * 8 + c] = descale(tmp13 - tmp0, CB - P1);
    }
    auto limit = [](long v) { v += 128; return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); };
    for (int r = 0; r < 8; r++) {
        const long *w = ws + r * 8;
        uint8_t *o = out + r * out_stride;
        if (!(w[1] | w[2] | w[3] | w[4] | w[5] | w[6] | w[7])) {
            const uint8_t dc = limit(descale(w[0], P1 + 3));
            for (int c = 0; c < 8; c++) o[c] = dc;
            continue;
        }
        long z2 = w[2], z3 = w[6];
        long z1 = (z2 + z3) * F0541;
        long tmp2 = z1 + z3 * (-F1847), tmp3 = z1 + z2 * F0765;
        long tmp0 = (w[0] + w[4]) * (1L << CB), tmp1 = (w[0] - w[4]) * (1L << CB);
        const long tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
        tmp0 = w[7]; tmp1 = w[5]; tmp2 = w[3]; tmp3 = w[1];
        z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
        long z4 = tmp1 + tmp3;
        const long z5 = (z3 + z4) * F1175;
        tmp0 *= F0298; tmp1 *= F2053; tmp2 *= F3072; tmp3 *= F1501;
        z1 *= -F0899; z2 *= -F2562; z3 *= -F1961; z4 *= -F0390;
        z3 += z5; z4 += z5;
        tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
        const int S = CB + P1 + 3;
        o[0] = limit(descale(tmp10 + tmp3, S)); o[7] = limit(descale(tmp10 - tmp3, S));
        o[1] = limit(descale(tmp11 + tmp2, S)); o[6] = limit(descale(tmp11 - tmp2, S));
        o[2] = limit(descale(tmp12 + tmp1, S)); o[5] = limit(descale(tmp12 - tmp1, S));
        o[3] = limit(descale(tmp13 + tmp0, S)); o[4] = limit(descale(tmp13 - tmp0, S));
    }
}

inline bool decode_jpeg(const std::vector<uint8_t> &b, Mat8u &img)
{
    static const int ZZ[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                               41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                               30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
    if (b.size() < 4 || b[0] != 0xFF || b[1] != 0xD8) return false;
    int quant[4][64] = {};
    bool have_q[4] = {false, false, false, false};
    JpegHuff dc[4], ac[4];
    struct Comp { int id, h, v, tq, td = 0, ta = 0, pred = 0; } comp[4];
    int ncomp = 0, width = 0, height = 0, restart = 0;
    size_t p = 2;
    while (p + 4 <= b.size()) {
        if (b[p] != 0xFF) return false;
        const int m = b[p + 1];
        if (m == 0xFF) { p++; continue; }                    // fill byte
        p += 2;
        if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) continue;
        if (m == 0xD9) return false;                          // EOI before any scan
        if (p + 2 > b.size()) return false;
        const size_t len = ((size_t)b[p] << 8) | b[p + 1];
        if (len < 2 || p + len > b.size()) return false;
        const uint8_t *d = &b[p + 2];
        const size_t n = len - 2;
        if (m == 0xDB) {                                      // DQT
            size_t q = 0;
            while (q < n) {
                const int pq = d[q] >> 4, tq = d[q] & 15;
                q++;
                if (tq > 3 || q + (pq ? 128u : 64u) > n) return false;
                for (int i = 0; i < 64; i++) { quant[tq][ZZ[i]] = pq ? ((d[q] << 8) | d[q + 1]) : d[q]; q += pq ? 2 : 1; }
                have_q[tq] = true;
            }
        } else if (m == 0xC0 || m == 0xC1) {                  // SOF0 / SOF1: sequential, Huffman
            if (n < 6 || d[0] != 8) return false;
            height = (d[1] << 8) | d[2]; width = (d[3] << 8) | d[4]; ncomp = d[5];
            if (width < 1 || height < 1 || (ncomp != 1 && ncomp != 3) || n < 6 + 3u * ncomp) return false;
            for (int c = 0; c < ncomp; c++) {
                comp[c].id = d[6 + 3 * c]; comp[c].h = d[7 + 3 * c] >> 4; comp[c].v = d[7 + 3 * c] & 15; comp[c].tq = d[8 + 3 * c] & 3;
                if (comp[c].h < 1 || comp[c].h > 4 || comp[c].v < 1 || comp[c].v > 4) return false;
            }
        } else if (m == 0xC2 || (m >= 0xC3 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC)) {
            return false;                                     // progressive, lossless, arithmetic: not decoded
        } else if (m == 0xC4) {                               // DHT
            size_t q = 0;
            while (q + 17 <= n) {
                const int tc = d[q] >> 4, th = d[q] & 15;
                if (tc > 1 || th > 3) return false;
                int total = 0;
                for (int i = 0; i < 16; i++) total += d[q + 1 + i];
                if (total > 256 || q + 17 + total > n) return false;
                (tc ? ac[th] : dc[th]).build(&d[q + 1], &d[q + 17], total);
                q += 17 + total;
            }
        } else if (m == 0xDD) {                               // DRI
            if (n < 2) return false;
            restart = (d[0] << 8) | d[1];
        } else if (m == 0xDA) {                               // SOS: the one scan of a sequential file
            if (ncomp == 0 || n < 1 || d[0] != ncomp || n < 1 + 2u * ncomp + 3) return false;
            for (int s = 0; s < ncomp; s++) {
                int c = 0;
                while (c < ncomp && comp[c].id != d[1 + 2 * s]) c++;
                if (c != s) return false;                     // components in frame order (what encoders write)
                comp[c].td = d[2 + 2 * s] >> 4; comp[c].ta = d[2 + 2 * s] & 15;
                if (comp[c].td > 3 || comp[c].ta > 3 || !dc[comp[c].td].set || !ac[comp[c].ta].set || !have_q[comp[c].tq]) return false;
            }
            int hmax = 1, vmax = 1;
            for (int c = 0; c < ncomp; c++) { hmax = std::max(hmax, comp[c].h); vmax = std::max(vmax, comp[c].v); }
            if (comp[0].h != hmax || comp[0].v != vmax) return false;      // Y at full resolution
            if (ncomp == 1) { comp[0].h = comp[0].v = 1; hmax = vmax = 1; } // a single-component scan is not interleaved
            const int mcu_w = 8 * hmax, mcu_h = 8 * vmax;
            const int mcus_x = (width + mcu_w - 1) / mcu_w, mcus_y = (height + mcu_h - 1) / mcu_h;
            const size_t pw = (size_t)mcus_x * mcu_w, ph = (size_t)mcus_y * mcu_h;
            std::vector<uint8_t> plane(pw * ph);
            JpegBits bits{&b[p + len], b.data() + b.size()};
            int todo = restart, next_rst = 0;
            for (int my = 0; my < mcus_y; my++)
                for (int mx = 0; mx < mcus_x; mx++) {
                    if (restart && todo == 0) {               // RSTn: byte-align, skip the marker, reset the predictors
                        bits.reset();
                        while (bits.p + 1 < bits.end && !(bits.p[0] == 0xFF && bits.p[1] >= 0xD0 && bits.p[1] <= 0xD7)) bits.p++;
                        if (bits.p + 1 >= bits.end || bits.p[1] != 0xD0 + next_rst) return false;
                        bits.p += 2;
                        next_rst = (next_rst + 1) & 7;
                        for (int c = 0; c < ncomp; c++) comp[c].pred = 0;
                        todo = restart;
                    }
                    for (int c = 0; c < ncomp; c++)
                        for (int by = 0; by < comp[c].v; by++)
                            for (int bx = 0; bx < comp[c].h; bx++) {
                                int coef[64] = {0};
                                const int s = bits.decode(dc[comp[c].td]);
                                comp[c].pred += jpeg_extend(bits.receive(s), s);
                                coef[0] = comp[c].pred * quant[comp[c].tq][0];
                                for (int k = 1; k < 64;) {
                                    const int rs = bits.decode(ac[comp[c].ta]), r = rs >> 4, sz = rs & 15;
                                    if (sz == 0) {
                                        if (r != 15) break;
                                        k += 16;
                                        continue;
                                    }
                                    k += r;
                                    if (k > 63) { bits.bad = true; break; }
                                    coef[ZZ[k]] = jpeg_extend(bits.receive(sz), sz) * quant[comp[c].tq][ZZ[k]];
                                    k++;
                                }
                                if (bits.bad) return false;
                                if (c == 0)
                                    jpeg_idct_islow(coef, &plane[((size_t)my * mcu_h + 8 * by) * pw + (size_t)mx * mcu_w + 8 * bx], pw);
                            }
                    if (restart) todo--;
                }
            img.cols = width; img.rows = height;
            img.data.resize((size_t)width * height);
            for (int y = 0; y < height; y++) std::memcpy(&img.data[(size_t)y * width], &plane[(size_t)y * pw], (size_t)width);
            return true;
        }
        p += len;
    }
    return false;
}

// cv::imread(name, 0): an empty image when the file cannot be read or decoded
inline Mat8u imread_grey(const std::string &name)
{
    Mat8u img;
    std::vector<uint8_t> bytes;
    if (!read_file(name, bytes) || bytes.size() < 8) return img;
    bool ok = false;
    if (bytes[0] == 'P' && bytes[1] == '5') ok = decode_pgm(bytes, img);
    else if (bytes[0] == 0xFF && bytes[1] == 0xD8) ok = decode_jpeg(bytes, img);
    else ok = decode_png(bytes, img);
    if (!ok) img = Mat8u();
    return img;
}

}  // namespace image_io
}  // namespace visgeom_b200
