// gi_png.cpp — PNG decode / encode without Qt: what QImage did for imageTexture (material.h:51-81: load a
// file, width / height / hasAlphaChannel / pixelColor) and for saving the frame (gui.h:39-45).  zlib does inflate / deflate /
// crc32; chunk parsing, the five scanline filters, Adam7 and the conversion to 8-bit RGBA are here.
// Decoded pixels are what QImage::pixelColor(x, y).red()/green()/blue()/alpha() return for the same file: grey expands to
// r = g = b, palette entries come from PLTE (+ tRNS alpha), sub-byte depths are scaled to 0..255, 16-bit samples keep their
// high byte, a tRNS colour key makes matching pixels transparent.  has_alpha = colour type 4 / 6 or a tRNS chunk
// (QImage::hasAlphaChannel).  Gamma / colour-profile chunks are ignored (Qt applies gAMA only on request).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <zlib.h>

#include "gi_scene.hpp"

namespace {

uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
void put32(std::vector<uint8_t>& v, uint32_t x) { v.push_back((uint8_t)(x >> 24)); v.push_back((uint8_t)(x >> 16)); v.push_back((uint8_t)(x >> 8)); v.push_back((uint8_t)x); }

int paeth(int a, int b, int c)
{
    int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// undo the scanline filter of one pass: `rows` scanlines of `stride` bytes each, every line preceded by its filter byte
bool unfilter(uint8_t* d, size_t rows, size_t stride, size_t bpp)
{
    std::vector<uint8_t> zero(stride, 0);
    const uint8_t* prev = zero.data();
    for (size_t y = 0; y < rows; y++) {
        uint8_t ft = d[0];
        uint8_t* cur = d + 1;
        if (ft > 4) return false;
        for (size_t i = 0; i < stride; i++) {
            int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0, x = cur[i];
            switch (ft) {
            case 1: x += a; break;
            case 2: x += b; break;
            case 3: x += (a + b) >> 1; break;
            case 4: x += paeth(a, b, c); break;
            default: break;
            }
            cur[i] = (uint8_t)x;
        }
        prev = cur;
        d += stride + 1;
    }
    return true;
}

}  // namespace

bool gi_png_decode(const char* path, int& width, int& height, bool& has_alpha, std::vector<uint8_t>& rgba)
{
    FILE* f = std::fopen(path, "rb");
    if (!f) return false;
    std::vector<uint8_t> file;
    uint8_t buf[65536];
    size_t n;
    while ((n = std::fread(buf, 1, sizeof(buf), f)) > 0) file.insert(file.end(), buf, buf + n);
    std::fclose(f);
    static const uint8_t sig[8] = { 0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A };
    if (file.size() < 8 + 25 || std::memcmp(file.data(), sig, 8) != 0) return false;
    uint32_t w = 0, h = 0;
    int depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, plte, trns;
    bool have_ihdr = false, have_trns = false, have_end = false;
    for (size_t pos = 8; pos + 12 <= file.size();) {
        uint32_t len = be32(&file[pos]);
        if ((uint64_t)pos + 12ull + len > file.size()) return false;
        const uint8_t* type = &file[pos + 4];
        const uint8_t* data = &file[pos + 8];
        if (be32(data + len) != (uint32_t)crc32(crc32(0L, Z_NULL, 0), type, len + 4)) return false;   // chunk CRC covers type + data
        if (!std::memcmp(type, "IHDR", 4)) {
            if (len != 13) return false;
            w = be32(data); h = be32(data + 4); depth = data[8]; ctype = data[9]; interlace = data[12];
            if (data[10] != 0 || data[11] != 0 || interlace > 1) return false;
            have_ihdr = true;
        } else if (!std::memcmp(type, "PLTE", 4)) plte.assign(data, data + len);
        else if (!std::memcmp(type, "tRNS", 4)) { trns.assign(data, data + len); have_trns = true; }
        else if (!std::memcmp(type, "IDAT", 4)) idat.insert(idat.end(), data, data + len);
        else if (!std::memcmp(type, "IEND", 4)) { have_end = true; break; }
        pos += 12 + (size_t)len;
    }
    if (!have_ihdr || !have_end || w == 0 || h == 0 || w > 65535u || h > 65535u) return false;
    int channels;
    switch (ctype) {
    case 0: channels = 1; break;
    case 2: channels = 3; break;
    case 3: channels = 1; break;
    case 4: channels = 2; break;
    case 6: channels = 4; break;
    default: return false;
    }
    const bool depth_ok = (ctype == 0 && (depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)) || (ctype == 3 && (depth == 1 || depth == 2 || depth == 4 || depth == 8)) ||
                          ((ctype == 2 || ctype == 4 || ctype == 6) && (depth == 8 || depth == 16));
    if (!depth_ok || (ctype == 3 && plte.size() < 3)) return false;
    const size_t bits_pp = (size_t)channels * depth, bpp = (bits_pp + 7) / 8;
    // pass geometry: one pass, or the seven Adam7 passes
    static const int px0[7] = { 0, 4, 0, 2, 0, 1, 0 }, py0[7] = { 0, 0, 4, 0, 2, 0, 1 }, pdx[7] = { 8, 8, 4, 4, 2, 2, 1 }, pdy[7] = { 8, 8, 8, 4, 4, 2, 2 };
    const int passes = interlace ? 7 : 1;
    size_t total = 0;
    size_t pw[7], ph[7];
    for (int p = 0; p < passes; p++) {
        pw[p] = interlace ? (w + pdx[p] - 1 - px0[p]) / pdx[p] : w;
        ph[p] = interlace ? (h + pdy[p] - 1 - py0[p]) / pdy[p] : h;
        if (pw[p] && ph[p]) total += ph[p] * (1 + (pw[p] * bits_pp + 7) / 8);
    }
    std::vector<uint8_t> raw(total);
    uLongf out_len = (uLongf)total;
    int zr = uncompress(raw.data(), &out_len, idat.data(), (uLong)idat.size());
    if (zr != Z_OK || out_len != total) return false;
    width = (int)w; height = (int)h;
    has_alpha = ctype == 4 || ctype == 6 || have_trns;
    rgba.assign((size_t)w * h * 4, 255);
    const int maxv = (1 << (depth > 8 ? 8 : depth)) - 1;
    auto sample = [&](const uint8_t* line, size_t idx) -> int {   // idx-th sample of the line, reduced to 8 bits (16-bit: high byte)
        if (depth == 8) return line[idx];
        if (depth == 16) return line[idx * 2];
        const size_t bit = idx * depth;
        return (line[bit >> 3] >> (8 - depth - (bit & 7))) & maxv;
    };
    auto raw16 = [&](const uint8_t* line, size_t idx) -> int { return depth == 16 ? (line[idx * 2] << 8) | line[idx * 2 + 1] : sample(line, idx); };
    size_t off = 0;
    for (int p = 0; p < passes; p++) {
        if (!pw[p] || !ph[p]) continue;
        const size_t stride = (pw[p] * bits_pp + 7) / 8;
        if (!unfilter(&raw[off], ph[p], stride, bpp)) return false;
        for (size_t yy = 0; yy < ph[p]; yy++) {
            const uint8_t* line = &raw[off + yy * (stride + 1) + 1];
            const size_t y = interlace ? (size_t)py0[p] + yy * pdy[p] : yy;
            for (size_t xx = 0; xx < pw[p]; xx++) {
                const size_t x = interlace ? (size_t)px0[p] + xx * pdx[p] : xx;
                uint8_t* o = &rgba[(y * w + x) * 4];
                if (ctype == 3) {
                    size_t k = (size_t)sample(line, xx);
                    if (k * 3 + 2 >= plte.size()) k = 0;   // index past the palette: libpng reports an error; entry 0 here
                    o[0] = plte[k * 3]; o[1] = plte[k * 3 + 1]; o[2] = plte[k * 3 + 2];
                    o[3] = k < trns.size() ? trns[k] : 255;
                } else if (ctype == 0 || ctype == 4) {
                    const int g = sample(line, xx * channels);
                    const uint8_t g8 = (uint8_t)(depth < 8 ? g * 255 / maxv : g);
                    o[0] = o[1] = o[2] = g8;
                    if (ctype == 4) o[3] = (uint8_t)sample(line, xx * 2 + 1);
                    else if (have_trns && trns.size() >= 2 && raw16(line, xx) == ((trns[0] << 8) | trns[1])) o[3] = 0;
                } else {
                    o[0] = (uint8_t)sample(line, xx * channels); o[1] = (uint8_t)sample(line, xx * channels + 1); o[2] = (uint8_t)sample(line, xx * channels + 2);
                    if (ctype == 6) o[3] = (uint8_t)sample(line, xx * 4 + 3);
                    else if (have_trns && trns.size() >= 6 && raw16(line, xx * 3) == ((trns[0] << 8) | trns[1]) && raw16(line, xx * 3 + 1) == ((trns[2] << 8) | trns[3]) &&
                             raw16(line, xx * 3 + 2) == ((trns[4] << 8) | trns[5]))
                        o[3] = 0;
                }
            }
        }
        off += ph[p] * (stride + 1);
    }
    return true;
}

// 8-bit RGB, filter 0 on every line, one IDAT
bool gi_png_encode(const char* path, int width, int height, const uint8_t* rgb)
{
    if (width <= 0 || height <= 0 || !rgb) return false;
    const size_t stride = (size_t)width * 3;
    std::vector<uint8_t> raw((stride + 1) * (size_t)height);
    for (int y = 0; y < height; y++) {
        raw[(stride + 1) * y] = 0;
        std::memcpy(&raw[(stride + 1) * y + 1], rgb + stride * y, stride);
    }
    uLongf clen = compressBound((uLong)raw.size());
    std::vector<uint8_t> comp(clen);
    if (compress2(comp.data(), &clen, raw.data(), (uLong)raw.size(), 6) != Z_OK) return false;
    comp.resize(clen);
    std::vector<uint8_t> out = { 0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A };
    auto chunk = [&](const char* type, const std::vector<uint8_t>& data) {
        put32(out, (uint32_t)data.size());
        size_t start = out.size();
        out.insert(out.end(), type, type + 4);
        out.insert(out.end(), data.begin(), data.end());
        put32(out, (uint32_t)crc32(crc32(0L, Z_NULL, 0), &out[start], (uInt)(out.size() - start)));
    };
    std::vector<uint8_t> ihdr;
    put32(ihdr, (uint32_t)width); put32(ihdr, (uint32_t)height);
    ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    chunk("IHDR", ihdr);
    chunk("IDAT", comp);
    chunk("IEND", {});
    FILE* f = std::fopen(path, "wb");
    if (!f) return false;
    bool ok = std::fwrite(out.data(), 1, out.size(), f) == out.size();
    std::fclose(f);
    return ok;
}

bool Image::writePNG(const char* path) const { return gi_png_encode(path, _w, _h, rgb.data()); }

extern "C" {
// test / tool entry points: decode into caller memory (rgba may be NULL to query the size), encode from caller memory
int gih_png_decode(const char* path, int* w, int* h, int* has_alpha, uint8_t* rgba, size_t cap)
{
    int ww = 0, hh = 0; bool a = false;
    std::vector<uint8_t> px;
    if (!gi_png_decode(path, ww, hh, a, px)) return GI_ERR_INVALID;
    if (w) *w = ww;
    if (h) *h = hh;
    if (has_alpha) *has_alpha = a ? 1 : 0;
    if (rgba) { if (cap < px.size()) return GI_ERR_INVALID; std::memcpy(rgba, px.data(), px.size()); }
    return GI_OK;
}
int gih_png_encode(const char* path, int w, int h, const uint8_t* rgb) { return gi_png_encode(path, w, h, rgb) ? GI_OK : GI_ERR_INVALID; }
}
