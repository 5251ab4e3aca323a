// gi_jpg.cpp — JPEG decode without Qt: the other file format QImage(fname) reads for imageTexture
// (material.h:51-81).  Qt hands JPEG files to libjpeg with its defaults (quality >= 50: the accurate integer inverse DCT, "fancy"
// chroma upsampling, fixed-point YCbCr -> RGB), so a texture's bytes are defined by those three algorithms, and they are restated here
// in their published form (JPEG Annex A / F / G for the entropy coding; Loeffler-Ligtenberg-Moschytz 13-bit fixed-point IDCT;
// triangle-filter upsampling with libjpeg's rounding biases; 16-bit fixed-point colour tables) so that every pixel equals what
// QImage::pixelColor returns.  Checked byte for byte against an independent libjpeg build (PIL) in tests/test_jpg.py.
// Supported: 8-bit baseline / extended sequential and progressive Huffman JPEGs, 1 (grey) or 3 components (YCbCr, or RGB when an
// Adobe marker or the component ids say so), any sampling factors up to 4, restart intervals, multiple scans.  Not supported (the
// loader then falls back to the raw sidecar): arithmetic coding, 12-bit, lossless, CMYK / YCCK.  A JPEG has no alpha channel.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "gi_scene.hpp"

namespace {

const uint8_t kZigzag[64 + 16] = { 0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13,
                                   6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31,
                                   39, 46, 53, 60, 61, 54, 47, 55, 62, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63 };

struct Huff {
    bool present = false;
    uint8_t bits[17] = { 0 }, vals[256] = { 0 };
    int32_t maxcode[18];   // largest code of each length (-1: none), Annex F.2.2.3
    int32_t valptr[17];
    uint16_t mincode[17];
    uint16_t look[512];    // 9-bit prefix -> (length << 8) | value, 0 = longer code
    void build()
    {
        int code = 0, k = 0;
        for (int l = 1; l <= 16; l++) {
            valptr[l] = k; mincode[l] = (uint16_t)code;
            code += bits[l]; k += bits[l];
            maxcode[l] = bits[l] ? code - 1 : -1;
            code <<= 1;
        }
        maxcode[17] = 0x7fffffff;
        std::memset(look, 0, sizeof(look));
        code = 0; k = 0;
        for (int l = 1; l <= 9; l++) {
            for (int i = 0; i < bits[l]; i++, k++, code++) {
                const int first = code << (9 - l), n = 1 << (9 - l);
                for (int j = 0; j < n; j++) look[first + j] = (uint16_t)((l << 8) | vals[k]);
            }
            code <<= 1;
        }
        present = true;
    }
};

struct Comp {
    int id = 0, h = 1, v = 1, tq = 0;
    int bw = 0, bh = 0;          // blocks per row / column of the coefficient array (padded to whole MCUs)
    int dw = 0, dh = 0;          // downsampled size in samples: ceil(W * h / hmax), ceil(H * v / vmax)
    std::vector<int16_t> coef;   // bw * bh blocks of 64, natural order
    std::vector<uint8_t> plane;  // (bw * 8) x (bh * 8) samples after the inverse DCT
    int dc_pred = 0;
};

struct Bits {
    const uint8_t* p; const uint8_t* end;
    uint32_t acc = 0; int n = 0;
    bool hit_marker = false;
    void fill()
    {
        while (n <= 24) {
            int b = 0;
            if (!hit_marker && p < end) {
                b = *p++;
                if (b == 0xFF) {
                    if (p < end && *p == 0x00) p++;            // stuffed zero
                    else { hit_marker = true; p--; b = 0; }    // a marker: feed zeros from here on
                }
            }
            acc |= (uint32_t)b << (24 - n);
            n += 8;
        }
    }
    int peek(int k) { if (n < k) fill(); return (int)(acc >> (32 - k)); }
    void skip(int k) { acc <<= k; n -= k; }
    int get(int k) { if (k == 0) return 0; int v = peek(k); skip(k); return v; }
    int bit() { return get(1); }
    void reset() { acc = 0; n = 0; hit_marker = false; }
};

int decode_huff(Bits& b, const Huff& h)
{
    const int look = h.look[b.peek(9)];
    if (look) { b.skip(look >> 8); return look & 0xff; }
    int code = b.peek(16);
    for (int l = 10; l <= 16; l++) {
        const int c = code >> (16 - l);
        if (h.maxcode[l] >= 0 && c <= h.maxcode[l] && c >= h.mincode[l]) { b.skip(l); return h.vals[h.valptr[l] + c - h.mincode[l]]; }
    }
    b.skip(16);
    return -1;
}
inline int extend(int v, int s) { return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v; }   // Annex F.2.2.1

// ---- inverse DCT: Loeffler-Ligtenberg-Moschytz, 13-bit constants, 2 extra bits after the column pass (libjpeg "islow") ----------------
const int CB = 13, P1 = 2;
inline int32_t descale(int32_t x, int n) { return (x + (1 << (n - 1))) >> n; }
inline uint8_t clamp8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }
void idct_block(const int16_t* c, const uint16_t* q, uint8_t* out, int stride)
{
    const int32_t F0298 = 2446, F0390 = 3196, F0541 = 4433, F0765 = 6270, F0899 = 7373, F1175 = 9633, F1501 = 12299, F1847 = 15137, F1961 = 16069, F2053 = 16819,
                  F2562 = 20995, F3072 = 25172;
    int32_t ws[64];
    for (int x = 0; x < 8; x++) {
        int32_t in[8];
        for (int y = 0; y < 8; y++) in[y] = (int32_t)c[8 * y + x] * (int32_t)q[8 * y + x];
        int32_t z2 = in[2], z3 = in[6];
        int32_t z1 = (z2 + z3) * F0541;
        int32_t t2 = z1 + z3 * (-F1847), t3 = z1 + z2 * F0765;
        z2 = in[0]; z3 = in[4];
        int32_t t0 = (z2 + z3) * (1 << CB), t1 = (z2 - z3) * (1 << CB);
        const int32_t t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
        t0 = in[7]; t1 = in[5]; t2 = in[3]; t3 = in[1];
        z1 = t0 + t3; z2 = t1 + t2; z3 = t0 + t2; int32_t z4 = t1 + t3;
        const int32_t z5 = (z3 + z4) * F1175;
        t0 *= F0298; t1 *= F2053; t2 *= F3072; t3 *= F1501;
        z1 *= -F0899; z2 *= -F2562; z3 *= -F1961; z4 *= -F0390;
        z3 += z5; z4 += z5;
        t0 += z1 + z3; t1 += z2 + z4; t2 += z2 + z3; t3 += z1 + z4;
        ws[0 * 8 + x] = descale(t10 + t3, CB - P1); ws[7 * 8 + x] = descale(t10 - t3, CB - P1);
        ws[1 * 8 + x] = descale(t11 + t2, CB - P1); ws[6 * 8 + x] = descale(t11 - t2, CB - P1);
        ws[2 * 8 + x] = descale(t12 + t1, CB - P1); ws[5 * 8 + x] = descale(t12 - t1, CB - P1);
        ws[3 * 8 + x] = descale(t13 + t0, CB - P1); ws[4 * 8 + x] = descale(t13 - t0, CB - P1);
    }
    for (int y = 0; y < 8; y++) {
        const int32_t* w = ws + 8 * y;
        int32_t z2 = w[2], z3 = w[6];
        int32_t z1 = (z2 + z3) * F0541;
        int32_t t2 = z1 + z3 * (-F1847), t3 = z1 + z2 * F0765;
        int32_t t0 = (w[0] + w[4]) * (1 << CB), t1 = (w[0] - w[4]) * (1 << CB);
        const int32_t t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
        t0 = w[7]; t1 = w[5]; t2 = w[3]; t3 = w[1];
        z1 = t0 + t3; z2 = t1 + t2; z3 = t0 + t2; int32_t z4 = t1 + t3;
        const int32_t z5 = (z3 + z4) * F1175;
        t0 *= F0298; t1 *= F2053; t2 *= F3072; t3 *= F1501;
        z1 *= -F0899; z2 *= -F2562; z3 *= -F1961; z4 *= -F0390;
        z3 += z5; z4 += z5;
        t0 += z1 + z3; t1 += z2 + z4; t2 += z2 + z3; t3 += z1 + z4;
        uint8_t* o = out + (size_t)y * stride;
        const int S = CB + P1 + 3;
        o[0] = clamp8(descale(t10 + t3, S) + 128); o[7] = clamp8(descale(t10 - t3, S) + 128);
        o[1] = clamp8(descale(t11 + t2, S) + 128); o[6] = clamp8(descale(t11 - t2, S) + 128);
        o[2] = clamp8(descale(t12 + t1, S) + 128); o[5] = clamp8(descale(t12 - t1, S) + 128);
        o[3] = clamp8(descale(t13 + t0, S) + 128); o[4] = clamp8(descale(t13 - t0, S) + 128);
    }
}

struct Decoder {
    std::vector<uint8_t> file;
    int W = 0, H = 0, ncomp = 0, hmax = 1, vmax = 1;
    bool progressive = false, have_sof = false;
    Comp comp[4];
    uint16_t qt[4][64];
    bool have_qt[4] = { false, false, false, false };
    Huff hdc[4], hac[4];
    int restart = 0;
    int adobe_transform = -1; bool jfif = false;
    int eobrun = 0;

    static int be16(const uint8_t* p) { return (p[0] << 8) | p[1]; }

    bool parse()
    {
        const uint8_t* p = file.data(); const uint8_t* end = p + file.size();
        if (end - p < 4 || p[0] != 0xFF || p[1] != 0xD8) return false;
        p += 2;
        for (;;) {
            while (p < end && *p != 0xFF) p++;          // garbage before a marker is skipped
            while (p < end && *p == 0xFF) p++;          // fill bytes
            if (p >= end) return false;
            const int m = *p++;
            if (m == 0xD9) break;                        // EOI
            if (m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
            if (end - p < 2) return false;
            const int len = be16(p);
            if (len < 2 || p + len > end) return false;
            const uint8_t* s = p + 2; const uint8_t* se = p + len;
            p += len;
            switch (m) {
            case 0xDB:   // DQT
                while (s < se) {
                    const int pq = s[0] >> 4, tq = s[0] & 15; s++;
                    if (tq > 3 || pq > 1 || se - s < (pq ? 128 : 64)) return false;
                    for (int i = 0; i < 64; i++) { qt[tq][kZigzag[i]] = (uint16_t)(pq ? be16(s + 2 * i) : s[i]); }
                    s += pq ? 128 : 64; have_qt[tq] = true;
                }
                break;
            case 0xC4:   // DHT
                while (s < se) {
                    if (se - s < 17) return false;
                    const int tc = s[0] >> 4, th = s[0] & 15;
                    if (tc > 1 || th > 3) return false;
                    Huff& h = tc ? hac[th] : hdc[th];
                    int n = 0;
                    for (int l = 1; l <= 16; l++) { h.bits[l] = s[l]; n += s[l]; }
                    s += 17;
                    if (n > 256 || se - s < n) return false;
                    std::memcpy(h.vals, s, (size_t)n); s += n;
                    h.build();
                }
                break;
            case 0xC0: case 0xC1: case 0xC2: {   // SOF0 / 1 / 2
                if (have_sof || se - s < 6) return false;
                if (s[0] != 8) return false;   // sample precision
                H = be16(s + 1); W = be16(s + 3); ncomp = s[5];
                if (W <= 0 || H <= 0 || (ncomp != 1 && ncomp != 3) || se - s < 6 + 3 * ncomp) return false;
                for (int i = 0; i < ncomp; i++) {
                    Comp& c = comp[i];
                    c.id = s[6 + 3 * i]; c.h = s[7 + 3 * i] >> 4; c.v = s[7 + 3 * i] & 15; c.tq = s[8 + 3 * i];
                    if (c.h < 1 || c.h > 4 || c.v < 1 || c.v > 4 || c.tq > 3) return false;
                    if (c.h > hmax) hmax = c.h;
                    if (c.v > vmax) vmax = c.v;
                }
                if (ncomp == 1) { comp[0].h = comp[0].v = 1; hmax = vmax = 1; }   // a single component is never subsampled against itself
                const int mcux = (W + 8 * hmax - 1) / (8 * hmax), mcuy = (H + 8 * vmax - 1) / (8 * vmax);
                for (int i = 0; i < ncomp; i++) {
                    Comp& c = comp[i];
                    c.bw = mcux * c.h; c.bh = mcuy * c.v;
                    c.dw = (W * c.h + hmax - 1) / hmax; c.dh = (H * c.v + vmax - 1) / vmax;
                    c.coef.assign((size_t)c.bw * c.bh * 64, 0);
                }
                progressive = m == 0xC2; have_sof = true;
                break;
            }
            case 0xC3: case 0xC5: case 0xC6: case 0xC7: case 0xC9: case 0xCA: case 0xCB: case 0xCD: case 0xCE: case 0xCF:
                return false;   // lossless, hierarchical, arithmetic
            case 0xDD: if (se - s < 2) return false; restart = be16(s); break;
            case 0xE0: if (se - s >= 5 && std::memcmp(s, "JFIF", 5) == 0) jfif = true; break;
            case 0xEE: if (se - s >= 12 && std::memcmp(s, "Adobe", 5) == 0) adobe_transform = s[11]; break;
            case 0xDA: {   // SOS + entropy-coded data
                if (!have_sof || se - s < 1) return false;
                const int ns = s[0];
                if (ns < 1 || ns > ncomp || se - s < 1 + 2 * ns + 3) return false;
                int ci[4], td[4], ta[4];
                for (int i = 0; i < ns; i++) {
                    int k = -1;
                    for (int j = 0; j < ncomp; j++) if (comp[j].id == s[1 + 2 * i]) k = j;
                    if (k < 0) return false;
                    ci[i] = k; td[i] = s[2 + 2 * i] >> 4; ta[i] = s[2 + 2 * i] & 15;
                    if (td[i] > 3 || ta[i] > 3) return false;
                }
                const int ss = s[1 + 2 * ns], se_ = s[2 + 2 * ns], ah = s[3 + 2 * ns] >> 4, al = s[3 + 2 * ns] & 15;
                const uint8_t* q = p;
                if (!scan(q, end, ns, ci, td, ta, ss, se_, ah, al)) return false;
                p = q;
                break;
            }
            default: break;   // APPn, COM, DNL, ... skipped
            }
        }
        return have_sof;
    }

    // one block of a sequential scan (Annex F.2.2)
    bool block_baseline(Bits& b, Comp& c, int16_t* blk, const Huff& dc, const Huff& ac)
    {
        int t = decode_huff(b, dc);
        if (t < 0 || t > 15) return false;
        const int diff = t ? extend(b.get(t), t) : 0;
        c.dc_pred += diff;
        blk[0] = (int16_t)c.dc_pred;
        for (int k = 1; k < 64;) {
            const int rs = decode_huff(b, ac);
            if (rs < 0) return false;
            const int r = rs >> 4, sz = rs & 15;
            if (sz == 0) { if (r != 15) break; k += 16; continue; }
            k += r;
            if (k > 63) return false;
            blk[kZigzag[k]] = (int16_t)extend(b.get(sz), sz);
            k++;
        }
        return true;
    }
    // progressive scans (Annex G.1.2)
    bool block_dc_prog(Bits& b, Comp& c, int16_t* blk, const Huff& dc, int ah, int al)
    {
        if (ah == 0) {
            const int t = decode_huff(b, dc);
            if (t < 0 || t > 15) return false;
            const int diff = t ? extend(b.get(t), t) : 0;
            c.dc_pred += diff;
            blk[0] = (int16_t)(c.dc_pred * (1 << al));
        } else if (b.bit()) blk[0] = (int16_t)(blk[0] | (1 << al));
        return true;
    }
    bool block_ac_prog(Bits& b, int16_t* blk, const Huff& ac, int ss, int se, int ah, int al)
    {
        if (ah == 0) {
            if (eobrun) { eobrun--; return true; }
            for (int k = ss; k <= se;) {
                const int rs = decode_huff(b, ac);
                if (rs < 0) return false;
                const int r = rs >> 4, sz = rs & 15;
                if (sz == 0) {
                    if (r < 15) { eobrun = (1 << r) - 1; if (r) eobrun += b.get(r); break; }
                    k += 16;
                } else {
                    k += r;
                    if (k > 63) return false;
                    blk[kZigzag[k]] = (int16_t)(extend(b.get(sz), sz) * (1 << al));
                    k++;
                }
            }
            return true;
        }
        // refinement: every coefficient with history gets a correction bit; new +-1 coefficients are placed after `r` zero-history ones
        const int p1 = 1 << al, m1 = -(1 << al);
        int k = ss;
        if (eobrun == 0) {
            for (; k <= se;) {
                const int rs = decode_huff(b, ac);
                if (rs < 0) return false;
                int r = rs >> 4;
                const int sz = rs & 15;
                int val = 0;
                if (sz == 0) {
                    if (r < 15) { eobrun = (1 << r); if (r) eobrun += b.get(r); break; }
                } else {
                    if (sz != 1) return false;
                    val = b.bit() ? p1 : m1;
                }
                while (k <= se) {
                    int16_t* cf = blk + kZigzag[k++];
                    if (*cf != 0) {
                        if (b.bit() && (*cf & p1) == 0) *cf = (int16_t)(*cf >= 0 ? *cf + p1 : *cf + m1);
                    } else {
                        if (r == 0) { if (val) *cf = (int16_t)val; break; }
                        r--;
                    }
                }
            }
        }
        if (eobrun > 0) {
            for (; k <= se; k++) {
                int16_t* cf = blk + kZigzag[k];
                if (*cf != 0 && b.bit() && (*cf & p1) == 0) *cf = (int16_t)(*cf >= 0 ? *cf + p1 : *cf + m1);
            }
            eobrun--;
        }
        return true;
    }

    bool scan(const uint8_t*& p, const uint8_t* end, int ns, const int* ci, const int* td, const int* ta, int ss, int se, int ah, int al)
    {
        if (progressive) {
            if (ss > se || se > 63 || al > 13 || (ss == 0 && se != 0) || (ss > 0 && ns != 1)) return false;
        } else if (ss != 0 || se != 63 || ah != 0 || al != 0) return false;
        for (int i = 0; i < ns; i++) {
            if ((!progressive || ss == 0) && ah == 0 && !hdc[td[i]].present) return false;
            if ((!progressive || ss > 0) && !hac[ta[i]].present) return false;
            comp[ci[i]].dc_pred = 0;
        }
        eobrun = 0;
        Bits b; b.p = p; b.end = end;
        // a scan of one component covers ceil(dw / 8) x ceil(dh / 8) blocks; an interleaved scan whole MCUs
        const bool inter = ns > 1;
        const int mcux = inter ? (W + 8 * hmax - 1) / (8 * hmax) : (comp[ci[0]].dw + 7) / 8;
        const int mcuy = inter ? (H + 8 * vmax - 1) / (8 * vmax) : (comp[ci[0]].dh + 7) / 8;
        int until_restart = restart;
        for (int my = 0; my < mcuy; my++) {
            for (int mx = 0; mx < mcux; mx++) {
                if (restart && until_restart == 0) {
                    // byte-align, expect RSTn
                    b.reset();
                    const uint8_t* q = b.p;
                    while (q + 1 < end && !(q[0] == 0xFF && q[1] >= 0xD0 && q[1] <= 0xD7)) {
                        if (q[0] == 0xFF && q[1] != 0x00 && q[1] != 0xFF) return false;   // some other marker: the scan is short
                        q++;
                    }
                    if (q + 1 >= end) return false;
                    b.p = q + 2;
                    for (int i = 0; i < ns; i++) comp[ci[i]].dc_pred = 0;
                    eobrun = 0;
                    until_restart = restart;
                }
                for (int i = 0; i < ns; i++) {
                    Comp& c = comp[ci[i]];
                    const int nh = inter ? c.h : 1, nv = inter ? c.v : 1;
                    for (int by = 0; by < nv; by++)
                        for (int bx = 0; bx < nh; bx++) {
                            const int X = mx * nh + bx, Y = my * nv + by;
                            int16_t* blk = c.coef.data() + ((size_t)Y * c.bw + X) * 64;
                            bool ok;
                            if (!progressive) ok = block_baseline(b, c, blk, hdc[td[i]], hac[ta[i]]);
                            else if (ss == 0) ok = block_dc_prog(b, c, blk, hdc[td[i]], ah, al);
                            else ok = block_ac_prog(b, blk, hac[ta[i]], ss, se, ah, al);
                            if (!ok) return false;
                        }
                }
                if (restart) until_restart--;
            }
        }
        // the next marker: the bit reader stopped in front of it, or has not reached it yet (padding bits): look for it
        const uint8_t* q = b.p;
        while (q + 1 < end && !(q[0] == 0xFF && q[1] != 0x00 && q[1] != 0xFF && !(q[1] >= 0xD0 && q[1] <= 0xD7))) q++;
        p = q;
        return true;
    }

    void inverse_dct()
    {
        for (int i = 0; i < ncomp; i++) {
            Comp& c = comp[i];
            const int stride = c.bw * 8;
            c.plane.assign((size_t)stride * c.bh * 8, 0);
            for (int y = 0; y < c.bh; y++)
                for (int x = 0; x < c.bw; x++) idct_block(c.coef.data() + ((size_t)y * c.bw + x) * 64, qt[c.tq], c.plane.data() + (size_t)y * 8 * stride + x * 8, stride);
        }
    }

    // component i at full resolution, W x H.  Row context beyond the component's real rows repeats the edge row (what the decoder's
    // row buffers do); the horizontal edge cases of the triangle filter use the component's real width.
    void upsample(int i, std::vector<uint8_t>& out) const
    {
        const Comp& c = comp[i];
        const int hx = hmax / c.h, vx = vmax / c.v, stride = c.bw * 8;
        const bool integral = hmax % c.h == 0 && vmax % c.v == 0;
        out.assign((size_t)W * H, 0);
        auto row = [&](int y) { if (y < 0) y = 0; if (y > c.dh - 1) y = c.dh - 1; return c.plane.data() + (size_t)y * stride; };
        if (hx == 1 && vx == 1 && integral) {
            for (int y = 0; y < H; y++) std::memcpy(out.data() + (size_t)y * W, row(y), (size_t)W);
        } else if (integral && hx == 2 && vx == 1 && c.dw > 2) {   // h2v1, triangle filter
            std::vector<uint8_t> line((size_t)c.dw * 2);
            for (int y = 0; y < H; y++) {
                const uint8_t* in = row(y);
                const int n = c.dw;
                line[0] = in[0]; line[1] = (uint8_t)((in[0] * 3 + in[1] + 2) >> 2);
                for (int x = 1; x < n - 1; x++) { const int v = in[x] * 3; line[2 * x] = (uint8_t)((v + in[x - 1] + 1) >> 2); line[2 * x + 1] = (uint8_t)((v + in[x + 1] + 2) >> 2); }
                line[2 * n - 2] = (uint8_t)((in[n - 1] * 3 + in[n - 2] + 1) >> 2); line[2 * n - 1] = in[n - 1];
                std::memcpy(out.data() + (size_t)y * W, line.data(), (size_t)W);
            }
        } else if (integral && hx == 2 && vx == 2 && c.dw > 2) {   // h2v2, triangle filter in both directions
            std::vector<uint8_t> line((size_t)c.dw * 2);
            for (int y = 0; y < H; y++) {
                const int sy = y >> 1;
                const uint8_t* in0 = row(sy);
                const uint8_t* in1 = (y & 1) ? row(sy + 1) : row(sy - 1);   // the nearer neighbour row
                const int n = c.dw;
                int thiscol = in0[0] * 3 + in1[0], nextcol = in0[1] * 3 + in1[1], lastcol;
                line[0] = (uint8_t)((thiscol * 4 + 8) >> 4); line[1] = (uint8_t)((thiscol * 3 + nextcol + 7) >> 4);
                lastcol = thiscol; thiscol = nextcol;
                for (int x = 1; x < n - 1; x++) {
                    nextcol = in0[x + 1] * 3 + in1[x + 1];
                    line[2 * x] = (uint8_t)((thiscol * 3 + lastcol + 8) >> 4); line[2 * x + 1] = (uint8_t)((thiscol * 3 + nextcol + 7) >> 4);
                    lastcol = thiscol; thiscol = nextcol;
                }
                line[2 * n - 2] = (uint8_t)((thiscol * 3 + lastcol + 8) >> 4); line[2 * n - 1] = (uint8_t)((thiscol * 4 + 7) >> 4);
                std::memcpy(out.data() + (size_t)y * W, line.data(), (size_t)W);
            }
        } else if (integral && hx == 1 && vx == 2) {               // h1v2, triangle filter vertically
            for (int y = 0; y < H; y++) {
                const int sy = y >> 1;
                const uint8_t* in0 = row(sy);
                const uint8_t* in1 = (y & 1) ? row(sy + 1) : row(sy - 1);
                const int bias = (y & 1) ? 2 : 1;
                for (int x = 0; x < W; x++) out[(size_t)y * W + x] = (uint8_t)((in0[x] * 3 + in1[x] + bias) >> 2);
            }
        } else if (integral) {                                      // any other integral factor: replication
            for (int y = 0; y < H; y++) {
                const uint8_t* in = row(y / vx);
                for (int x = 0; x < W; x++) out[(size_t)y * W + x] = in[x / hx];
            }
        } else {                                                    // fractional factors (e.g. 3:2): nearest sample of the scaled grid
            for (int y = 0; y < H; y++) {
                const uint8_t* in = row(y * c.v / vmax);
                for (int x = 0; x < W; x++) out[(size_t)y * W + x] = in[x * c.h / hmax];
            }
        }
    }
};

}  // namespace

bool gi_jpg_decode(const char* path, int& width, int& height, std::vector<uint8_t>& rgba)
{
    FILE* f = std::fopen(path, "rb");
    if (!f) return false;
    Decoder D;
    std::fseek(f, 0, SEEK_END);
    const long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    if (sz < 4) { std::fclose(f); return false; }
    D.file.resize((size_t)sz);
    const bool rd = std::fread(D.file.data(), 1, (size_t)sz, f) == (size_t)sz;
    std::fclose(f);
    if (!rd || !D.parse()) return false;
    for (int i = 0; i < D.ncomp; i++) if (!D.have_qt[D.comp[i].tq]) return false;
    D.inverse_dct();
    width = D.W; height = D.H;
    rgba.assign((size_t)D.W * D.H * 4, 255);
    const size_t n = (size_t)D.W * D.H;
    if (D.ncomp == 1) {
        std::vector<uint8_t> y; D.upsample(0, y);
        for (size_t i = 0; i < n; i++) { rgba[4 * i] = rgba[4 * i + 1] = rgba[4 * i + 2] = y[i]; }
        return true;
    }
    std::vector<uint8_t> c0, c1, c2;
    D.upsample(0, c0); D.upsample(1, c1); D.upsample(2, c2);
    // colour space (libjpeg's rules): an Adobe marker decides (transform 0 = RGB, 1 = YCbCr); else JFIF means YCbCr; else the component
    // ids 'R','G','B' mean RGB; else YCbCr
    bool ycc = true;
    if (D.adobe_transform >= 0) ycc = D.adobe_transform != 0;
    else if (!D.jfif && D.comp[0].id == 'R' && D.comp[1].id == 'G' && D.comp[2].id == 'B') ycc = false;
    if (!ycc) {
        for (size_t i = 0; i < n; i++) { rgba[4 * i] = c0[i]; rgba[4 * i + 1] = c1[i]; rgba[4 * i + 2] = c2[i]; }
        return true;
    }
    // YCbCr -> RGB with 16-bit fixed-point tables: R = Y + 1.402 Cr', G = Y - 0.34414 Cb' - 0.71414 Cr', B = Y + 1.772 Cb'
    int32_t cr_r[256], cb_b[256], cr_g[256], cb_g[256];
    const int32_t half = 1 << 15;
    auto fix = [](double x) { return (int32_t)(x * 65536.0 + 0.5); };
    for (int i = 0; i < 256; i++) {
        const int32_t x = i - 128;
        cr_r[i] = (fix(1.40200) * x + half) >> 16;
        cb_b[i] = (fix(1.77200) * x + half) >> 16;
        cr_g[i] = -fix(0.71414) * x;
        cb_g[i] = -fix(0.34414) * x + half;
    }
    for (size_t i = 0; i < n; i++) {
        const int y = c0[i], cb = c1[i], cr = c2[i];
        rgba[4 * i] = clamp8(y + cr_r[cr]);
        rgba[4 * i + 1] = clamp8(y + ((cb_g[cb] + cr_g[cr]) >> 16));
        rgba[4 * i + 2] = clamp8(y + cb_b[cb]);
    }
    return true;
}

extern "C" {
// test / tool entry point: decode into caller memory (rgba may be NULL to query the size)
int gih_jpg_decode(const char* path, int* w, int* h, uint8_t* rgba, size_t cap)
{
    int ww = 0, hh = 0;
    std::vector<uint8_t> px;
    if (!gi_jpg_decode(path, ww, hh, px)) return GI_ERR_INVALID;
    if (w) *w = ww;
    if (h) *h = hh;
    if (rgba) { if (cap < px.size()) return GI_ERR_INVALID; std::memcpy(rgba, px.data(), px.size()); }
    return GI_OK;
}
}
