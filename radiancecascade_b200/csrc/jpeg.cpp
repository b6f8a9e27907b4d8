// jpeg.cpp — JPEG decode for map_Kd / map_Bump textures: the stand-in for
// image::ImageReader::open(p).decode() (src/primitives.rs:391-404) when the file is a JPEG.
//
// Baseline / extended-sequential (SOF0, SOF1) and progressive (SOF2) Huffman JPEG, 8-bit, 1 or 3
// components, restart intervals.  The sample pipeline follows the integer reference pipeline of the
// Independent JPEG Group's decoder, which image decoders are commonly validated against: 13-bit
// fixed-point "islow" inverse DCT, triangle-filter ("fancy") chroma upsampling for 2x1 / 1x2 / 2x2,
// 16-bit fixed-point YCbCr -> RGB.  (The reference's decoder, zune-jpeg 0.4.13, is not vendored; a
// +-1 LSB IDCT/upsampling difference between JPEG decoders is expected)  Arithmetic
// coding, 12-bit samples, CMYK and hierarchical files are rejected: the caller then warns and uses the
// empty texture exactly as the reference does for an undecodable file (src/renderer.rs:424-430).
#include <cstdint>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "scene.h"

namespace rc {
namespace {

const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct Huff {
    bool present = false;
    uint8_t bits[17] = {0};
    uint8_t vals[256] = {0};
    int mincode[18], maxcode[18], valptr[18];
    int16_t fast[512];   // 9-bit lookup: (len << 8) | symbol, or -1
    void build()
    {
        int code = 0, k = 0;
        for (int l = 1; l <= 16; l++) {
            valptr[l] = k;
            mincode[l] = code;
            code += bits[l];
            k += bits[l];
            maxcode[l] = bits[l] ? code - 1 : -1;
            code <<= 1;
        }
        maxcode[17] = 0x7fffffff;
        for (auto& f : fast) f = -1;
        code = 0; k = 0;
        for (int l = 1; l <= 9; l++) {
            for (int i = 0; i < bits[l]; i++, k++, code++) {
                const int first = code << (9 - l), n = 1 << (9 - l);
                for (int j = 0; j < n; j++) fast[first + j] = (int16_t)((l << 8) | vals[k]);
            }
            code <<= 1;
        }
    }
};

struct BitReader {
    const uint8_t* p;
    const uint8_t* end;
    uint32_t buf = 0;
    int cnt = 0;
    bool hit_marker = false;
    void fill()
    {
        while (cnt <= 24) {
            uint32_t b = 0;
            if (!hit_marker && p < end) {
                b = *p;
                if (b == 0xff) {
                    const uint8_t n = p + 1 < end ? p[1] : 0xd9;
                    if (n == 0) p += 2;              // stuffed zero
                    else { hit_marker = true; b = 0; }   // a marker ends the entropy segment: feed zeros
                } else p++;
            }
            buf |= b << (24 - cnt);
            cnt += 8;
        }
    }
    inline int peek(int n) { if (cnt < n) fill(); return (int)(buf >> (32 - n)); }
    inline void skip(int n) { buf <<= n; cnt -= n; }
    inline int get(int n) { if (n == 0) return 0; const int v = peek(n); skip(n); return v; }
    inline int bit() { return get(1); }
    void reset() { buf = 0; cnt = 0; hit_marker = false; }
};

inline int extend(int v, int s) { return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v; }   // callers guarantee 1 <= s <= 15

int decode_symbol(BitReader& br, const Huff& h)
{
    const int look = br.peek(9);
    const int f = h.fast[look];
    if (f >= 0) { br.skip(f >> 8); return f & 0xff; }
    int code = br.peek(16), l;
    for (l = 10; l <= 16; l++) if ((code >> (16 - l)) <= h.maxcode[l]) break;
    if (l > 16) { br.skip(16); return 0; }
    br.skip(l);
    const int c = code >> (16 - l);
    return h.vals[(h.valptr[l] + c - h.mincode[l]) & 0xff];
}

struct Component {
    int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
    int bw = 0, bh = 0;        // blocks per line / column, padded to whole MCUs
    int cw = 0, chh = 0;       // real blocks per line / column (non-interleaved scans)
    int sw = 0, sh = 0;        // real samples per line / column
    int pred = 0;
    std::vector<int16_t> coef; // bw*bh*64
    std::vector<uint8_t> pix;  // (bw*8) x (bh*8)
};

struct Decoder {
    const uint8_t* data;
    size_t size;
    int width = 0, height = 0, ncomp = 0, hmax = 1, vmax = 1;
    bool progressive = false;
    uint16_t qt[4][64];
    bool qt_present[4] = {false, false, false, false};
    Huff dc[4], ac[4];
    Component comp[3];
    int restart_interval = 0;
    int mcus_x = 0, mcus_y = 0;
    int adobe_transform = -1;
    std::string err;

    bool fail(const char* m) { err = m; return false; }
    bool bad_stream = false;   // set by the entropy decoders on a symbol no valid 8-bit stream can contain
    bool corrupt(const char* m) { if (!bad_stream) { bad_stream = true; err = m; } return false; }

    // ---- entropy decoding of one block ----------------------------------------------
    bool block_baseline(BitReader& br, Component& c, int16_t* b)
    {
        const int s = decode_symbol(br, dc[c.td]);
        if (s > 11) return corrupt("DC coefficient category above 11");          // 8-bit JPEG: DC differences need <= 11 bits
        const int diff = s ? extend(br.get(s), s) : 0;
        c.pred += diff;
        b[0] = (int16_t)c.pred;
        for (int k = 1; k < 64;) {
            const int rs = decode_symbol(br, ac[c.ta]);
            const int r = rs >> 4, sz = rs & 15;
            if (sz == 0) { if (r == 15) { k += 16; continue; } break; }
            if (sz > 10) return corrupt("AC coefficient category above 10");
            k += r;
            if (k > 63) break;
            b[kZigzag[k]] = (int16_t)extend(br.get(sz), sz);
            k++;
        }
        return true;
    }

    void block_dc_first(BitReader& br, Component& c, int16_t* b, int al)
    {
        const int s = decode_symbol(br, dc[c.td]);
        if (s > 11) { corrupt("DC coefficient category above 11"); return; }
        const int diff = s ? extend(br.get(s), s) : 0;
        c.pred += diff;
        b[0] = (int16_t)(c.pred * (1 << al));
    }
    void block_dc_refine(BitReader& br, int16_t* b, int al) { if (br.bit()) b[0] |= (int16_t)(1 << al); }

    void block_ac_first(BitReader& br, Component& c, int16_t* b, int ss, int se, int al, int& eobrun)
    {
        if (eobrun > 0) { eobrun--; return; }
        for (int k = ss; k <= se;) {
            const int rs = decode_symbol(br, ac[c.ta]);
            const int r = rs >> 4, s = rs & 15;
            if (s == 0) {
                if (r < 15) { eobrun = (1 << r) - 1; if (r) eobrun += br.get(r); break; }
                k += 16;
            } else {
                if (s > 10) { corrupt("AC coefficient category above 10"); return; }
                k += r;
                if (k > 63) break;
                b[kZigzag[k]] = (int16_t)(extend(br.get(s), s) * (1 << al));
                k++;
            }
        }
    }

    void block_ac_refine(BitReader& br, Component& c, int16_t* b, int ss, int se, int al, int& eobrun)
    {
        const int p1 = 1 << al, m1 = -(1 << al);
        int k = ss;
        if (eobrun == 0) {
            for (; k <= se; k++) {
                const int rs = decode_symbol(br, ac[c.ta]);
                int r = rs >> 4, s = rs & 15;
                if (s) s = br.bit() ? p1 : m1;
                else if (r != 15) { eobrun = 1 << r; if (r) eobrun += br.get(r); break; }
                while (k <= se) {
                    int16_t* cp = &b[kZigzag[k]];
                    if (*cp != 0) {
                        if (br.bit() && (*cp & p1) == 0) *cp = (int16_t)(*cp + (*cp >= 0 ? p1 : m1));
                    } else {
                        if (--r < 0) break;
                    }
                    k++;
                }
                if (s && k <= 63) b[kZigzag[k]] = (int16_t)s;
            }
        }
        if (eobrun > 0) {
            for (; k <= se; k++) {
                int16_t* cp = &b[kZigzag[k]];
                if (*cp != 0 && br.bit() && (*cp & p1) == 0) *cp = (int16_t)(*cp + (*cp >= 0 ? p1 : m1));
            }
            eobrun--;
        }
    }

    // ---- one scan ---------------------------------------------------------------------
    bool decode_scan(const uint8_t*& p, const uint8_t* end, int ns, const int* scan_comp, int ss, int se, int ah, int al)
    {
        BitReader br;
        br.p = p; br.end = end;
        for (int i = 0; i < ns; i++) comp[scan_comp[i]].pred = 0;
        int eobrun = 0, restarts = restart_interval;
        auto do_block = [&](Component& c, int bx, int by) {
            int16_t* b = &c.coef[((size_t)by * c.bw + bx) * 64];
            if (!progressive) block_baseline(br, c, b);
            else if (ss == 0) { if (ah == 0) block_dc_first(br, c, b, al); else block_dc_refine(br, b, al); }
            else { if (ah == 0) block_ac_first(br, c, b, ss, se, al, eobrun); else block_ac_refine(br, c, b, ss, se, al, eobrun); }
        };
        auto restart = [&]() {
            // byte-align, expect RSTn, reset predictors
            br.reset();
            const uint8_t* q = br.p;
            while (q + 1 < end && !(q[0] == 0xff && q[1] >= 0xd0 && q[1] <= 0xd7)) {
                if (q[0] == 0xff && q[1] != 0 && q[1] != 0xff) break;
                q++;
            }
            if (q + 1 < end && q[0] == 0xff && q[1] >= 0xd0 && q[1] <= 0xd7) q += 2;
            br.p = q;
            for (int i = 0; i < ns; i++) comp[scan_comp[i]].pred = 0;
            eobrun = 0;
            restarts = restart_interval;
        };
        if (ns == 1) {
            Component& c = comp[scan_comp[0]];
            for (int by = 0; by < c.chh; by++)
                for (int bx = 0; bx < c.cw; bx++) {
                    if (restart_interval && restarts == 0) restart();
                    do_block(c, bx, by);
                    restarts--;
                }
        } else {
            for (int my = 0; my < mcus_y; my++)
                for (int mx = 0; mx < mcus_x; mx++) {
                    if (restart_interval && restarts == 0) restart();
                    for (int i = 0; i < ns; i++) {
                        Component& c = comp[scan_comp[i]];
                        for (int v = 0; v < c.v; v++)
                            for (int h = 0; h < c.h; h++) do_block(c, mx * c.h + h, my * c.v + v);
                    }
                    restarts--;
                }
        }
        // advance to the next marker
        const uint8_t* q = br.p;
        while (q + 1 < end && !(q[0] == 0xff && q[1] != 0 && q[1] != 0xff && !(q[1] >= 0xd0 && q[1] <= 0xd7))) q++;
        p = q;
        return !bad_stream;
    }

    // ---- islow inverse DCT (13-bit fixed point, two passes) ------------------------------
    static inline int descale(long x, int n) { return (int)((x + (1L << (n - 1))) >> n); }
    static inline uint8_t range_limit(int x)
    {
        x = ((x + 512) & 1023) - 512;   // the reference pipeline masks to 10 bits before its clamp table
        x += 128;
        return (uint8_t)(x < 0 ? 0 : (x > 255 ? 255 : x));
    }
    static void idct(const int16_t* in, const uint16_t* q, uint8_t* out, int stride)
    {
        constexpr long F0_298 = 2446, F0_390 = 3196, F0_541 = 4433, F0_765 = 6270, F0_899 = 7373, F1_175 = 9633, F1_501 = 12299,
                       F1_847 = 15137, F1_961 = 16069, F2_053 = 16819, F2_562 = 20995, F3_072 = 25172;
        constexpr int CB = 13, P1 = 2;
        long ws[64];
        for (int c = 0; c < 8; c++) {
            const int16_t* i = in + c;
            const uint16_t* qq = q + c;
            if (!(i[8] | i[16] | i[24] | i[32] | i[40] | i[48] | i[56])) {
                const long dcv = ((long)i[0] * qq[0]) * (1L << P1);
                for (int r = 0; r < 8; r++) ws[r * 8 + c] = dcv;
                continue;
            }
            long z2 = (long)i[16] * qq[16], z3 = (long)i[48] * qq[48];
            long z1 = (z2 + z3) * F0_541;
            long tmp2 = z1 + z3 * (-F1_847), tmp3 = z1 + z2 * F0_765;
            z2 = (long)i[0] * qq[0]; z3 = (long)i[32] * qq[32];
            long tmp0 = (z2 + z3) * (1L << CB), tmp1 = (z2 - z3) * (1L << CB);
            const long tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
            tmp0 = (long)i[56] * qq[56]; tmp1 = (long)i[40] * qq[40]; tmp2 = (long)i[24] * qq[24]; tmp3 = (long)i[8] * qq[8];
            z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
            long z4 = tmp1 + tmp3;
            const long z5 = (z3 + z4) * F1_175;
            tmp0 *= F0_298; tmp1 *= F2_053; tmp2 *= F3_072; tmp3 *= F1_501;
            z1 *= -F0_899; z2 *= -F2_562; z3 *= -F1_961; z4 *= -F0_390;
            z3 += z5; z4 += z5;
            tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
            ws[0 * 8 + c] = descale(tmp10 + tmp3, CB - P1); ws[7 * 8 + c] = descale(tmp10 - tmp3, CB - P1);
            ws[1 * 8 + c] = descale(tmp11 + tmp2, CB - P1); ws[6 * 8 + c] = descale(tmp11 - tmp2, CB - P1);
            ws[2 * 8 + c] = descale(tmp12 + tmp1, CB - P1); ws[5 * 8 + c] = descale(tmp12 - tmp1, CB - P1);
            ws[3 * 8 + c] = descale(tmp13 + tmp0, CB - P1); ws[4 * 8 + c] = descale(tmp13 - tmp0, CB - P1);
        }
        for (int r = 0; r < 8; r++) {
            const long* w = ws + r * 8;
            uint8_t* o = out + (size_t)r * stride;
            long z2 = w[2], z3 = w[6];
            long z1 = (z2 + z3) * F0_541;
            long tmp2 = z1 + z3 * (-F1_847), tmp3 = z1 + z2 * F0_765;
            long tmp0 = (w[0] + w[4]) * (1L << CB), tmp1 = (w[0] - w[4]) * (1L << CB);
            const long tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
            tmp0 = w[7]; tmp1 = w[5]; tmp2 = w[3]; tmp3 = w[1];
            z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
            long z4 = tmp1 + tmp3;
            const long z5 = (z3 + z4) * F1_175;
            tmp0 *= F0_298; tmp1 *= F2_053; tmp2 *= F3_072; tmp3 *= F1_501;
            z1 *= -F0_899; z2 *= -F2_562; z3 *= -F1_961; z4 *= -F0_390;
            z3 += z5; z4 += z5;
            tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
            constexpr int S = CB + P1 + 3;
            o[0] = range_limit(descale(tmp10 + tmp3, S)); o[7] = range_limit(descale(tmp10 - tmp3, S));
            o[1] = range_limit(descale(tmp11 + tmp2, S)); o[6] = range_limit(descale(tmp11 - tmp2, S));
            o[2] = range_limit(descale(tmp12 + tmp1, S)); o[5] = range_limit(descale(tmp12 - tmp1, S));
            o[3] = range_limit(descale(tmp13 + tmp0, S)); o[4] = range_limit(descale(tmp13 - tmp0, S));
        }
    }

    // ---- header parsing -------------------------------------------------------------------
    bool run(Image& out)
    {
        const uint8_t* p = data + 2;
        const uint8_t* end = data + size;
        bool have_frame = false;
        while (p + 4 <= end) {
            if (p[0] != 0xff) { p++; continue; }
            const uint8_t m = p[1];
            if (m == 0xff) { p++; continue; }
            if (m == 0xd9) break;                                  // EOI
            if (m == 0x01 || (m >= 0xd0 && m <= 0xd7)) { p += 2; continue; }
            const int len = (p[2] << 8) | p[3];
            const uint8_t* s = p + 4;
            const uint8_t* se_ = p + 2 + len;
            if (se_ > end || len < 2) return fail("truncated JPEG segment");
            if (m == 0xdb) {                                       // DQT
                while (s < se_) {
                    const int pq = s[0] >> 4, tq = s[0] & 15;
                    s++;
                    if (tq > 3 || pq > 1 || s + (pq ? 128 : 64) > se_) return fail("bad DQT");
                    for (int i = 0; i < 64; i++) {
                        qt[tq][kZigzag[i]] = pq ? (uint16_t)((s[0] << 8) | s[1]) : s[0];
                        s += pq ? 2 : 1;
                    }
                    qt_present[tq] = true;
                }
            } else if (m == 0xc4) {                                // DHT
                while (s < se_) {
                    if (s + 17 > se_) return fail("bad DHT");
                    const int tc = s[0] >> 4, th = s[0] & 15;
                    s++;
                    if (th > 3 || tc > 1) return fail("bad DHT");
                    Huff& h = tc ? ac[th] : dc[th];
                    int n = 0;
                    h.bits[0] = 0;
                    for (int i = 1; i <= 16; i++) { h.bits[i] = s[i - 1]; n += s[i - 1]; }
                    s += 16;
                    if (n > 256 || s + n > se_) return fail("bad DHT");
                    memcpy(h.vals, s, n);
                    s += n;
                    h.present = true;
                    h.build();
                }
            } else if (m == 0xc0 || m == 0xc1 || m == 0xc2) {      // SOF0/1/2
                if (have_frame) return fail("more than one frame header");
                if (s + 6 > se_) return fail("truncated SOF");
                progressive = (m == 0xc2);
                if (s[0] != 8) return fail("only 8-bit JPEG is supported");
                height = (s[1] << 8) | s[2];
                width = (s[3] << 8) | s[4];
                ncomp = s[5];
                if ((ncomp != 1 && ncomp != 3) || !width || !height) return fail("unsupported JPEG component count");
                if (s + 6 + 3 * ncomp > se_) return fail("truncated SOF");
                for (int i = 0; i < ncomp; i++) {
                    comp[i].id = s[6 + 3 * i];
                    comp[i].h = s[7 + 3 * i] >> 4;
                    comp[i].v = s[7 + 3 * i] & 15;
                    comp[i].tq = s[8 + 3 * i];
                    if (!comp[i].h || !comp[i].v || comp[i].h > 4 || comp[i].v > 4 || comp[i].tq > 3) return fail("bad SOF");
                    hmax = comp[i].h > hmax ? comp[i].h : hmax;
                    vmax = comp[i].v > vmax ? comp[i].v : vmax;
                }
                if ((size_t)width * height > ((size_t)1 << 28)) return fail("JPEG larger than 2^28 pixels");
                mcus_x = (width + 8 * hmax - 1) / (8 * hmax);
                mcus_y = (height + 8 * vmax - 1) / (8 * vmax);
                for (int i = 0; i < ncomp; i++) {
                    Component& c = comp[i];
                    c.bw = mcus_x * c.h; c.bh = mcus_y * c.v;
                    c.sw = (width * c.h + hmax - 1) / hmax; c.sh = (height * c.v + vmax - 1) / vmax;
                    c.cw = (c.sw + 7) / 8; c.chh = (c.sh + 7) / 8;
                    c.coef.assign((size_t)c.bw * c.bh * 64, 0);
                }
                have_frame = true;
            } else if (m == 0xc3 || (m >= 0xc5 && m <= 0xcf && m != 0xc8 && m != 0xcc)) {
                return fail("unsupported JPEG process (lossless / hierarchical / arithmetic)");
            } else if (m == 0xdd) {                                // DRI
                if (s + 2 > se_) return fail("truncated DRI");
                restart_interval = (s[0] << 8) | s[1];
            } else if (m == 0xee && len >= 14 && !memcmp(s, "Adobe", 5)) {
                adobe_transform = s[11];
            } else if (m == 0xda) {                                // SOS
                if (!have_frame) return fail("SOS before SOF");
                if (s + 1 > se_) return fail("truncated SOS");
                const int ns = s[0];
                if (ns < 1 || ns > ncomp || s + 4 + 2 * ns > se_) return fail("bad SOS");
                int sc[3];
                for (int i = 0; i < ns; i++) {
                    int ci = -1;
                    for (int j = 0; j < ncomp; j++) if (comp[j].id == s[1 + 2 * i]) ci = j;
                    if (ci < 0) return fail("SOS names an unknown component");
                    comp[ci].td = s[2 + 2 * i] >> 4;
                    comp[ci].ta = s[2 + 2 * i] & 15;
                    if (comp[ci].td > 3 || comp[ci].ta > 3) return fail("bad SOS table id");
                    sc[i] = ci;
                }
                const int ss = s[1 + 2 * ns], se = s[2 + 2 * ns], ah = s[3 + 2 * ns] >> 4, al = s[3 + 2 * ns] & 15;
                if (progressive) {
                    // spectral selection / successive approximation of ITU T.81 G.1.1.1: 0 <= Ss <= Se <= 63, a DC scan has
                    // Se = 0, an AC scan names one component; Al <= 13 keeps 1 << al inside an int16 coefficient
                    if (ss > se || se > 63 || al > 13 || ah > 13) return fail("bad progressive scan parameters");
                    if (ss == 0 && se != 0) return fail("progressive DC scan with Se != 0");
                    if (ss > 0 && ns != 1) return fail("progressive AC scan over several components");
                }
                for (int i = 0; i < ns; i++) {
                    const Component& cc = comp[sc[i]];
                    if ((!progressive || ss == 0) && !dc[cc.td].present && !(progressive && ah != 0)) return fail("scan uses a missing DC table");
                    if ((!progressive || ss > 0) && !ac[cc.ta].present) return fail("scan uses a missing AC table");
                }
                const uint8_t* q = se_;
                if (!decode_scan(q, end, ns, sc, progressive ? ss : 0, progressive ? se : 63, progressive ? ah : 0, progressive ? al : 0))
                    return false;
                p = q;
                continue;
            }
            p = se_;
        }
        if (!have_frame) return fail("no frame header");
        if (ncomp == 3 && adobe_transform == 0) return fail("Adobe RGB JPEG without a colour transform is not supported");

        // dequantise + inverse DCT into per-component planes
        for (int i = 0; i < ncomp; i++) {
            Component& c = comp[i];
            if (!qt_present[c.tq]) return fail("missing quantisation table");
            const int stride = c.bw * 8;
            c.pix.assign((size_t)stride * c.bh * 8, 0);
            for (int by = 0; by < c.bh; by++)
                for (int bx = 0; bx < c.bw; bx++)
                    idct(&c.coef[((size_t)by * c.bw + bx) * 64], qt[c.tq], &c.pix[(size_t)by * 8 * stride + bx * 8], stride);
        }
        // upsample + colour convert
        out.width = (uint32_t)width; out.height = (uint32_t)height;
        out.rgba.assign((size_t)width * height * 4, 255);
        std::vector<uint8_t> up[3];
        for (int i = 0; i < ncomp; i++) {
            if (!upsample(comp[i], up[i])) return fail("unsupported chroma subsampling");
        }
        if (ncomp == 1) {
            for (size_t i = 0; i < (size_t)width * height; i++) { const uint8_t y = up[0][i]; uint8_t* o = &out.rgba[4 * i]; o[0] = o[1] = o[2] = y; }
            return true;
        }
        for (size_t i = 0; i < (size_t)width * height; i++) {
            const int y = up[0][i], cb = up[1][i] - 128, cr = up[2][i] - 128;
            // 16-bit fixed point: FIX(1.40200) = 91881, FIX(1.77200) = 116130, FIX(0.71414) = 46802, FIX(0.34414) = 22554
            const int r = y + ((91881 * cr + 32768) >> 16);
            const int g = y + ((-22554 * cb + 32768 - 46802 * cr) >> 16);
            const int b = y + ((116130 * cb + 32768) >> 16);
            uint8_t* o = &out.rgba[4 * i];
            o[0] = (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
            o[1] = (uint8_t)(g < 0 ? 0 : (g > 255 ? 255 : g));
            o[2] = (uint8_t)(b < 0 ? 0 : (b > 255 ? 255 : b));
        }
        return true;
    }

    // Component plane -> full-resolution plane [height][width] with triangle-filter upsampling.
    bool upsample(const Component& c, std::vector<uint8_t>& dst) const
    {
        dst.assign((size_t)width * height, 0);
        const int stride = c.bw * 8;
        const int hs = hmax / c.h, vs = vmax / c.v;
        if (hmax % c.h || vmax % c.v) return false;
        auto src = [&](int x, int y) -> int {
            x = x < 0 ? 0 : (x >= c.sw ? c.sw - 1 : x);
            y = y < 0 ? 0 : (y >= c.sh ? c.sh - 1 : y);
            return c.pix[(size_t)y * stride + x];
        };
        if (hs == 1 && vs == 1) {
            for (int y = 0; y < height; y++) memcpy(&dst[(size_t)y * width], &c.pix[(size_t)y * stride], width);
            return true;
        }
        if (hs == 2 && vs == 1) {   // h2v1 fancy
            for (int y = 0; y < height; y++)
                for (int x = 0; x < width; x++) {
                    const int i = x >> 1;
                    int v;
                    if (c.sw == 1) v = src(0, y);
                    else if (x == 0) v = src(0, y);
                    else if (x == 2 * c.sw - 1) v = src(c.sw - 1, y);
                    else if (x & 1) v = (3 * src(i, y) + src(i + 1, y) + 2) >> 2;
                    else v = (3 * src(i, y) + src(i - 1, y) + 1) >> 2;
                    dst[(size_t)y * width + x] = (uint8_t)v;
                }
            return true;
        }
        if (hs == 1 && vs == 2) {   // h1v2 fancy
            for (int y = 0; y < height; y++) {
                const int j = y >> 1, other = (y & 1) ? j + 1 : j - 1, bias = (y & 1) ? 2 : 1;
                for (int x = 0; x < width; x++) dst[(size_t)y * width + x] = (uint8_t)((3 * src(x, j) + src(x, other) + bias) >> 2);
            }
            return true;
        }
        if (hs == 2 && vs == 2) {   // h2v2 fancy: 3/4 nearer row + 1/4 further row, then 3/4-1/4 across columns
            for (int y = 0; y < height; y++) {
                const int j = y >> 1, other = (y & 1) ? j + 1 : j - 1;
                for (int x = 0; x < width; x++) {
                    const int i = x >> 1;
                    const int cur = 3 * src(i, j) + src(i, other);
                    int v;
                    if (c.sw == 1) v = (cur * 4 + 8) >> 4;
                    else if (x == 0) v = (cur * 4 + 8) >> 4;
                    else if (x == 2 * c.sw - 1) v = (cur * 4 + 7) >> 4;
                    else if (x & 1) v = (cur * 3 + (3 * src(i + 1, j) + src(i + 1, other)) + 7) >> 4;
                    else v = (cur * 3 + (3 * src(i - 1, j) + src(i - 1, other)) + 8) >> 4;
                    dst[(size_t)y * width + x] = (uint8_t)v;
                }
            }
            return true;
        }
        // other ratios: box replication
        for (int y = 0; y < height; y++)
            for (int x = 0; x < width; x++) dst[(size_t)y * width + x] = (uint8_t)src(x / hs, y / vs);
        return true;
    }
};

}  // namespace

bool decode_jpeg(const std::vector<uint8_t>& file, Image& out, std::string* why)
{
    if (file.size() < 4 || file[0] != 0xff || file[1] != 0xd8) { if (why) *why = "not a JPEG"; return false; }
    Decoder d;
    d.data = file.data();
    d.size = file.size();
    memset(d.qt, 0, sizeof(d.qt));
    try {
        if (!d.run(out)) { if (why) *why = d.err; return false; }
    } catch (const std::bad_alloc&) {      // a header may announce up to 65535 x 65535 samples
        if (why) *why = "out of memory while decoding the JPEG";
        return false;
    }
    return true;
}

}  // namespace rc
