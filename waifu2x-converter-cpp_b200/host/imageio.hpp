// imageio.hpp -- cv::imread(path, IMREAD_COLOR) / cv::imwrite(path, image) for the drop-in CLI (src/main.cpp:74,190)
// without OpenCV: 8-bit PNG (zlib) and binary PPM/PGM.  Pixels come back as 3-channel B,G,R bytes like imread does.
#ifndef W2X_IMAGEIO_HPP_
#define W2X_IMAGEIO_HPP_

#include <zlib.h>

#include <cctype>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <algorithm>
#include <string>
#include <thread>
#include <vector>

namespace w2xio {

struct Image8 {
    int width = 0, height = 0;
    std::vector<uint8_t> bgr;   // [h][w][3]
    bool empty() const { return bgr.empty(); }
};

namespace detail {
inline uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
inline void put32(std::vector<uint8_t> &v, uint32_t x) { v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x); }
inline int paeth(int a, int b, int c) {
    int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

inline bool read_png(const std::vector<uint8_t> &f, Image8 &out, std::string &err) {
    static const uint8_t sig[8] = {137, 80, 78, 71, 13, 10, 26, 10};
    if (f.size() < 8 || std::memcmp(f.data(), sig, 8)) { err = "not a PNG file"; return false; }
    size_t pos = 8;
    uint32_t w = 0, h = 0;
    int depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, plte;
    while (pos + 12 <= f.size()) {
        uint32_t len = be32(&f[pos]);
        const uint8_t *type = &f[pos + 4], *data = &f[pos + 8];
        if (pos + 12 + len > f.size()) { err = "truncated PNG chunk"; return false; }
        if (!std::memcmp(type, "IHDR", 4)) {
            if (len != 13) { err = "bad PNG IHDR length"; return false; }
            w = be32(data); h = be32(data + 4); depth = data[8]; ctype = data[9]; interlace = data[12];
        } else if (!std::memcmp(type, "PLTE", 4)) plte.assign(data, data + len);
        else if (!std::memcmp(type, "IDAT", 4)) idat.insert(idat.end(), data, data + len);
        else if (!std::memcmp(type, "IEND", 4)) break;
        pos += 12 + len;
    }
    if (!w || !h) { err = "PNG without IHDR"; return false; }
    if (w > 65535u || h > 65535u || (uint64_t)w * h > ((uint64_t)1 << 28)) { err = "PNG dimensions out of range"; return false; }   // (a crafted header must not drive the allocations below)
    if (interlace) { err = "interlaced PNG is not supported"; return false; }
    if (depth != 8 && depth != 16) { err = "only 8/16-bit PNG is supported"; return false; }
    int ch = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    if (!ch || (ctype == 3 && depth != 8)) { err = "unsupported PNG colour type"; return false; }
    const size_t bpp = (size_t)ch * depth / 8, stride = bpp * w;
    std::vector<uint8_t> raw((stride + 1) * h);
    uLongf rawlen = (uLongf)raw.size();
    if (uncompress(raw.data(), &rawlen, idat.data(), (uLong)idat.size()) != Z_OK || rawlen != raw.size()) { err = "PNG inflate failed"; return false; }
    std::vector<uint8_t> img(stride * h);
    for (uint32_t y = 0; y < h; y++) {
        const uint8_t *src = &raw[(stride + 1) * y];
        uint8_t *cur = &img[stride * y];
        const uint8_t *prev = y ? &img[stride * (y - 1)] : nullptr;
        int ft = src[0];
        for (size_t i = 0; i < stride; i++) {
            int a = i >= bpp ? cur[i - bpp] : 0, b = prev ? prev[i] : 0, c = (prev && i >= bpp) ? prev[i - bpp] : 0;
            int x = src[i + 1];
            switch (ft) {
                case 0: break;
                case 1: x += a; break;
                case 2: x += b; break;
                case 3: x += (a + b) / 2; break;
                case 4: x += paeth(a, b, c); break;
                default: err = "bad PNG filter"; return false;
            }
            cur[i] = (uint8_t)x;
        }
    }
    out.width = (int)w; out.height = (int)h;
    out.bgr.resize((size_t)w * h * 3);
    const size_t step = depth / 8;   // 16-bit: keep the high byte (imread IMREAD_COLOR converts to 8 bit)
    for (size_t i = 0; i < (size_t)w * h; i++) {
        const uint8_t *p = &img[i * bpp];
        uint8_t r, g, b;
        if (ctype == 0 || ctype == 4) r = g = b = p[0];
        else if (ctype == 3) {
            size_t k = (size_t)p[0] * 3;
            if (k + 2 < plte.size()) { r = plte[k]; g = plte[k + 1]; b = plte[k + 2]; }
            else { r = g = b = 0; }
        } else { r = p[0]; g = p[step]; b = p[2 * step]; }
        out.bgr[i * 3] = b; out.bgr[i * 3 + 1] = g; out.bgr[i * 3 + 2] = r;
    }
    return true;
}

inline bool read_pnm(const std::vector<uint8_t> &f, Image8 &out, std::string &err) {
    size_t pos = 2;
    auto next_int = [&](int &v) {
        while (pos < f.size()) {
            if (f[pos] == '#') { while (pos < f.size() && f[pos] != '\n') pos++; }
            else if (std::isspace(f[pos])) pos++;
            else break;
        }
        v = 0;
        bool any = false;
        while (pos < f.size() && std::isdigit(f[pos])) { v = v * 10 + (f[pos++] - '0'); any = true; }
        return any;
    };
    int w, h, mx;
    if (!next_int(w) || !next_int(h) || !next_int(mx) || mx != 255) { err = "unsupported PNM header"; return false; }
    pos++;
    const int ch = f[1] == '6' ? 3 : 1;
    if (pos + (size_t)w * h * ch > f.size()) { err = "truncated PNM"; return false; }
    out.width = w; out.height = h;
    out.bgr.resize((size_t)w * h * 3);
    for (size_t i = 0; i < (size_t)w * h; i++) {
        const uint8_t *p = &f[pos + i * ch];
        uint8_t r = p[0], g = ch == 3 ? p[1] : p[0], b = ch == 3 ? p[2] : p[0];
        out.bgr[i * 3] = b; out.bgr[i * 3 + 1] = g; out.bgr[i * 3 + 2] = r;
    }
    return true;
}
}  // namespace detail

// cv::imread(path, cv::IMREAD_COLOR): empty image on failure (the reference does not check, src/main.cpp:74)
inline Image8 imread(const std::string &path, std::string *err_out = nullptr) {
    Image8 img;
    std::string err;
    std::ifstream in(path, std::ios::binary);
    if (!in) { err = "cannot open " + path; }
    else {
        std::vector<uint8_t> f((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
        if (f.size() > 2 && f[0] == 'P' && (f[1] == '6' || f[1] == '5')) detail::read_pnm(f, img, err);
        else detail::read_png(f, img, err);
        if (!err.empty()) img = Image8();
    }
    if (err_out) *err_out = err;
    return img;
}

// cv::imwrite(path, image): format by extension (.png, .ppm)
inline bool imwrite(const std::string &path, const uint8_t *bgr, int w, int h) {
    std::string ext = path.size() >= 4 ? path.substr(path.size() - 4) : "";
    for (auto &c : ext) c = (char)std::tolower(c);
    if (ext != ".ppm" && ext != ".png") return false;          // cv::imwrite picks the encoder by extension; PNG and binary PPM are what this front end has
    std::ofstream out(path, std::ios::binary);
    if (!out) return false;
    if (ext == ".ppm") {
        out << "P6\n" << w << " " << h << "\n255\n";
        std::vector<uint8_t> row((size_t)w * 3);
        for (int y = 0; y < h; y++) {
            for (int x = 0; x < w; x++) { const uint8_t *p = bgr + ((size_t)y * w + x) * 3; row[x * 3] = p[2]; row[x * 3 + 1] = p[1]; row[x * 3 + 2] = p[0]; }
            out.write(reinterpret_cast<const char *>(row.data()), (std::streamsize)row.size());
        }
        return (bool)out;
    }
    // PNG, 8-bit RGB.  Encoder settings are cv::imwrite's defaults (filter SUB, Z_BEST_SPEED, Z_RLE: grfmt_png.cpp); the rows are
    // deflated in independent bands on several threads -- every band but the last ends on a full flush (byte-aligned, no final
    // block), so the concatenation is one valid deflate stream; the zlib trailer is the combined Adler-32 of the bands.
    const size_t stride = (size_t)w * 3 + 1;
    int nband = (int)std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()), 32);
    nband = std::max(1, std::min(nband, h / 16));
    if ((size_t)w * h < (1u << 16)) nband = 1;
    std::vector<std::vector<uint8_t>> parts((size_t)nband);
    std::vector<uLong> adler((size_t)nband), lens((size_t)nband);
    std::vector<int> ok((size_t)nband, 0);
    auto work = [&](int bi) {
        const int y0 = (int)((long)h * bi / nband), y1 = (int)((long)h * (bi + 1) / nband);
        std::vector<uint8_t> raw(stride * (size_t)(y1 - y0));
        for (int y = y0; y < y1; y++) {
            uint8_t *r = &raw[stride * (size_t)(y - y0)];
            const uint8_t *p = bgr + (size_t)y * w * 3;
            r[0] = 1;                                                            // filter type 1 (SUB): byte - byte of the pixel to the left
            uint8_t pr = 0, pg = 0, pb = 0;
            for (int x = 0; x < w; x++, p += 3) {
                r[1 + x * 3] = (uint8_t)(p[2] - pr); r[2 + x * 3] = (uint8_t)(p[1] - pg); r[3 + x * 3] = (uint8_t)(p[0] - pb);
                pr = p[2]; pg = p[1]; pb = p[0];
            }
        }
        adler[(size_t)bi] = adler32(adler32(0L, Z_NULL, 0), raw.data(), (uInt)raw.size());
        lens[(size_t)bi] = (uLong)raw.size();
        z_stream zs;
        std::memset(&zs, 0, sizeof zs);
        if (deflateInit2(&zs, Z_BEST_SPEED, Z_DEFLATED, -15, 8, Z_RLE) != Z_OK) return;
        std::vector<uint8_t> &o = parts[(size_t)bi];
        o.resize(deflateBound(&zs, (uLong)raw.size()) + 16);
        zs.next_in = raw.data(); zs.avail_in = (uInt)raw.size();
        zs.next_out = o.data(); zs.avail_out = (uInt)o.size();
        const int rc = deflate(&zs, bi == nband - 1 ? Z_FINISH : Z_FULL_FLUSH);
        if ((bi == nband - 1 && rc == Z_STREAM_END) || (bi != nband - 1 && rc == Z_OK && zs.avail_in == 0)) ok[(size_t)bi] = 1;
        o.resize(o.size() - zs.avail_out);
        deflateEnd(&zs);
    };
    if (nband == 1) work(0);
    else {
        std::vector<std::thread> ts;
        for (int bi = 0; bi < nband; bi++) ts.emplace_back(work, bi);
        for (auto &t : ts) t.join();
    }
    // one IDAT chunk per band (their concatenation is the zlib stream): header in front of the first, Adler-32 behind the last
    uLong ad = adler[0];
    for (int bi = 0; bi < nband; bi++) {
        if (!ok[(size_t)bi]) return false;
        if (bi > 0) ad = adler32_combine(ad, adler[(size_t)bi], (z_off_t)lens[(size_t)bi]);
    }
    auto chunk = [&](const char *type, const uint8_t *head, size_t nhead, const std::vector<uint8_t> &data, const uint8_t *tail, size_t ntail) {
        uint8_t len[4], crcb[4];
        const uint32_t n = (uint32_t)(nhead + data.size() + ntail);
        len[0] = n >> 24; len[1] = n >> 16; len[2] = n >> 8; len[3] = n;
        uLong crc = crc32(0L, reinterpret_cast<const Bytef *>(type), 4);
        if (nhead) crc = crc32(crc, head, (uInt)nhead);
        if (!data.empty()) crc = crc32(crc, data.data(), (uInt)data.size());
        if (ntail) crc = crc32(crc, tail, (uInt)ntail);
        crcb[0] = crc >> 24; crcb[1] = crc >> 16; crcb[2] = crc >> 8; crcb[3] = crc;
        out.write(reinterpret_cast<const char *>(len), 4);
        out.write(type, 4);
        if (nhead) out.write(reinterpret_cast<const char *>(head), (std::streamsize)nhead);
        if (!data.empty()) out.write(reinterpret_cast<const char *>(data.data()), (std::streamsize)data.size());
        if (ntail) out.write(reinterpret_cast<const char *>(tail), (std::streamsize)ntail);
        out.write(reinterpret_cast<const char *>(crcb), 4);
    };
    static const uint8_t sig[8] = {137, 80, 78, 71, 13, 10, 26, 10};
    out.write(reinterpret_cast<const char *>(sig), 8);
    std::vector<uint8_t> ihdr;
    detail::put32(ihdr, (uint32_t)w); detail::put32(ihdr, (uint32_t)h);
    ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    chunk("IHDR", nullptr, 0, ihdr, nullptr, 0);
    static const uint8_t zhdr[2] = {0x78, 0x01};                                 // zlib header: deflate, 32 KB window, fastest
    const uint8_t ztail[4] = {(uint8_t)(ad >> 24), (uint8_t)(ad >> 16), (uint8_t)(ad >> 8), (uint8_t)ad};
    for (int bi = 0; bi < nband; bi++)
        chunk("IDAT", bi == 0 ? zhdr : nullptr, bi == 0 ? 2 : 0, parts[(size_t)bi], bi == nband - 1 ? ztail : nullptr, bi == nband - 1 ? 4 : 0);
    chunk("IEND", nullptr, 0, {}, nullptr, 0);
    return (bool)out;
}

}  // namespace w2xio
#endif
