// rz_image.hpp -- host-side image I/O for the C++ mirror of the crate API (header-only, no dependencies).
//
//   read_png   what Texture::from_png_file (texture.rs:26-45) gets from the `png` crate (0.16.8, default
//              transformations EXPAND | STRIP_16): 8-bit samples, palette and low-bit-depth images expanded.
//              Grey images are replicated to RGB so that the result is always RGB8 or RGBA8 -- the two
//              layouts Texture::read_texel knows (texture.rs:47-63).
//   write_png  the resolved 0xAARRGGBB framebuffer (rasterizer/buffers.rs:121-124) as an RGB8 PNG; the
//              reference only presents it in a minifb window (render.rs:116-127), so this is the headless
//              replacement.  Stored (uncompressed) deflate blocks: simple and exact.
//   write_ppm  the same image as a binary P6 file.
//
// Everything here is plain byte/integer work on the host; nothing touches the raster path.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace rz {
namespace image {

struct Image {
    std::vector<uint8_t> pixels; // row-major u8[height][width][channels], origin top-left
    uint32_t width = 0, height = 0, channels = 0;
};

inline uint32_t crc32(const uint8_t *p, size_t n, uint32_t crc = 0) {
    static uint32_t table[256];
    static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < 256; i++) {
            uint32_t c = i;
            for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        init = true;
    }
    crc = ~crc;
    for (size_t i = 0; i < n; i++) crc = table[(crc ^ p[i]) & 0xFF] ^ (crc >> 8);
    return ~crc;
}

inline uint32_t adler32(const uint8_t *p, size_t n) {
    uint32_t a = 1, b = 0;
    for (size_t i = 0; i < n; i++) {
        a = (a + p[i]) % 65521u;
        b = (b + a) % 65521u;
    }
    return (b << 16) | a;
}

// ---- inflate (RFC 1951) ------------------------------------------------------------------------------
namespace detail {

struct BitReader {
    const uint8_t *p;
    size_t n, pos = 0;
    uint32_t acc = 0;
    int cnt = 0;
    uint32_t bits(int k) {
        while (cnt < k) {
            if (pos >= n) throw std::runtime_error("png: truncated deflate stream");
            acc |= (uint32_t)p[pos++] << cnt;
            cnt += 8;
        }
        const uint32_t v = acc & ((k == 32) ? 0xFFFFFFFFu : ((1u << k) - 1u));
        acc >>= k;
        cnt -= k;
        return v;
    }
    void align() {
        acc = 0;
        cnt = 0;
    }
};

struct Huffman { // canonical code: counts per length + symbols sorted by (length, value)
    uint16_t count[16];
    std::vector<uint16_t> symbol;
    void build(const uint8_t *lengths, int n) {
        std::memset(count, 0, sizeof count);
        for (int i = 0; i < n; i++) count[lengths[i]]++;
        count[0] = 0;
        uint16_t offs[16];
        offs[1] = 0;
        for (int l = 1; l < 15; l++) offs[l + 1] = offs[l] + count[l];
        symbol.assign(n, 0);
        for (int i = 0; i < n; i++)
            if (lengths[i]) symbol[offs[lengths[i]]++] = (uint16_t)i;
    }
    int decode(BitReader &br) const {
        int code = 0, first = 0, index = 0;
        for (int l = 1; l <= 15; l++) {
            code |= (int)br.bits(1);
            const int c = count[l];
            if (code - c < first) return symbol[index + (code - first)];
            index += c;
            first += c;
            first <<= 1;
            code <<= 1;
        }
        throw std::runtime_error("png: bad Huffman code");
    }
};

inline std::vector<uint8_t> inflate(const uint8_t *src, size_t n, size_t expect) {
    static const uint16_t LBASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const uint8_t LEXT[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    static const uint16_t DBASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    static const uint8_t DEXT[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    static const uint8_t ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    // `expect` is a HARD cap on the output (the caller knows the image size from IHDR): a stream that inflates to more
    // is rejected instead of growing without bound (zip bomb)
    std::vector<uint8_t> out;
    out.reserve(expect);
    BitReader br{src, n};
    for (bool last = false; !last;) {
        last = br.bits(1) != 0;
        const uint32_t type = br.bits(2);
        if (type == 0) { // stored
            br.align();
            if (br.pos + 4 > n) throw std::runtime_error("png: truncated stored block");
            const uint32_t len = src[br.pos] | (src[br.pos + 1] << 8), nlen = src[br.pos + 2] | (src[br.pos + 3] << 8);
            br.pos += 4;
            if ((len ^ 0xFFFFu) != nlen || br.pos + len > n) throw std::runtime_error("png: bad stored block");
            if (out.size() + len > expect) throw std::runtime_error("png: image data longer than the header says");
            out.insert(out.end(), src + br.pos, src + br.pos + len);
            br.pos += len;
            continue;
        }
        if (type == 3) throw std::runtime_error("png: bad deflate block type");
        Huffman lit, dist;
        uint8_t lengths[320];
        if (type == 1) { // fixed codes
            int i = 0;
            for (; i < 144; i++) lengths[i] = 8;
            for (; i < 256; i++) lengths[i] = 9;
            for (; i < 280; i++) lengths[i] = 7;
            for (; i < 288; i++) lengths[i] = 8;
            lit.build(lengths, 288);
            for (i = 0; i < 30; i++) lengths[i] = 5;
            dist.build(lengths, 30);
        } else { // dynamic codes
            const int nlen = (int)br.bits(5) + 257, ndist = (int)br.bits(5) + 1, ncode = (int)br.bits(4) + 4;
            if (nlen > 286 || ndist > 30) throw std::runtime_error("png: bad dynamic header");
            uint8_t cl[19] = {0};
            for (int i = 0; i < ncode; i++) cl[ORDER[i]] = (uint8_t)br.bits(3);
            Huffman lencode;
            lencode.build(cl, 19);
            int i = 0;
            while (i < nlen + ndist) {
                const int sym = lencode.decode(br);
                if (sym < 16) {
                    lengths[i++] = (uint8_t)sym;
                } else {
                    int rep, val = 0;
                    if (sym == 16) {
                        if (i == 0) throw std::runtime_error("png: repeat with no previous length");
                        val = lengths[i - 1];
                        rep = 3 + (int)br.bits(2);
                    } else if (sym == 17) {
                        rep = 3 + (int)br.bits(3);
                    } else {
                        rep = 11 + (int)br.bits(7);
                    }
                    if (i + rep > nlen + ndist) throw std::runtime_error("png: too many code lengths");
                    while (rep--) lengths[i++] = (uint8_t)val;
                }
            }
            lit.build(lengths, nlen);
            dist.build(lengths + nlen, ndist);
        }
        for (;;) {
            const int sym = lit.decode(br);
            if (sym < 256) {
                if (out.size() >= expect) throw std::runtime_error("png: image data longer than the header says");
                out.push_back((uint8_t)sym);
            } else if (sym == 256) {
                break;
            } else {
                if (sym > 285) throw std::runtime_error("png: bad length symbol");
                const int len = LBASE[sym - 257] + (int)br.bits(LEXT[sym - 257]);
                const int ds = dist.decode(br);
                if (ds > 29) throw std::runtime_error("png: bad distance symbol");
                const size_t d = DBASE[ds] + br.bits(DEXT[ds]);
                if (d > out.size()) throw std::runtime_error("png: distance too far back");
                if (out.size() + (size_t)len > expect) throw std::runtime_error("png: image data longer than the header says");
                for (int k = 0; k < len; k++) out.push_back(out[out.size() - d]);
            }
        }
    }
    return out;
}

inline uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | (p[1] << 16) | (p[2] << 8) | p[3]; }
inline void put_be32(std::vector<uint8_t> &v, uint32_t x) {
    v.push_back((uint8_t)(x >> 24)); v.push_back((uint8_t)(x >> 16)); v.push_back((uint8_t)(x >> 8)); v.push_back((uint8_t)x);
}
inline int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = p > a ? p - a : a - p, pb = p > b ? p - b : b - p, pc = p > c ? p - c : c - p;
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

} // namespace detail

// Decode a PNG held in memory (non-interlaced; bit depths 1, 2, 4, 8, 16; all five colour types).
inline Image decode_png(const uint8_t *data, size_t size) {
    using namespace detail;
    static const uint8_t SIG[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (size < 8 || std::memcmp(data, SIG, 8) != 0) throw std::runtime_error("png: bad signature");
    uint32_t W = 0, H = 0;
    int depth = 0, ctype = -1;
    std::vector<uint8_t> idat, plte, trns;
    for (size_t pos = 8; pos + 12 <= size;) {
        const uint32_t len = be32(data + pos);
        const uint8_t *type = data + pos + 4, *body = data + pos + 8;
        if (pos + 12 + (size_t)len > size) throw std::runtime_error("png: truncated chunk");
        if (crc32(type, 4 + len) != be32(body + len)) throw std::runtime_error("png: chunk CRC mismatch");
        if (!std::memcmp(type, "IHDR", 4)) {
            if (len != 13) throw std::runtime_error("png: bad IHDR");
            W = be32(body); H = be32(body + 4); depth = body[8]; ctype = body[9];
            if (body[10] != 0 || body[11] != 0) throw std::runtime_error("png: unknown compression/filter method");
            if (body[12] != 0) throw std::runtime_error("png: interlaced images are not supported");
        } else if (!std::memcmp(type, "PLTE", 4)) {
            plte.assign(body, body + len);
        } else if (!std::memcmp(type, "tRNS", 4)) {
            trns.assign(body, body + len);
        } else if (!std::memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), body, body + len);
        } else if (!std::memcmp(type, "IEND", 4)) {
            break;
        }
        pos += 12 + (size_t)len;
    }
    if (W == 0 || H == 0 || ctype < 0) throw std::runtime_error("png: missing IHDR");
    // dimensions are untrusted: bound them (the framebuffers and textures of this library are <= 65535 on a side), so
    // that (stride + 1) * H below cannot wrap size_t and a header cannot ask for an absurd allocation
    if (W > 65535u || H > 65535u) throw std::runtime_error("png: image larger than 65535 pixels on a side");
    int samples;
    switch (ctype) {
    case 0: samples = 1; break;
    case 2: samples = 3; break;
    case 3: samples = 1; break;
    case 4: samples = 2; break;
    case 6: samples = 4; break;
    default: throw std::runtime_error("png: bad colour type");
    }
    const bool depth_ok = depth == 8 || (depth == 16 && ctype != 3) || ((depth == 1 || depth == 2 || depth == 4) && (ctype == 0 || ctype == 3));
    if (!depth_ok) throw std::runtime_error("png: bad bit depth for the colour type");
    if (ctype == 3 && plte.empty()) throw std::runtime_error("png: palette image without PLTE");
    const size_t bpp = (size_t)(samples * depth + 7) / 8;           // filter unit, bytes
    const size_t stride = ((size_t)W * samples * depth + 7) / 8;    // bytes per scanline
    if (idat.size() < 6) throw std::runtime_error("png: no image data");
    std::vector<uint8_t> raw = inflate(idat.data() + 2, idat.size() - 2, (stride + 1) * H); // skip the zlib header
    if (raw.size() < (stride + 1) * H) throw std::runtime_error("png: image data too short");
    // un-filter in place (filter types 0..4)
    std::vector<uint8_t> prev(stride, 0);
    for (uint32_t y = 0; y < H; y++) {
        uint8_t *line = &raw[(stride + 1) * y + 1];
        const int ft = line[-1];
        for (size_t i = 0; i < stride; i++) {
            const int a = i >= bpp ? line[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
            int add;
            switch (ft) {
            case 0: add = 0; break;
            case 1: add = a; break;
            case 2: add = b; break;
            case 3: add = (a + b) >> 1; break;
            case 4: add = paeth(a, b, c); break;
            default: throw std::runtime_error("png: bad filter type");
            }
            line[i] = (uint8_t)(line[i] + add);
        }
        std::memcpy(prev.data(), line, stride);
    }
    // expand to RGB8 / RGBA8
    const bool alpha = ctype == 4 || ctype == 6 || !trns.empty();
    Image img;
    img.width = W; img.height = H; img.channels = alpha ? 4 : 3;
    img.pixels.resize((size_t)W * H * img.channels);
    const uint32_t maxv = (1u << depth) - 1u;
    auto sample = [&](const uint8_t *line, size_t idx) -> uint32_t { // idx-th sample of the scanline, native depth
        if (depth == 8) return line[idx];
        if (depth == 16) return ((uint32_t)line[2 * idx] << 8) | line[2 * idx + 1];
        const size_t bit = idx * depth;
        return (line[bit >> 3] >> (8 - depth - (bit & 7))) & maxv;
    };
    auto to8 = [&](uint32_t v) -> uint8_t { // STRIP_16 keeps the high byte; low depths are scaled like EXPAND
        if (depth == 8) return (uint8_t)v;
        if (depth == 16) return (uint8_t)(v >> 8);
        return (uint8_t)(v * 255u / maxv);
    };
    for (uint32_t y = 0; y < H; y++) {
        const uint8_t *line = &raw[(stride + 1) * y + 1];
        uint8_t *dst = &img.pixels[(size_t)y * W * img.channels];
        for (uint32_t x = 0; x < W; x++, dst += img.channels) {
            uint8_t r, g, b, a = 255;
            if (ctype == 3) {
                const uint32_t i = sample(line, x);
                if (3 * i + 2 >= plte.size()) throw std::runtime_error("png: palette index out of range");
                r = plte[3 * i]; g = plte[3 * i + 1]; b = plte[3 * i + 2];
                if (i < trns.size()) a = trns[i];
            } else if (ctype == 0 || ctype == 4) {
                const uint32_t v = sample(line, (size_t)x * samples);
                r = g = b = to8(v);
                if (ctype == 4) a = to8(sample(line, (size_t)x * 2 + 1));
                else if (trns.size() >= 2 && v == (((uint32_t)trns[0] << 8) | trns[1])) a = 0;
            } else {
                const uint32_t vr = sample(line, (size_t)x * samples), vg = sample(line, (size_t)x * samples + 1),
                               vb = sample(line, (size_t)x * samples + 2);
                r = to8(vr); g = to8(vg); b = to8(vb);
                if (ctype == 6) a = to8(sample(line, (size_t)x * 4 + 3));
                else if (trns.size() >= 6 && vr == (((uint32_t)trns[0] << 8) | trns[1]) && vg == (((uint32_t)trns[2] << 8) | trns[3]) &&
                         vb == (((uint32_t)trns[4] << 8) | trns[5]))
                    a = 0;
            }
            dst[0] = r; dst[1] = g; dst[2] = b;
            if (alpha) dst[3] = a;
        }
    }
    return img;
}

inline std::vector<uint8_t> read_file(const std::string &path) {
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("cannot open " + path);
    std::vector<uint8_t> buf;
    uint8_t tmp[65536];
    size_t n;
    while ((n = std::fread(tmp, 1, sizeof tmp, f)) > 0) buf.insert(buf.end(), tmp, tmp + n);
    std::fclose(f);
    return buf;
}

inline Image read_png(const std::string &path) {
    const std::vector<uint8_t> buf = read_file(path);
    return decode_png(buf.data(), buf.size());
}

// Encode RGB8 rows (u8[height][width][3]) as a PNG with stored deflate blocks.
inline std::vector<uint8_t> encode_png_rgb(const uint8_t *rgb, uint32_t W, uint32_t H) {
    using namespace detail;
    std::vector<uint8_t> raw; // filter byte 0 + row
    raw.reserve(((size_t)W * 3 + 1) * H);
    for (uint32_t y = 0; y < H; y++) {
        raw.push_back(0);
        raw.insert(raw.end(), rgb + (size_t)y * W * 3, rgb + (size_t)(y + 1) * W * 3);
    }
    std::vector<uint8_t> z = {0x78, 0x01};
    for (size_t pos = 0; pos < raw.size() || pos == 0;) {
        const size_t n = raw.size() - pos < 65535 ? raw.size() - pos : 65535;
        z.push_back(pos + n >= raw.size() ? 1 : 0);
        z.push_back((uint8_t)n); z.push_back((uint8_t)(n >> 8));
        z.push_back((uint8_t)~n); z.push_back((uint8_t)(~n >> 8));
        z.insert(z.end(), raw.begin() + pos, raw.begin() + pos + n);
        pos += n;
        if (n == 0) break;
    }
    put_be32(z, adler32(raw.data(), raw.size()));
    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    auto chunk = [&](const char *type, const std::vector<uint8_t> &body) {
        put_be32(out, (uint32_t)body.size());
        const size_t start = out.size();
        out.insert(out.end(), type, type + 4);
        out.insert(out.end(), body.begin(), body.end());
        put_be32(out, crc32(&out[start], out.size() - start));
    };
    std::vector<uint8_t> ihdr;
    put_be32(ihdr, W); put_be32(ihdr, H);
    ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    chunk("IHDR", ihdr);
    chunk("IDAT", z);
    chunk("IEND", {});
    return out;
}

// 0xAARRGGBB framebuffer (rasterizer/buffers.rs:121-124) -> RGB8 rows
inline std::vector<uint8_t> framebuffer_to_rgb(const uint32_t *fb, size_t W, size_t H) {
    std::vector<uint8_t> rgb(W * H * 3);
    for (size_t i = 0; i < W * H; i++) {
        rgb[3 * i] = (uint8_t)(fb[i] >> 16); rgb[3 * i + 1] = (uint8_t)(fb[i] >> 8); rgb[3 * i + 2] = (uint8_t)fb[i];
    }
    return rgb;
}

inline void write_file(const std::string &path, const uint8_t *p, size_t n) {
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot create " + path);
    const size_t w = std::fwrite(p, 1, n, f);
    std::fclose(f);
    if (w != n) throw std::runtime_error("short write to " + path);
}

inline void write_png(const std::string &path, const uint32_t *fb, size_t W, size_t H) {
    const std::vector<uint8_t> rgb = framebuffer_to_rgb(fb, W, H);
    const std::vector<uint8_t> png = encode_png_rgb(rgb.data(), (uint32_t)W, (uint32_t)H);
    write_file(path, png.data(), png.size());
}

inline void write_ppm(const std::string &path, const uint32_t *fb, size_t W, size_t H) {
    const std::vector<uint8_t> rgb = framebuffer_to_rgb(fb, W, H);
    std::string head = "P6\n" + std::to_string(W) + " " + std::to_string(H) + "\n255\n";
    std::vector<uint8_t> out(head.begin(), head.end());
    out.insert(out.end(), rgb.begin(), rgb.end());
    write_file(path, out.data(), out.size());
}

} // namespace image
} // namespace rz
