// mt_assets.cpp -- asset pipeline of the cloud pass without stb / PIL: decodes the reference's texture
// files into the RGBA8 arrays mtUploadTexture* takes.
//   Sky::CreateCloudResources (Sky.cpp:25-58) -> ImageLoadingUtility::create3DTextureFromMany2DTextures
//   (ImageLoadingUtility.cpp:75-139: slice z = "<folder><base>(z+1)<ext>", stbi_load(..., STBI_rgb_alpha), memcpy into
//   volume[z][y][x][rgba]) and loadImageFromFile for the 2D textures.
// Formats covered = what the reference ships: TGA true-colour (uncompressed or RLE, 24/32 bpp, either origin) and PNG
// (non-interlaced, 8 or 16 bit, grey / grey+alpha / RGB / RGBA; 16-bit samples keep their high byte like stb_image).
// The inflate below is a plain RFC 1951 decoder (stored, fixed and dynamic Huffman blocks).  A decoded volume can be
// cached as a ".mtvol" file (16-byte header + raw RGBA8).  Host-only code: no CUDA, no context.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/meteoros_b200.h"

namespace {

// ---- file helpers --------------------------------------------------------------------------------------------------
bool read_file(const char* path, std::vector<uint8_t>& out)
{
    FILE* f = std::fopen(path, "rb");
    if (!f) return false;
    std::fseek(f, 0, SEEK_END);
    long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    if (n < 0) { std::fclose(f); return false; }
    out.resize((size_t)n);
    size_t got = n ? std::fread(out.data(), 1, (size_t)n, f) : 0;
    std::fclose(f);
    return got == (size_t)n;
}

// ---- TGA -----------------------------------------------------------------------------------------------------------
// A malformed or hostile file must be rejected, never crash the host: dimensions are bounded, every allocation is bounded
// by what the bytes actually present could decode to, and the C entry points catch whatever is left (bad_alloc).
constexpr int MT_MAX_IMAGE_DIM = 16384;

bool decode_tga(const uint8_t* d, size_t n, std::vector<uint8_t>& rgba, int& w, int& h)
{
    if (n < 18) return false;
    const int idlen = d[0], cmaptype = d[1], type = d[2];
    w = d[12] | (d[13] << 8);
    h = d[14] | (d[15] << 8);
    const int bpp = d[16], desc = d[17];
    if (cmaptype != 0 || (type != 2 && type != 10) || (bpp != 24 && bpp != 32) || w <= 0 || h <= 0) return false;
    if (w > MT_MAX_IMAGE_DIM || h > MT_MAX_IMAGE_DIM) return false;
    const int bytes = bpp / 8;
    size_t pos = 18 + (size_t)idlen;
    const size_t npix = (size_t)w * h;
    if (pos > n) return false;
    // raw: every pixel is in the file; RLE: a packet of 1 + bytes input bytes yields at most 128 pixels
    const size_t avail = n - pos;
    if (type == 2 ? npix * (size_t)bytes > avail : npix > (avail / (size_t)(1 + bytes) + 1) * 128) return false;
    rgba.assign(npix * 4, 255);
    size_t i = 0;
    auto put = [&](size_t k, const uint8_t* px) {  // file order is BGR(A)
        const size_t row = k / w, col = k % w;
        const size_t y = (desc & 0x20) ? row : (size_t)h - 1 - row;  // bit 5 clear: bottom-left origin -> flip to top-down
        const size_t x = (desc & 0x10) ? (size_t)w - 1 - col : col;
        uint8_t* o = &rgba[(y * w + x) * 4];
        o[0] = px[2]; o[1] = px[1]; o[2] = px[0];
        o[3] = bytes == 4 ? px[3] : 255;
    };
    if (type == 2) {
        if (pos + npix * bytes > n) return false;
        for (; i < npix; ++i) put(i, d + pos + i * bytes);
        return true;
    }
    while (i < npix) {  // RLE packets
        if (pos >= n) return false;
        const int hdr = d[pos++];
        const size_t cnt = (size_t)(hdr & 0x7f) + 1;
        if (i + cnt > npix) return false;
        if (hdr & 0x80) {
            if (pos + bytes > n) return false;
            for (size_t k = 0; k < cnt; ++k) put(i + k, d + pos);
            pos += bytes;
        } else {
            if (pos + cnt * bytes > n) return false;
            for (size_t k = 0; k < cnt; ++k) put(i + k, d + pos + k * bytes);
            pos += cnt * bytes;
        }
        i += cnt;
    }
    return true;
}

// ---- inflate (RFC 1951) ----------------------------------------------------------------------------------------------
struct BitReader {
    const uint8_t* p;
    size_t n, pos = 0;
    uint32_t buf = 0;
    int cnt = 0;
    bool ok = true;
    uint32_t bits(int k)
    {
        while (cnt < k) {
            if (pos >= n) { ok = false; return 0; }
            buf |= (uint32_t)p[pos++] << cnt;
            cnt += 8;
        }
        uint32_t v = buf & ((k == 32) ? 0xffffffffu : ((1u << k) - 1u));
        buf >>= k;
        cnt -= k;
        return v;
    }
};
struct Huffman {
    uint16_t count[16] = {};
    uint16_t symbol[320] = {};
    void build(const uint8_t* len, int n)
    {
        std::memset(count, 0, sizeof(count));
        for (int i = 0; i < n; ++i) count[len[i]]++;
        count[0] = 0;
        uint16_t offs[16];
        offs[1] = 0;
        for (int i = 1; i < 15; ++i) offs[i + 1] = offs[i] + count[i];
        for (int i = 0; i < n; ++i)
            if (len[i]) symbol[offs[len[i]]++] = (uint16_t)i;
    }
    int decode(BitReader& br) const
    {
        int code = 0, first = 0, index = 0;
        for (int l = 1; l <= 15; ++l) {
            code |= (int)br.bits(1);
            if (!br.ok) return -1;
            int c = count[l];
            if (code - c < first) return symbol[index + (code - first)];
            index += c;
            first += c;
            first <<= 1;
            code <<= 1;
        }
        return -1;
    }
};
bool inflate_raw(const uint8_t* src, size_t n, std::vector<uint8_t>& out, size_t max_out)
{
    static const uint16_t lbase[29] = { 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258 };
    static const uint16_t lext[29] = { 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0 };
    static const uint16_t dbase[30] = { 1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577 };
    static const uint16_t dext[30] = { 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13 };
    BitReader br{ src, n };
    int last;
    do {
        last = (int)br.bits(1);
        const int type = (int)br.bits(2);
        if (!br.ok) return false;
        if (type == 0) {
            br.buf = 0; br.cnt = 0;  // skip to the byte boundary
            if (br.pos + 4 > n) return false;
            const unsigned len = src[br.pos] | (src[br.pos + 1] << 8);
            br.pos += 4;
            if (br.pos + len > n || out.size() + len > max_out) return false;
            out.insert(out.end(), src + br.pos, src + br.pos + len);
            br.pos += len;
            continue;
        }
        if (type == 3) return false;
        Huffman lit, dist;
        uint8_t lengths[320];
        if (type == 1) {
            int i = 0;
            for (; i < 144; ++i) lengths[i] = 8;
            for (; i < 256; ++i) lengths[i] = 9;
            for (; i < 280; ++i) lengths[i] = 7;
            for (; i < 288; ++i) lengths[i] = 8;
            lit.build(lengths, 288);
            for (i = 0; i < 30; ++i) lengths[i] = 5;
            dist.build(lengths, 30);
        } else {
            static const uint8_t order[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };
            const int nlen = (int)br.bits(5) + 257, ndist = (int)br.bits(5) + 1, ncode = (int)br.bits(4) + 4;
            if (!br.ok || nlen > 286 || ndist > 30) return false;
            uint8_t cl[19] = {};
            for (int i = 0; i < ncode; ++i) cl[order[i]] = (uint8_t)br.bits(3);
            Huffman clh;
            clh.build(cl, 19);
            int idx = 0;
            while (idx < nlen + ndist) {
                int sym = clh.decode(br);
                if (sym < 0) return false;
                if (sym < 16) lengths[idx++] = (uint8_t)sym;
                else {
                    int rep, val = 0;
                    if (sym == 16) { if (idx == 0) return false; val = lengths[idx - 1]; rep = 3 + (int)br.bits(2); }
                    else if (sym == 17) rep = 3 + (int)br.bits(3);
                    else rep = 11 + (int)br.bits(7);
                    if (idx + rep > nlen + ndist) return false;
                    while (rep--) lengths[idx++] = (uint8_t)val;
                }
            }
            lit.build(lengths, nlen);
            dist.build(lengths + nlen, ndist);
        }
        for (;;) {
            int sym = lit.decode(br);
            if (sym < 0 || !br.ok) return false;
            if (sym < 256) {
                if (out.size() >= max_out) return false;
                out.push_back((uint8_t)sym);
            } else if (sym == 256) break;
            else {
                sym -= 257;
                if (sym >= 29) return false;
                const size_t len = lbase[sym] + br.bits(lext[sym]);
                const int ds = dist.decode(br);
                if (ds < 0 || ds >= 30) return false;
                const size_t d = dbase[ds] + br.bits(dext[ds]);
                if (d > out.size() || out.size() + len > max_out) return false;
                const size_t start = out.size() - d;
                for (size_t k = 0; k < len; ++k) out.push_back(out[start + k]);
            }
        }
    } while (!last);
    return br.ok;
}

// ---- PNG -----------------------------------------------------------------------------------------------------------
inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
bool decode_png(const uint8_t* d, size_t n, std::vector<uint8_t>& rgba, int& w, int& h)
{
    static const uint8_t sig[8] = { 0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n' };
    if (n < 8 || std::memcmp(d, sig, 8)) return false;
    size_t pos = 8;
    int depth = 0, ctype = 0;
    std::vector<uint8_t> z;
    bool have_hdr = false;
    while (pos + 12 <= n) {
        const uint32_t len = be32(d + pos);
        const uint8_t* type = d + pos + 4;
        if (pos + 12 + (size_t)len > n) return false;
        const uint8_t* body = d + pos + 8;
        if (!std::memcmp(type, "IHDR", 4)) {
            if (len < 13) return false;
            w = (int)be32(body); h = (int)be32(body + 4);
            depth = body[8]; ctype = body[9];
            if (body[10] || body[11] || body[12]) return false;  // compression / filter method 0, no interlace
            have_hdr = true;
        } else if (!std::memcmp(type, "IDAT", 4)) z.insert(z.end(), body, body + len);
        else if (!std::memcmp(type, "IEND", 4)) break;
        pos += 12 + (size_t)len;
    }
    if (!have_hdr || w <= 0 || h <= 0 || (depth != 8 && depth != 16)) return false;
    if (w > MT_MAX_IMAGE_DIM || h > MT_MAX_IMAGE_DIM) return false;
    int chans;
    switch (ctype) {
        case 0: chans = 1; break;
        case 2: chans = 3; break;
        case 4: chans = 2; break;
        case 6: chans = 4; break;
        default: return false;  // palette images are not used by the reference
    }
    if (z.size() < 6) return false;
    const size_t bpp = (size_t)chans * (depth / 8), stride = (size_t)w * bpp;
    const size_t need = (stride + 1) * (size_t)h;
    if (need / 1032 > z.size()) return false;  // deflate cannot expand by more than 1032:1: the pixels are not in this file
    std::vector<uint8_t> raw;
    raw.reserve(need);
    if (!inflate_raw(z.data() + 2, z.size() - 2, raw, need)) return false;  // skip the 2-byte zlib header; adler32 not checked
    if (raw.size() < need) return false;
    std::vector<uint8_t> prev(stride, 0), cur(stride);
    rgba.assign((size_t)w * h * 4, 255);
    for (int y = 0; y < h; ++y) {
        const uint8_t* line = &raw[(stride + 1) * (size_t)y];
        const int ft = line[0];
        for (size_t i = 0; i < stride; ++i) {
            const int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
            int v = line[1 + i];
            switch (ft) {
                case 0: break;
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) >> 1; break;
                case 4: {
                    const int p = a + b - c, pa = p > a ? p - a : a - p, pb = p > b ? p - b : b - p, pc = p > c ? p - c : c - p;
                    v += (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
                    break;
                }
                default: return false;
            }
            cur[i] = (uint8_t)v;
        }
        for (int x = 0; x < w; ++x) {
            uint8_t s[4];
            for (int c = 0; c < chans; ++c) s[c] = cur[(size_t)x * bpp + (size_t)c * (depth / 8)];  // 16 bit: the high byte (stb: v >> 8)
            uint8_t* o = &rgba[((size_t)y * w + x) * 4];
            if (chans == 1) { o[0] = o[1] = o[2] = s[0]; }
            else if (chans == 2) { o[0] = o[1] = o[2] = s[0]; o[3] = s[1]; }
            else { o[0] = s[0]; o[1] = s[1]; o[2] = s[2]; if (chans == 4) o[3] = s[3]; }
        }
        prev.swap(cur);
    }
    return true;
}

bool has_ext(const char* path, const char* ext)
{
    const size_t n = std::strlen(path), m = std::strlen(ext);
    if (n < m) return false;
    for (size_t i = 0; i < m; ++i) {
        char a = path[n - m + i], b = ext[i];
        if (a >= 'A' && a <= 'Z') a = (char)(a - 'A' + 'a');
        if (a != b) return false;
    }
    return true;
}

bool decode_any(const char* path, std::vector<uint8_t>& rgba, int& w, int& h)
{
    std::vector<uint8_t> file;
    if (!read_file(path, file)) return false;
    if (has_ext(path, ".png")) return decode_png(file.data(), file.size(), rgba, w, h);
    if (has_ext(path, ".tga")) return decode_tga(file.data(), file.size(), rgba, w, h);
    return decode_png(file.data(), file.size(), rgba, w, h) || decode_tga(file.data(), file.size(), rgba, w, h);
}

// No exception may cross the C boundary: allocation failure becomes MT_ERR_OOM, anything else MT_ERR_INVALID.
template <class F>
MtStatus guarded(F&& body) noexcept
{
    try {
        return body();
    } catch (const std::bad_alloc&) {
        return MT_ERR_OOM;
    } catch (...) {
        return MT_ERR_INVALID;
    }
}

}  // namespace

extern "C" {

MtStatus mtxDecodeImage(const uint8_t* file_bytes, size_t n, int is_png, uint8_t* rgba8_out, size_t out_bytes, uint32_t* w, uint32_t* h)
{
    return guarded([&]() -> MtStatus {
        if (!file_bytes || !w || !h) return MT_ERR_INVALID;
        std::vector<uint8_t> px;
        int iw = 0, ih = 0;
        const bool ok = is_png ? decode_png(file_bytes, n, px, iw, ih) : decode_tga(file_bytes, n, px, iw, ih);
        if (!ok) return MT_ERR_INVALID;
        *w = (uint32_t)iw;
        *h = (uint32_t)ih;
        if (rgba8_out) {
            if (out_bytes < px.size()) return MT_ERR_INVALID;
            std::memcpy(rgba8_out, px.data(), px.size());
        }
        return MT_OK;
    });
}

MtStatus mtxLoadImageFile(const char* path, uint8_t* rgba8_out, size_t out_bytes, uint32_t* w, uint32_t* h)
{
    return guarded([&]() -> MtStatus {
        if (!path || !w || !h) return MT_ERR_INVALID;
        std::vector<uint8_t> px;
        int iw = 0, ih = 0;
        if (!decode_any(path, px, iw, ih)) return MT_ERR_INVALID;
        *w = (uint32_t)iw;
        *h = (uint32_t)ih;
        if (rgba8_out) {
            if (out_bytes < px.size()) return MT_ERR_INVALID;
            std::memcpy(rgba8_out, px.data(), px.size());
        }
        return MT_OK;
    });
}

MtStatus mtxLoadVolumeFromSlices(const char* folder, const char* base_name, const char* extension, uint32_t w, uint32_t h, uint32_t d,
                                 uint8_t* rgba8_out, size_t out_bytes)
{
    return guarded([&]() -> MtStatus {
        if (!folder || !base_name || !extension || !rgba8_out) return MT_ERR_INVALID;
        const size_t slice = (size_t)w * h * 4;
        if (out_bytes < slice * d) return MT_ERR_INVALID;
        for (uint32_t z = 0; z < d; ++z) {  // ImageLoadingUtility.cpp:87-98
            const std::string path = std::string(folder) + base_name + "(" + std::to_string(z + 1) + ")" + extension;
            std::vector<uint8_t> px;
            int iw = 0, ih = 0;
            if (!decode_any(path.c_str(), px, iw, ih) || (uint32_t)iw != w || (uint32_t)ih != h) return MT_ERR_INVALID;
            std::memcpy(rgba8_out + slice * z, px.data(), slice);
        }
        return MT_OK;
    });
}

// ".mtvol": "MTVOL001" + u32 w, h, d (little endian) + u32 reserved, then w*h*d*4 bytes of RGBA8.
MtStatus mtxSaveVolume(const char* path, uint32_t w, uint32_t h, uint32_t d, const uint8_t* rgba8)
{
    return guarded([&]() -> MtStatus {
        if (!path || !rgba8 || !w || !h || !d) return MT_ERR_INVALID;
        FILE* f = std::fopen(path, "wb");
        if (!f) return MT_ERR_INVALID;
        const uint32_t hdr[4] = { w, h, d, 0 };
        const size_t n = (size_t)w * h * d * 4;
        const bool ok = std::fwrite("MTVOL001", 1, 8, f) == 8 && std::fwrite(hdr, 4, 4, f) == 4 && std::fwrite(rgba8, 1, n, f) == n;
        std::fclose(f);
        return ok ? MT_OK : MT_ERR_INVALID;
    });
}
MtStatus mtxLoadVolume(const char* path, uint8_t* rgba8_out, size_t out_bytes, uint32_t* w, uint32_t* h, uint32_t* d)
{
    return guarded([&]() -> MtStatus {
        if (!path || !w || !h || !d) return MT_ERR_INVALID;
        FILE* f = std::fopen(path, "rb");
        if (!f) return MT_ERR_INVALID;
        char magic[8];
        uint32_t hdr[4];
        bool ok = std::fread(magic, 1, 8, f) == 8 && !std::memcmp(magic, "MTVOL001", 8) && std::fread(hdr, 4, 4, f) == 4;
        if (ok) {
            *w = hdr[0]; *h = hdr[1]; *d = hdr[2];
            const size_t n = (size_t)hdr[0] * hdr[1] * hdr[2] * 4;
            if (rgba8_out) ok = out_bytes >= n && std::fread(rgba8_out, 1, n, f) == n;
        }
        std::fclose(f);
        return ok ? MT_OK : MT_ERR_INVALID;
    });
}

}  // extern "C"
