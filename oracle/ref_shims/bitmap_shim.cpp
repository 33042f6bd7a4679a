/* oracle/_ref only -- replaces the OpenEXR-backed src/bitmap.cpp:33-139 of the reference with a
 * dependency-free implementation of the SAME nori::Bitmap API (include/nori/bitmap.h:32-54):
 *   Bitmap(filename)  : reads scan-line OpenEXR (NONE / ZIPS / ZIP compression, HALF or FLOAT R,G,B)
 *   save(filename)    : writes an uncompressed FLOAT R,G,B scan-line OpenEXR
 *   saveToLDR(file)   : sRGB PNG through ext/stb_image_write.h, as the reference does
 * OpenEXR itself cannot be built here (CMake 4 rejects ext/openexr); nothing on the rendering hot
 * path depends on it.  This file is test infrastructure and is original code. */
#include <nori/bitmap.h>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <zlib.h>
#define STB_IMAGE_WRITE_IMPLEMENTATION
#include <stb_image_write.h>

NORI_NAMESPACE_BEGIN

namespace {

struct Channel { std::string name; int type; };   // type: 0 uint, 1 half, 2 float

float halfToFloat(uint16_t h) {
    uint32_t s = (h >> 15) & 1, e = (h >> 10) & 31, m = h & 1023, out;
    if (e == 0) {
        if (m == 0) out = s << 31;
        else { e = 127 - 15 + 1; while (!(m & 1024)) { m <<= 1; --e; } out = (s << 31) | (e << 23) | ((m & 1023) << 13); }
    } else if (e == 31) out = (s << 31) | 0x7f800000u | (m << 13);
    else out = (s << 31) | ((e + 112) << 23) | (m << 13);
    float f; memcpy(&f, &out, 4); return f;
}

std::string readStr(const std::vector<uint8_t> &b, size_t &p) {
    std::string s; while (b[p]) s += (char) b[p++]; ++p; return s;
}
template <typename T> T rd(const std::vector<uint8_t> &b, size_t p) { T v; memcpy(&v, &b[p], sizeof(T)); return v; }

void unzipBlock(const uint8_t *src, size_t n, std::vector<uint8_t> &out) {
    std::vector<uint8_t> tmp(out.size());
    uLongf len = tmp.size();
    if (uncompress(tmp.data(), &len, src, n) != Z_OK || len != tmp.size())
        throw NoriException("EXR shim: zlib failure");
    for (size_t i = 1; i < len; ++i) tmp[i] = (uint8_t) (tmp[i - 1] + tmp[i] - 128);   // predictor
    size_t half = (len + 1) / 2;                                                       // de-interleave
    for (size_t i = 0; i < len; ++i) out[i] = (i & 1) ? tmp[half + i / 2] : tmp[i / 2];
}

} // namespace

Bitmap::Bitmap(const std::string &filename) {
    std::ifstream is(filename, std::ios::binary);
    if (!is) throw NoriException("EXR shim: cannot open \"%s\"", filename);
    std::vector<uint8_t> b((std::istreambuf_iterator<char>(is)), std::istreambuf_iterator<char>());
    if (b.size() < 8 || rd<uint32_t>(b, 0) != 20000630u) throw NoriException("EXR shim: bad magic");
    if (rd<uint32_t>(b, 4) & 0x200) throw NoriException("EXR shim: tiled files unsupported");
    size_t p = 8;
    std::vector<Channel> ch; int comp = 0; int dw[4] = {0, 0, 0, 0};
    while (b[p]) {
        std::string name = readStr(b, p), type = readStr(b, p);
        uint32_t size = rd<uint32_t>(b, p); p += 4;
        if (name == "channels") {
            size_t q = p;
            while (b[q]) { Channel c; c.name = readStr(b, q); c.type = rd<int32_t>(b, q); q += 16; ch.push_back(c); }
        } else if (name == "compression") comp = b[p];
        else if (name == "dataWindow") for (int i = 0; i < 4; ++i) dw[i] = rd<int32_t>(b, p + 4 * i);
        p += size;
    }
    ++p;
    int W = dw[2] - dw[0] + 1, H = dw[3] - dw[1] + 1;
    resize(H, W);
    cout << "Reading a " << cols() << "x" << rows() << " OpenEXR file from \"" << filename << "\"" << endl;
    int linesPerBlock = comp == 3 ? 16 : 1;
    if (comp != 0 && comp != 2 && comp != 3) throw NoriException("EXR shim: unsupported compression %i", comp);
    size_t bytesPerLine = 0; for (auto &c : ch) bytesPerLine += (c.type == 1 ? 2 : 4) * (size_t) W;
    int nBlocks = (H + linesPerBlock - 1) / linesPerBlock;
    for (int blk = 0; blk < nBlocks; ++blk) {
        size_t off = (size_t) rd<uint64_t>(b, p + 8 * (size_t) blk);
        int y0 = rd<int32_t>(b, off) - dw[1]; uint32_t n = rd<uint32_t>(b, off + 4);
        int lines = std::min(linesPerBlock, H - y0);
        std::vector<uint8_t> raw(bytesPerLine * lines);
        if (comp == 0 || n == raw.size()) memcpy(raw.data(), &b[off + 8], raw.size());
        else unzipBlock(&b[off + 8], n, raw);
        for (int l = 0; l < lines; ++l) {
            size_t q = bytesPerLine * l;
            for (auto &c : ch) {
                std::string nm = toLower(c.name);
                int k = (nm == "r" || endsWith(nm, ".r")) ? 0 : (nm == "g" || endsWith(nm, ".g")) ? 1
                      : (nm == "b" || endsWith(nm, ".b")) ? 2 : -1;
                for (int x = 0; x < W; ++x) {
                    float v = c.type == 1 ? halfToFloat(rd<uint16_t>(raw, q + 2 * x)) : rd<float>(raw, q + 4 * x);
                    if (k >= 0) coeffRef(y0 + l, x)[k] = v;
                }
                q += (c.type == 1 ? 2 : 4) * (size_t) W;
            }
        }
    }
}

void Bitmap::save(const std::string &filename) {
    cout << "Writing a " << cols() << "x" << rows() << " OpenEXR file to \"" << filename << "\"" << endl;
    int W = (int) cols(), H = (int) rows();
    std::vector<uint8_t> o;
    auto put = [&](const void *d, size_t n) { o.insert(o.end(), (const uint8_t *) d, (const uint8_t *) d + n); };
    auto puts = [&](const char *s) { put(s, strlen(s) + 1); };
    auto puti = [&](int32_t v) { put(&v, 4); };
    uint32_t magic = 20000630u, version = 2; put(&magic, 4); put(&version, 4);
    puts("channels"); puts("chlist"); puti(3 * 18 + 1);
    for (const char *c : {"B", "G", "R"}) { puts(c); puti(2); puti(0); puti(1); puti(1); }
    o.push_back(0);
    puts("compression"); puts("compression"); puti(1); o.push_back(0);
    int box[4] = {0, 0, W - 1, H - 1};
    puts("dataWindow"); puts("box2i"); puti(16); put(box, 16);
    puts("displayWindow"); puts("box2i"); puti(16); put(box, 16);
    puts("lineOrder"); puts("lineOrder"); puti(1); o.push_back(0);
    float one = 1.f, zero2[2] = {0.f, 0.f};
    puts("pixelAspectRatio"); puts("float"); puti(4); put(&one, 4);
    puts("screenWindowCenter"); puts("v2f"); puti(8); put(zero2, 8);
    puts("screenWindowWidth"); puts("float"); puti(4); put(&one, 4);
    o.push_back(0);
    size_t lineBytes = 8 + 12 * (size_t) W, tableAt = o.size();
    for (int y = 0; y < H; ++y) { uint64_t off = tableAt + 8 * (size_t) H + lineBytes * y; put(&off, 8); }
    std::vector<float> row(3 * (size_t) W);
    for (int y = 0; y < H; ++y) {
        puti(y); puti(12 * W);
        for (int x = 0; x < W; ++x) {
            row[x] = coeff(y, x)[2]; row[W + x] = coeff(y, x)[1]; row[2 * W + x] = coeff(y, x)[0];
        }
        put(row.data(), 12 * (size_t) W);
    }
    std::ofstream os(filename, std::ios::binary);
    os.write((const char *) o.data(), o.size());
}

void Bitmap::saveToLDR(const std::string &filename) {
    cout << "Writing a " << cols() << "x" << rows() << " PNG file to \"" << filename << "\"" << endl;
    std::unique_ptr<uint8_t[]> rgb8(new uint8_t[3 * cols() * rows()]);
    uint8_t *dst = rgb8.get();
    for (int y = 0; y < rows(); ++y)
        for (int x = 0; x < cols(); ++x)
            for (int k = 0; k < 3; ++k) {
                float v = coeff(y, x)[k];
                v = v <= 0.0031308f ? 12.92f * v : 1.055f * std::pow(v, 1.f / 2.4f) - 0.055f;
                *dst++ = (uint8_t) std::min(255.f, std::max(0.f, 255.f * v + 0.5f));
            }
    stbi_write_png(filename.c_str(), (int) cols(), (int) rows(), 3, rgb8.get(), 3 * (int) cols());
}

NORI_NAMESPACE_END
