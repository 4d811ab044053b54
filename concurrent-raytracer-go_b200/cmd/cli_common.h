// cli_common.h — shared helpers of the CLI mirrors (cmd/raytracer.cpp, cmd/benchmark.cpp): PNG writer, Go-style
// duration strings, JSON escaping.  The CLIs only use the public C ABI (include/gort.h).
#pragma once
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <sys/stat.h>
#include <vector>

namespace cli {

// image/png Encode of an *image.RGBA (renderer.go:438-451): 8-bit RGBA, one IDAT, filter 0 per row.
inline bool write_png(const std::string& path, const uint8_t* rgba, int w, int h) {
    std::vector<uint8_t> raw((size_t)h * ((size_t)w * 4 + 1));
    for (int y = 0; y < h; y++) {
        raw[(size_t)y * (w * 4 + 1)] = 0;
        memcpy(&raw[(size_t)y * (w * 4 + 1) + 1], rgba + (size_t)y * w * 4, (size_t)w * 4);
    }
    uLongf zlen = compressBound((uLong)raw.size());
    std::vector<uint8_t> z(zlen);
    if (compress2(z.data(), &zlen, raw.data(), (uLong)raw.size(), 6) != Z_OK) return false;
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    auto be32 = [](uint8_t* p, uint32_t v) { p[0] = v >> 24; p[1] = v >> 16; p[2] = v >> 8; p[3] = v; };
    auto chunk = [&](const char* type, const uint8_t* data, uint32_t len) {
        uint8_t hdr[8];
        be32(hdr, len);
        memcpy(hdr + 4, type, 4);
        fwrite(hdr, 1, 8, f);
        if (len) fwrite(data, 1, len, f);
        uLong c = crc32(0L, (const Bytef*)type, 4);
        if (len) c = crc32(c, data, len);
        uint8_t tail[4];
        be32(tail, (uint32_t)c);
        fwrite(tail, 1, 4, f);
    };
    const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    fwrite(sig, 1, 8, f);
    uint8_t ihdr[13];
    be32(ihdr, (uint32_t)w);
    be32(ihdr + 4, (uint32_t)h);
    ihdr[8] = 8; ihdr[9] = 6; ihdr[10] = 0; ihdr[11] = 0; ihdr[12] = 0;
    chunk("IHDR", ihdr, 13);
    chunk("IDAT", z.data(), (uint32_t)zlen);
    chunk("IEND", nullptr, 0);
    return fclose(f) == 0;
}

// time.Duration.String(): "2.42585975s", "57.7ms", "8.958µs"
inline std::string go_duration(double seconds) {
    char buf[64];
    auto trim = [&](double v, const char* unit) {
        snprintf(buf, sizeof(buf), "%.9f", v);
        std::string s(buf);
        while (!s.empty() && s.back() == '0') s.pop_back();
        if (!s.empty() && s.back() == '.') s.pop_back();
        return s + unit;
    };
    if (seconds == 0) return "0s";
    if (seconds >= 60) {
        int m = (int)(seconds / 60);
        std::string s = std::to_string(m) + "m" + trim(seconds - 60.0 * m, "s");
        return s;
    }
    if (seconds >= 1) return trim(seconds, "s");
    if (seconds >= 1e-3) return trim(seconds * 1e3, "ms");
    if (seconds >= 1e-6) return trim(seconds * 1e6, "\xC2\xB5s");
    return trim(seconds * 1e9, "ns");
}

inline std::string json_str(const std::string& s) {
    std::string o = "\"";
    for (char c : s) {
        if (c == '"' || c == '\\') { o += '\\'; o += c; }
        else if (c == '\n') o += "\\n";
        else o += c;
    }
    return o + "\"";
}

inline std::string dir_of(const std::string& p) {
    size_t k = p.find_last_of('/');
    return k == std::string::npos ? "." : (k == 0 ? "/" : p.substr(0, k));
}

// os.MkdirAll: every missing level of the path (errors surface when the file is opened)
inline void mkdir_all(const std::string& dir) {
    for (size_t k = 1; k <= dir.size(); k++)
        if (k == dir.size() || dir[k] == '/') mkdir(dir.substr(0, k).c_str(), 0755);
}

}  // namespace cli
