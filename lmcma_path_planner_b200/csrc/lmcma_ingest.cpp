// lmcma_ingest.cpp — on-disk map formats that feed the hot path.  Host-side file parsing
// only (no compute): 24-bit BMP with the reference's obstacle rule, binvox run-length voxel grids, and the
// comma-separated distance matrix the reference reads into EDT_Matrix.  Written from the format descriptions:
//   BMP     Signed_Distance_Fields_test, planner.cpp:505-523 (SDL_LoadBMP + SDL_GetRGB; g < 128 -> obstacle)
//   binvox  binvox2bt.cpp:164-285 (header "#binvox 1 / dim d h w / translate / scale / data", then (value, count)
//           byte pairs; voxel i -> y = i % W, z = (i / W) % H, x = i / (W * H))
//   text    populate_EDT_Matrix_old, planner.cpp:777-818 (one row per line, "," separated)
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/lmcma_b200.h"

namespace lmcma {
int set_error(int code, const char* fmt, ...);   // lmcma_capi.cu
}
using lmcma::set_error;

namespace {

bool read_file(const char* path, std::vector<unsigned char>* out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    f.seekg(0, std::ios::end);
    const std::streamoff n = f.tellg();
    if (n < 0) return false;
    f.seekg(0);
    out->resize((size_t)n);
    if (n > 0) f.read(reinterpret_cast<char*>(out->data()), n);
    return (bool)f;
}
uint32_t le32(const unsigned char* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
uint16_t le16(const unsigned char* p) { return (uint16_t)(p[0] | (p[1] << 8)); }

}  // namespace

extern "C" {

int lmcma_b200_load_bmp(const char* path, uint8_t* occ_out, int64_t capacity, int32_t* width, int32_t* height) {
    if (!path || !width || !height) return set_error(LMCMA_B200_ERR_ARG, "null pointer");
    std::vector<unsigned char> raw;
    if (!read_file(path, &raw)) return set_error(LMCMA_B200_ERR_ARG, "cannot read %s", path);
    if (raw.size() < 54 || raw[0] != 'B' || raw[1] != 'M') return set_error(LMCMA_B200_ERR_ARG, "%s is not a BMP file", path);
    const uint32_t data_off = le32(&raw[10]);
    const int32_t w = (int32_t)le32(&raw[18]), h_signed = (int32_t)le32(&raw[22]);
    const int bpp = le16(&raw[28]);
    const uint32_t compression = le32(&raw[30]);
    if ((bpp != 24 && bpp != 32) || compression != 0) return set_error(LMCMA_B200_ERR_ARG, "%s: only uncompressed 24/32-bit BMP is supported", path);
    const int32_t h = h_signed < 0 ? -h_signed : h_signed;
    if (w <= 0 || h <= 0) return set_error(LMCMA_B200_ERR_ARG, "%s: bad dimensions", path);
    *width = w; *height = h;
    if (!occ_out) return 0;                                     // size query
    if (capacity < (int64_t)w * h) return set_error(LMCMA_B200_ERR_ARG, "capacity %lld < %lld cells", (long long)capacity, (long long)w * h);
    const size_t bytes_pp = bpp / 8, stride = ((size_t)w * bytes_pp + 3) & ~(size_t)3;
    if (raw.size() < data_off + stride * (size_t)h) return set_error(LMCMA_B200_ERR_ARG, "%s is truncated", path);
    for (int32_t row = 0; row < h; ++row) {
        const int32_t y = h_signed > 0 ? h - 1 - row : row;     // positive height: bottom-up storage
        const unsigned char* line = &raw[data_off + stride * (size_t)row];
        for (int32_t x = 0; x < w; ++x) {
            const unsigned char g = line[x * bytes_pp + 1];     // B G R [A]
            occ_out[(size_t)y * w + x] = g < 128 ? 1 : 0;       // planner.cpp:515
        }
    }
    return 0;
}

int lmcma_b200_load_binvox(const char* path, uint8_t* occ_out, int64_t capacity, int32_t* shape_xyz, double* translate_xyz,
                           double* scale) {
    if (!path || !shape_xyz) return set_error(LMCMA_B200_ERR_ARG, "null pointer");
    std::vector<unsigned char> raw;
    if (!read_file(path, &raw)) return set_error(LMCMA_B200_ERR_ARG, "cannot read %s", path);
    // ---- text header up to the line "data" ----
    size_t pos = 0;
    auto token = [&]() -> std::string {
        while (pos < raw.size() && (raw[pos] == ' ' || raw[pos] == '\n' || raw[pos] == '\r' || raw[pos] == '\t')) ++pos;
        const size_t b = pos;
        while (pos < raw.size() && !(raw[pos] == ' ' || raw[pos] == '\n' || raw[pos] == '\r' || raw[pos] == '\t')) ++pos;
        return std::string(raw.begin() + b, raw.begin() + pos);
    };
    if (token() != "#binvox") return set_error(LMCMA_B200_ERR_ARG, "%s: first token is not #binvox", path);
    token();                                                    // version
    int depth = -1, hgt = -1, wid = -1;
    double tr[3] = {0, 0, 0}, sc = 1.0;
    bool have_data = false;
    while (pos < raw.size()) {
        const std::string key = token();
        if (key == "data") { have_data = true; break; }
        if (key == "dim") { depth = atoi(token().c_str()); hgt = atoi(token().c_str()); wid = atoi(token().c_str()); }
        else if (key == "translate") { for (int c = 0; c < 3; ++c) tr[c] = atof(token().c_str()); }
        else if (key == "scale") { sc = atof(token().c_str()); }
        else { while (pos < raw.size() && raw[pos] != '\n') ++pos; }   // unknown keyword: skip the line
    }
    if (!have_data || depth <= 0 || hgt <= 0 || wid <= 0) return set_error(LMCMA_B200_ERR_ARG, "%s: bad binvox header", path);
    if (pos < raw.size() && raw[pos] == '\r') ++pos;
    if (pos < raw.size() && raw[pos] == '\n') ++pos;             // the linefeed after "data"
    // voxel i -> y = i % W, z = (i / W) % H, x = i / (W * H): x spans `depth`, z spans `height`, y spans `width`
    shape_xyz[0] = depth; shape_xyz[1] = wid; shape_xyz[2] = hgt;
    if (translate_xyz) { translate_xyz[0] = tr[0]; translate_xyz[1] = tr[1]; translate_xyz[2] = tr[2]; }
    if (scale) *scale = sc;
    if (!occ_out) return 0;                                     // size query
    const int64_t size = (int64_t)depth * hgt * wid;
    if (capacity < size) return set_error(LMCMA_B200_ERR_ARG, "capacity %lld < %lld voxels", (long long)capacity, (long long)size);
    const int nx = depth, ny = wid;                             // dense output [z][y][x]
    int64_t index = 0;
    while (index < size && pos + 1 < raw.size()) {
        const unsigned char value = raw[pos], count = raw[pos + 1];
        pos += 2;
        if (index + count > size) return set_error(LMCMA_B200_ERR_ARG, "%s: run-length data overruns the grid", path);
        for (int64_t i = index; i < index + count; ++i) {
            const int64_t y = i % wid, z = (i / wid) % hgt, x = i / ((int64_t)wid * hgt);
            occ_out[((size_t)z * ny + (size_t)y) * nx + (size_t)x] = value ? 1 : 0;
        }
        index += count;
    }
    if (index != size) return set_error(LMCMA_B200_ERR_ARG, "%s: run-length data ends after %lld of %lld voxels", path, (long long)index, (long long)size);
    return 0;
}

int lmcma_b200_load_text_matrix(const char* path, double* out, int64_t capacity, int32_t* rows, int32_t* cols) {
    if (!path || !rows || !cols) return set_error(LMCMA_B200_ERR_ARG, "null pointer");
    std::ifstream f(path);
    if (!f) return set_error(LMCMA_B200_ERR_ARG, "cannot read %s", path);
    std::vector<double> vals;
    std::string line;
    int r = 0, c_first = -1;
    while (std::getline(f, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.find_first_not_of(" \t") == std::string::npos) continue;
        int c = 0;
        size_t b = 0;
        for (;;) {
            const size_t e = line.find(',', b);
            const std::string tok = line.substr(b, e == std::string::npos ? std::string::npos : e - b);
            char* endp = nullptr;
            const double v = strtod(tok.c_str(), &endp);
            if (endp == tok.c_str()) return set_error(LMCMA_B200_ERR_ARG, "%s: row %d, column %d is not a number", path, r, c);
            vals.push_back(v);
            ++c;
            if (e == std::string::npos) break;
            b = e + 1;
        }
        if (c_first < 0) c_first = c;
        else if (c != c_first) return set_error(LMCMA_B200_ERR_ARG, "%s: row %d has %d columns, expected %d", path, r, c, c_first);
        ++r;
    }
    if (r == 0) return set_error(LMCMA_B200_ERR_ARG, "%s holds no numbers", path);
    *rows = r; *cols = c_first;
    if (!out) return 0;
    if (capacity < (int64_t)vals.size()) return set_error(LMCMA_B200_ERR_ARG, "capacity too small");
    memcpy(out, vals.data(), vals.size() * sizeof(double));
    return 0;
}

}  // extern "C"
