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

// OctoMap binary tree (".bt", written by OcTree::writeBinary — binvox2bt.cpp:287-300 in the reference, read there by
// planner.cpp:152-163).  Text header up to the line "data" ("id OcTree", "size <nodes>", "res <leaf size>"), then the tree
// depth first: every inner node is 2 bytes = 8 x 2 bits, child i in bits (2i, 2i+1) of the little-endian bit order
// std::bitset uses: 00 no child (unknown), 10 free leaf, 01 occupied leaf, 11 inner child (its own 2 bytes follow, in child
// order).  Depth 16, child index = x | y << 1 | z << 2 of the key bit at that level; a leaf above the last level stands for
// the whole pruned cube below it.  Key k of an axis is the cell [ (k - 32768) res, (k - 32767) res ).
// Output: dense occupancy [z][y][x] over the bounding box of the OCCUPIED leaves (1 = occupied, 0 = free or unknown),
// the key of its first cell per axis and the leaf size.
namespace {
struct BtLeaf { uint32_t key[3]; uint32_t size; };
bool bt_walk(const std::vector<unsigned char>& raw, size_t* pos, uint32_t kx, uint32_t ky, uint32_t kz, int depth, std::vector<BtLeaf>* occ,
             int64_t* nodes) {
    if (*pos + 2 > raw.size()) return false;
    const unsigned bits = (unsigned)raw[*pos] | ((unsigned)raw[*pos + 1] << 8);
    *pos += 2;
    ++*nodes;
    const uint32_t half = 1u << (15 - depth);                  // children of a depth-`depth` node are cubes of `half` leaves
    for (int i = 0; i < 8; ++i) {
        const unsigned c = (bits >> (2 * i)) & 3u;              // bit 2i -> low bit
        if (c == 0u) continue;
        const uint32_t cx = kx + ((i & 1) ? half : 0u), cy = ky + ((i & 2) ? half : 0u), cz = kz + ((i & 4) ? half : 0u);
        ++*nodes;
        if (c == 2u) occ->push_back(BtLeaf{{cx, cy, cz}, half});   // bits (2i, 2i+1) = (0, 1): occupied leaf
        else if (c == 3u) {                                        // inner child
            --*nodes;                                              // counted when its own bytes are read
            if (depth + 1 >= 16) return false;
            if (!bt_walk(raw, pos, cx, cy, cz, depth + 1, occ, nodes)) return false;
        }                                                          // c == 1: bits (1, 0): free leaf
    }
    return true;
}
}  // namespace

int lmcma_b200_load_bt(const char* path, uint8_t* occ_out, int64_t capacity, int32_t* shape_xyz, int32_t* origin_key_xyz, double* res) {
    if (!path || !shape_xyz) return set_error(LMCMA_B200_ERR_ARG, "null pointer");
    std::vector<unsigned char> raw;
    if (!read_file(path, &raw)) return set_error(LMCMA_B200_ERR_ARG, "cannot read %s", path);
    size_t pos = 0;
    auto line = [&]() -> std::string {
        const size_t b = pos;
        while (pos < raw.size() && raw[pos] != '\n') ++pos;
        std::string l(raw.begin() + b, raw.begin() + pos);
        if (pos < raw.size()) ++pos;
        if (!l.empty() && l.back() == '\r') l.pop_back();
        return l;
    };
    if (line().compare(0, 28, "# Octomap OcTree binary file") != 0) return set_error(LMCMA_B200_ERR_ARG, "%s: not an OctoMap binary tree", path);
    double leaf = 0.0; long long size = -1; bool have_data = false, id_ok = false;
    while (pos < raw.size()) {
        const std::string l = line();
        if (l == "data") { have_data = true; break; }
        if (l.empty() || l[0] == '#') continue;
        if (l.compare(0, 3, "id ") == 0) id_ok = l.substr(3) == "OcTree";
        else if (l.compare(0, 5, "size ") == 0) size = atoll(l.c_str() + 5);
        else if (l.compare(0, 4, "res ") == 0) leaf = atof(l.c_str() + 4);
    }
    if (!have_data || !id_ok || leaf <= 0.0) return set_error(LMCMA_B200_ERR_ARG, "%s: bad OcTree header", path);
    std::vector<BtLeaf> occ;
    int64_t nodes = 0;
    if (size != 0 && !bt_walk(raw, &pos, 0u, 0u, 0u, 0, &occ, &nodes)) return set_error(LMCMA_B200_ERR_ARG, "%s: truncated or malformed tree data", path);
    if (size > 0 && nodes != size) return set_error(LMCMA_B200_ERR_ARG, "%s: header says %lld nodes, the data holds %lld", path, size, (long long)nodes);
    if (occ.empty()) return set_error(LMCMA_B200_ERR_ARG, "%s holds no occupied leaf", path);
    uint32_t lo[3] = {~0u, ~0u, ~0u}, hi[3] = {0u, 0u, 0u};
    for (const BtLeaf& l : occ)
        for (int c = 0; c < 3; ++c) { lo[c] = std::min(lo[c], l.key[c]); hi[c] = std::max(hi[c], l.key[c] + l.size); }
    for (int c = 0; c < 3; ++c) shape_xyz[c] = (int32_t)(hi[c] - lo[c]);
    if (origin_key_xyz) for (int c = 0; c < 3; ++c) origin_key_xyz[c] = (int32_t)lo[c];
    if (res) *res = leaf;
    if (!occ_out) return 0;                                     // size query
    const int64_t nx = shape_xyz[0], ny = shape_xyz[1], nz = shape_xyz[2];
    if (capacity < nx * ny * nz) return set_error(LMCMA_B200_ERR_ARG, "capacity %lld < %lld cells", (long long)capacity, (long long)(nx * ny * nz));
    memset(occ_out, 0, (size_t)(nx * ny * nz));
    for (const BtLeaf& l : occ)
        for (uint32_t z = l.key[2] - lo[2]; z < l.key[2] - lo[2] + l.size; ++z)
            for (uint32_t y = l.key[1] - lo[1]; y < l.key[1] - lo[1] + l.size; ++y)
                memset(occ_out + ((size_t)z * ny + y) * nx + (l.key[0] - lo[0]), 1, l.size);
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
