// lmcma_layout.hpp — HBM layout of the cost map.  The map is stored in 128-byte bricks (one L1/L2 line),
// so that the ~1-cell-apart samples of a rasterised segment — which the cost kernel assigns to
// consecutive lanes — fall into few lines whatever the direction of travel (a row-major map costs one
// line per sample on a diagonal, and the L1 wavefront rate then bounds the gather).
//
//   2-D f32: brick  8 x 4      cells        2-D u8: brick 16 x 8     cells
//   3-D f32: brick  4 x 4 x 2  cells        3-D u8: brick  8 x 4 x 4 cells
//
// Inside a brick the cells are ordered x-major: index = (x_in * BY + y_in) * BZ + z_in.  Two reasons:
//  * the x term of the offset is then just `ix << log2(BY * BZ)` (its low bits are the in-brick column, its high
//    bits the brick column), and the y / z terms are one multiply-add each: `iy * BZ + (iy >> log2 BY) * py` with
//    `py = brick_row_pitch - BY * BZ` — 3 integer instructions per 2-D sample on the hot path instead of 9;
//  * a 32-byte sector (the unit L2 -> L1 moves) is a 2 x 4 (f32) / 4 x 8 (u8) tile instead of a 1-cell-high strip, so
//    a diagonal run of samples touches about half as many sectors.
//
// The logical (row-major) cell index ((z*ny + y)*nx + x) stays the API-visible index (lmcma_b200_cost_trace).
#pragma once
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define LMCMA_HD __host__ __device__ __forceinline__
#else
#define LMCMA_HD inline
#endif

namespace lmcma {

struct BrickShape { int bx, by, bz; };   // cells per brick along x, y, z

template <int DIMS, int STORAGE>
LMCMA_HD BrickShape brick_shape() {
    if (DIMS == 2) return STORAGE == 0 ? BrickShape{8, 4, 1} : BrickShape{16, 8, 1};
    return STORAGE == 0 ? BrickShape{4, 4, 2} : BrickShape{8, 4, 4};
}

// log2 of the brick extents
template <int DIMS, int STORAGE> struct BrickLog2 {
    static constexpr int X = DIMS == 2 ? (STORAGE == 0 ? 3 : 4) : (STORAGE == 0 ? 2 : 3);
    static constexpr int Y = DIMS == 2 ? (STORAGE == 0 ? 2 : 3) : 2;
    static constexpr int Z = DIMS == 2 ? 0 : (STORAGE == 0 ? 1 : 2);
};

// the two pitch constants of brick_offset (host: computed per call; device hot path: hoisted out of the sample loop)
template <int DIMS, int STORAGE>
LMCMA_HD unsigned brick_pitch_y(unsigned nbx) {
    typedef BrickLog2<DIMS, STORAGE> L;
    return (nbx << (L::X + L::Y + L::Z)) - (1u << (L::Y + L::Z));
}
template <int DIMS, int STORAGE>
LMCMA_HD unsigned brick_pitch_z(unsigned nbx, unsigned nby) {
    typedef BrickLog2<DIMS, STORAGE> L;
    return ((nbx * nby) << (L::X + L::Y + L::Z)) - (1u << L::Z);
}

// element offset of cell (ix, iy, iz) given the hoisted pitches; 32-bit offsets: a map holds fewer than 2^32 stored
// elements (checked at upload)
template <int DIMS, int STORAGE>
LMCMA_HD unsigned brick_offset_p(unsigned ix, unsigned iy, unsigned iz, unsigned py, unsigned pz) {
    typedef BrickLog2<DIMS, STORAGE> L;
    unsigned o = (ix << (L::Y + L::Z)) + (iy << L::Z) + (iy >> L::Y) * py;
    if (DIMS == 3) o += iz + (iz >> L::Z) * pz;
    return o;
}

// nbx / nby = bricks per row / per column
template <int DIMS, int STORAGE>
LMCMA_HD unsigned brick_offset(unsigned ix, unsigned iy, unsigned iz, unsigned nbx, unsigned nby) {
    return brick_offset_p<DIMS, STORAGE>(ix, iy, iz, brick_pitch_y<DIMS, STORAGE>(nbx), brick_pitch_z<DIMS, STORAGE>(nbx, nby));
}

}  // namespace lmcma
