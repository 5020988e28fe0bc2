// lmcma_layout.hpp — HBM layout of the cost map.  The map is stored in 128-byte bricks (one L1/L2 line),
// so that the ~1-cell-apart samples of a rasterised segment — which the cost kernel assigns to
// consecutive lanes — fall into few lines whatever the direction of travel (a row-major map costs one
// line per sample on a diagonal, and the L1 wavefront rate then bounds the gather).
//
//   2-D f32: brick  8 x 4      cells        2-D u8: brick 16 x 8     cells
//   3-D f32: brick  4 x 4 x 2  cells        3-D u8: brick  8 x 4 x 4 cells
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

// element offset of cell (ix, iy, iz) in the bricked array; nbx / nby = bricks per row / per column.
// Cells are row-major inside a brick (a handful of shifts and masks on the hot path); 32-bit offsets:
// a map holds fewer than 2^32 stored elements (checked at upload).
template <int DIMS, int STORAGE>
LMCMA_HD unsigned brick_offset(unsigned ix, unsigned iy, unsigned iz, unsigned nbx, unsigned nby) {
    if (DIMS == 2) {
        if (STORAGE == 0) return (((iy >> 2) * nbx + (ix >> 3)) << 5) | ((iy & 3u) << 3) | (ix & 7u);
        return (((iy >> 3) * nbx + (ix >> 4)) << 7) | ((iy & 7u) << 4) | (ix & 15u);
    } else {
        if (STORAGE == 0) return ((((iz >> 1) * nby + (iy >> 2)) * nbx + (ix >> 2)) << 5) | ((iz & 1u) << 4) | ((iy & 3u) << 2) | (ix & 3u);
        return ((((iz >> 2) * nby + (iy >> 2)) * nbx + (ix >> 3)) << 7) | ((iz & 3u) << 5) | ((iy & 3u) << 3) | (ix & 7u);
    }
}

}  // namespace lmcma
