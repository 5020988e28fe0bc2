// lmcma_layout.hpp — HBM layout of the cost map.  The map is stored in 128-byte bricks (one L1/L2 line)
// whose 32-byte sectors are themselves small boxes, so that the ~1-cell-apart samples of a rasterised
// segment — which the cost kernel assigns to consecutive lanes — fall into few lines and few sectors
// whatever the direction of travel (a row-major map costs one line per sample on a diagonal).
//
//   2-D f32: brick  8 x 4      cells, sector 4 x 2        2-D u8: brick 16 x 8     cells, sector 8 x 4
//   3-D f32: brick  4 x 4 x 2  cells, sector 2 x 2 x 2    3-D u8: brick  8 x 4 x 4 cells, sector 4 x 4 x 2
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

// element offset of cell (ix, iy, iz) in the bricked array; nbx / nby = bricks per row / per column
template <int DIMS, int STORAGE>
LMCMA_HD size_t brick_offset(unsigned ix, unsigned iy, unsigned iz, unsigned nbx, unsigned nby) {
    if (DIMS == 2) {
        if (STORAGE == 0) {
            const size_t brick = (size_t)(iy >> 2) * nbx + (ix >> 3);
            const unsigned o = (ix & 3u) | ((iy & 1u) << 2) | (((ix >> 2) & 1u) << 3) | (((iy >> 1) & 1u) << 4);
            return brick * 32 + o;
        } else {
            const size_t brick = (size_t)(iy >> 3) * nbx + (ix >> 4);
            const unsigned o = (ix & 7u) | ((iy & 3u) << 3) | (((ix >> 3) & 1u) << 5) | (((iy >> 2) & 1u) << 6);
            return brick * 128 + o;
        }
    } else {
        if (STORAGE == 0) {
            const size_t brick = ((size_t)(iz >> 1) * nby + (iy >> 2)) * nbx + (ix >> 2);
            const unsigned o = (ix & 1u) | ((iy & 1u) << 1) | ((iz & 1u) << 2) | (((ix >> 1) & 1u) << 3) | (((iy >> 1) & 1u) << 4);
            return brick * 32 + o;
        } else {
            const size_t brick = ((size_t)(iz >> 2) * nby + (iy >> 2)) * nbx + (ix >> 3);
            const unsigned o = (ix & 3u) | ((iy & 3u) << 2) | ((iz & 1u) << 4) | (((ix >> 2) & 1u) << 5) | (((iz >> 1) & 1u) << 6);
            return brick * 128 + o;
        }
    }
}

}  // namespace lmcma
