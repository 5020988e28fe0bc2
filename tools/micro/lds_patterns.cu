// Micro-benchmark: what a shared-memory load costs the L1tex data pipe when the lanes of a warp read a handful of DISTINCT
// records (k_cost's per-sample segment record: lanes before a segment end read record s, lanes behind it record s + 1).
// One CTA of 1024 threads on one SM, 8 independent loads per iteration and lane; cycles per warp-load at saturation.
// (tools/, not part of the library)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o lds_patterns lds_patterns.cu && ./lds_patterns
#include <cuda_runtime.h>
#include <cstdio>

// WIDTH: 4 / 8 / 16 bytes per lane.  pattern: the record index of a lane
//   0 all lanes the same record          1 two records, split at lane 16      2 two records, split at lane 11
//   3 three records (0..9, 10..20, 21..) 4 four records, one per quarter-warp 5 every lane its own record (32-byte stride)
//   6 SHFL.IDX of one register per "load" from the lane that holds the record (pattern 2's source lanes)
__device__ __forceinline__ int record_of(int pattern, int lane) {
    switch (pattern) {
        case 0: return 0;
        case 1: return lane >> 4;
        case 2: return lane > 10 ? 1 : 0;
        case 3: return lane > 20 ? 2 : (lane > 9 ? 1 : 0);
        case 4: return lane >> 3;
        default: return lane;
    }
}

template <int WIDTH>
__global__ void __launch_bounds__(1024, 1) k_lds(float* out, long long* cyc, int iters, int pattern) {
    __shared__ __align__(16) float rec[64 * 8 + 64];               // 32-byte records
    for (int i = threadIdx.x; i < 64 * 8 + 64; i += blockDim.x) rec[i] = (float)i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const float* base = rec + record_of(pattern, lane) * 8;
    float acc = 0.f;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float* p = base + ((it + u) & 7) * 8;            // walk over 8 consecutive records: no loop-invariant loads
            if (WIDTH == 16) { float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"((unsigned)__cvta_generic_to_shared(p))); acc += (v.x + v.y) + (v.z + v.w); }
            else if (WIDTH == 8) { float2 v; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"((unsigned)__cvta_generic_to_shared(p))); acc += v.x + v.y; }
            else { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"((unsigned)__cvta_generic_to_shared(p))); acc += v; }
        }
    }
    const long long t1 = clock64();
    out[threadIdx.x] = acc;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

__global__ void __launch_bounds__(1024, 1) k_shfl(float* out, long long* cyc, int iters) {
    const int lane = threadIdx.x & 31;
    const int src = lane > 10 ? 1 : 0;
    float r = (float)threadIdx.x, acc = 0.f;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += __shfl_sync(0xffffffffu, r + (float)u, (src + it + u) & 31);
    }
    const long long t1 = clock64();
    out[threadIdx.x] = acc;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

int main() {
    float* out; long long* cyc; cudaMalloc(&out, 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 2048;
    const char* names[6] = {"uniform", "2 records @16", "2 records @11", "3 records", "4 records (quarters)", "32 records"};
    for (int width = 16; width >= 4; width >>= 1)
        for (int p = 0; p < 6; ++p) {
            for (int rep = 0; rep < 2; ++rep) {
                if (width == 16) k_lds<16><<<1, 1024>>>(out, cyc, iters, p);
                else if (width == 8) k_lds<8><<<1, 1024>>>(out, cyc, iters, p);
                else k_lds<4><<<1, 1024>>>(out, cyc, iters, p);
            }
            long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            printf("LDS.%-3d %-22s %6.2f cycles per warp-load (32 warps, 8 loads in flight each)\n", width * 8, names[p], (double)h / ((double)iters * 8 * 32));
        }
    for (int rep = 0; rep < 2; ++rep) k_shfl<<<1, 1024>>>(out, cyc, iters);
    long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("SHFL.IDX (2 source lanes)      %6.2f cycles per warp-shuffle\n", (double)h / ((double)iters * 8 * 32));
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
