// Micro-benchmark for the north-star question "should the cost map be staged with TMA?" (missing #4a):
// the same lane-consecutive gather along ~1-cell-per-sample poly-lines as gather_paths.cu (1024 poly-lines x 8192 samples,
// 4096^2 map), on the u8 storage, three ways:
//   LDG u8 bricked 16x8   the library's layout (lmcma_layout.hpp): one byte load per sample, lines of 16 x 8 cells
//   LDG u8 row-major      plain rows
//   TMA tiles -> smem     cp.async.bulk.tensor.2d (SASS UTMALDG): every run of 64 consecutive samples is served from a BOX x BOX
//                         u8 tile fetched into shared memory by one TMA request placed at the run's first cell (double-buffered
//                         per warp, the next run's tile in flight while this one is consumed); BOX = 64 (the 64 x 64 tile the
//                         verdict names, 4 KB per 64 samples) and BOX = 48 (the smallest 16-byte-granular box that always covers a run from a 16-byte aligned origin, 2.25 KB)
// Build: make -C tools/micro gather_tma ; run on the GPU box.  (tools/, not part of the library)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
constexpr int N = 4096, TRAJ = 1024, S = 8192, RUN = 64;

__device__ __forceinline__ void cell(int t, int traj, float slope, int& ix, int& iy) {
    const float off = (float)(traj & 63) * 0.37f;
    const float x = 64.f + off + (float)t * (3900.f / S);          // < 1 cell per sample
    const float y = 64.f + (float)(traj >> 6) * 1.3f + (float)t * (3900.f / S) * slope;
    ix = __float2int_rn(x); iy = __float2int_rn(y);
}
__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }

template <int MODE>
__global__ void __launch_bounds__(256, 7) k_ldg(const unsigned char* __restrict__ brick, const unsigned char* __restrict__ rowmaj, float slope, float* out) {
    const int traj = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc = 0.f;
    const int per = S / 8;
#pragma unroll 4
    for (int t0 = warp * per; t0 < (warp + 1) * per; t0 += 32) {
        int ix, iy; cell(t0 + lane, traj, slope, ix, iy);
        unsigned char g;
        // 16 x 8 bricks, x-major inside a brick: offset = (ix << 3) + (iy & 7) + (iy >> 3) * (N * 8)
        if (MODE == 0) g = __ldg(brick + (((unsigned)ix << 3) + ((unsigned)iy & 7u) + ((unsigned)iy >> 3) * (unsigned)(N * 8)));
        else g = __ldg(rowmaj + (unsigned)iy * N + ix);
        acc += (float)g;
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(~0u, acc, o);
    if (lane == 0) atomicAdd(out + traj, acc);
}

template <int BOX>
__global__ void __launch_bounds__(256) k_tma(const __grid_constant__ CUtensorMap tmap, float slope, float* out, int align_x) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar[8][2];
    const int traj = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* mine = smem + (size_t)warp * 2 * BOX * BOX;
    if (lane == 0) {
        for (int s = 0; s < 2; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[warp][s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const int runs = S / RUN / 8;                                  // runs of 64 samples per warp (warp w: runs w, w + 8, ...)
    auto origin = [&](int run, int& x0, int& y0) {                 // first cell of the run (slope >= 0: the run stays inside the box)
        cell(run * RUN, traj, slope, x0, y0);
        if (align_x) x0 &= ~15;                                    // 16-byte aligned tile origin (u8)
    };
    auto issue = [&](int k) {                                      // lane 0
        int x0, y0; origin(warp + 8 * k, x0, y0);
        const unsigned dst = smem_u32(mine + (size_t)(k & 1) * BOX * BOX), mb = smem_u32(&bar[warp][k & 1]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(BOX * BOX) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(dst), "l"(&tmap), "r"(x0), "r"(y0), "r"(mb) : "memory");
    };
    if (lane == 0) issue(0);
    float acc = 0.f;
    for (int k = 0; k < runs; ++k) {
        if (lane == 0 && k + 1 < runs) issue(k + 1);               // the other stage was consumed in iteration k - 1 (__syncwarp below)
        const unsigned mb = smem_u32(&bar[warp][k & 1]), parity = (unsigned)((k >> 1) & 1);
        unsigned ok = 0;
        while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(mb), "r"(parity) : "memory");
        int x0, y0; origin(warp + 8 * k, x0, y0);
        const unsigned char* tile = mine + (size_t)(k & 1) * BOX * BOX;
#pragma unroll
        for (int h = 0; h < RUN / 32; ++h) {
            int ix, iy; cell((warp + 8 * k) * RUN + h * 32 + lane, traj, slope, ix, iy);
            acc += (float)tile[(iy - y0) * BOX + (ix - x0)];
        }
        __syncwarp();
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(~0u, acc, o);
    if (lane == 0) atomicAdd(out + traj, acc);
}

typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                              const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    const int only = argc > 1 ? atoi(argv[1]) : -1;               // run one mode per process (a faulting mode poisons the context)
    const int align_x = argc > 2 ? atoi(argv[2]) : 1;
    std::vector<unsigned char> h((size_t)N * N), hb((size_t)N * N);
    for (int y = 0; y < N; ++y) for (int x = 0; x < N; ++x) {
        const unsigned char v = (unsigned char)(1 + (x * 7 + y * 13) % 97);
        h[(size_t)y * N + x] = v;
        hb[((size_t)x << 3) + (y & 7) + (size_t)(y >> 3) * (N * 8)] = v;
    }
    unsigned char *drow, *dbrick; float* dout;
    CK(cudaMalloc(&drow, h.size())); CK(cudaMalloc(&dbrick, h.size())); CK(cudaMalloc(&dout, TRAJ * 4 * 4));
    CK(cudaMemcpy(drow, h.data(), h.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dbrick, hb.data(), h.size(), cudaMemcpyHostToDevice));
    void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
    if (!fn) { printf("cuTensorMapEncodeTiled not available\n"); return 1; }
    CUtensorMap tm64, tm32;
    for (int which = 0; which < 2; ++which) {
        const cuuint32_t boxdim = which == 0 ? 64 : 48;
        cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)N}, strides[1] = {(cuuint64_t)N};
        cuuint32_t box[2] = {boxdim, boxdim}, estr[2] = {1, 1};
        CUresult r = ((encode_fn)fn)(which == 0 ? &tm64 : &tm32, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, drow, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); return 1; }
    }
    CK(cudaFuncSetAttribute(k_tma<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * 64 * 64));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char* names[4] = {"LDG u8 bricked 16x8", "LDG u8 row-major", "TMA 64x64 u8 tiles -> smem", "TMA 48x48 u8 tiles -> smem"};
    for (float slope : {1.0f, 0.5f, 0.05f}) {
        printf("slope %.2f\n", slope);
        for (int mode = 0; mode < 4; ++mode) {
            if (only >= 0 && mode != only) continue;
            CK(cudaMemset(dout, 0, TRAJ * 4 * 4));
            float best = 1e9f;
            for (int rep = 0; rep < 6; ++rep) {
                cudaEventRecord(e0);
                if (mode == 0) k_ldg<0><<<TRAJ, 256>>>(dbrick, drow, slope, dout);
                if (mode == 1) k_ldg<1><<<TRAJ, 256>>>(dbrick, drow, slope, dout + TRAJ);
                if (mode == 2) k_tma<64><<<TRAJ, 256, 8 * 2 * 64 * 64>>>(tm64, slope, dout + 2 * TRAJ, align_x);
                if (mode == 3) k_tma<48><<<TRAJ, 256, 8 * 2 * 48 * 48>>>(tm32, slope, dout + 3 * TRAJ, align_x);
                cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
            }
            CK(cudaGetLastError());
            std::vector<float> o(TRAJ); CK(cudaMemcpy(o.data(), dout + mode * TRAJ, TRAJ * 4, cudaMemcpyDeviceToHost));
            const double tile_mb = mode == 2 ? (double)TRAJ * (S / RUN) * 64 * 64 / 1e6 : (mode == 3 ? (double)TRAJ * (S / RUN) * 48 * 48 / 1e6 : 0.0);
            printf("  %-28s %8.2f us   %.1f Gsamples/s   check %.1f   (tile bytes requested %.0f MB)\n", names[mode], best * 1e3,
                   (double)TRAJ * S / best * 1e-6, o[5] / 5.0, tile_mb);
        }
    }
    return 0;
}
