// Micro-benchmark: what bounds a lane-consecutive gather along ~1-cell-per-sample diagonal paths through a 4096^2 f32 map?
// LDG from the bricked layout vs point-sampled TEX vs SULD from a block-linear cudaArray.  (tools/, not part of the library)
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#include <cmath>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
constexpr int N = 4096, TRAJ = 1024, S = 8192;

__device__ __forceinline__ void cell(int t, int traj, float slope, int& ix, int& iy) {
    const float off = (float)(traj & 63) * 0.37f;
    const float x = 64.f + off + (float)t * (3900.f / S);          // < 1 cell per sample
    const float y = 64.f + (float)(traj >> 6) * 1.3f + (float)t * (3900.f / S) * slope;
    ix = __float2int_rn(x); iy = __float2int_rn(y);
}
template <int MODE>
__global__ void __launch_bounds__(256, 7) k(const float* __restrict__ brick, const float* __restrict__ rowmaj, cudaTextureObject_t tex,
                                            cudaSurfaceObject_t surf, float slope, float* out) {
    const int traj = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc = 0.f;
    const int per = S / 8;
#pragma unroll 4
    for (int t0 = warp * per; t0 < (warp + 1) * per; t0 += 32) {
        int ix, iy; cell(t0 + lane, traj, slope, ix, iy);
        float g;
        if (MODE == 0) g = __ldg(brick + (((unsigned)ix << 2) + (unsigned)iy + ((unsigned)iy >> 2) * (unsigned)(N / 8 * 32 - 4)));
        else if (MODE == 1) g = __ldg(rowmaj + (unsigned)iy * N + ix);
        else if (MODE == 2) g = tex2D<float>(tex, (float)ix + 0.5f, (float)iy + 0.5f);
        else if (MODE == 3) g = surf2Dread<float>(surf, ix * 4, iy, cudaBoundaryModeZero);
        // control: the SAME arithmetic, but the 32 lanes read 32 consecutive words (one line): what the loop costs without a gather
        else g = __ldg(rowmaj + ((((unsigned)iy * N + (unsigned)ix) & ~31u) + lane));
        acc += fabsf(g);
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(~0u, acc, o);
    if (lane == 0) atomicAdd(out + traj, acc);
}
// Two consecutive samples per lane (they are <= 1 cell apart in x and in y, so both lie in ONE 2 x 2 footprint):
//   PAIR 0: two LDGs from the bricked layout (control: same arithmetic, twice the fetches of PAIR 1)
//   PAIR 1: ONE tex2Dgather (the four texels of the bilinear footprint anchored at the smaller x / y) serves both
template <int PAIR>
__global__ void __launch_bounds__(256, 7) k2(const float* __restrict__ brick, cudaTextureObject_t texg, float slope, float* out) {
    const int traj = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc = 0.f;
    const int per = S / 8;
#pragma unroll 2
    for (int t0 = warp * per; t0 < (warp + 1) * per; t0 += 64) {
        int ax, ay, bx, by;
        cell(t0 + 2 * lane, traj, slope, ax, ay);
        cell(t0 + 2 * lane + 1, traj, slope, bx, by);
        float ga, gb;
        if (PAIR == 0) {
            ga = __ldg(brick + (((unsigned)ax << 2) + (unsigned)ay + ((unsigned)ay >> 2) * (unsigned)(N / 8 * 32 - 4)));
            gb = __ldg(brick + (((unsigned)bx << 2) + (unsigned)by + ((unsigned)by >> 2) * (unsigned)(N / 8 * 32 - 4)));
        } else {
            const int x0 = min(ax, bx), y0 = min(ay, by);
            const float4 q = tex2Dgather<float4>(texg, (float)x0 + 1.0f, (float)y0 + 1.0f, 0);   // w (x0,y0)  z (x0+1,y0)  x (x0,y0+1)  y (x0+1,y0+1)
            const float lo_a = (ax == x0) ? q.w : q.z, hi_a = (ax == x0) ? q.x : q.y;
            const float lo_b = (bx == x0) ? q.w : q.z, hi_b = (bx == x0) ? q.x : q.y;
            ga = (ay == y0) ? lo_a : hi_a;
            gb = (by == y0) ? lo_b : hi_b;
        }
        acc += fabsf(ga) + fabsf(gb);
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(~0u, acc, o);
    if (lane == 0) atomicAdd(out + traj, acc);
}
int main() {
    std::vector<float> h((size_t)N * N), hb((size_t)N * N);
    for (int y = 0; y < N; ++y) for (int x = 0; x < N; ++x) {
        const float v = 1.f / (1.f + (float)((x * 7 + y * 13) % 97));
        h[(size_t)y * N + x] = v;
        hb[((size_t)x << 2) + y + (size_t)(y >> 2) * (N / 8 * 32 - 4)] = v;
    }
    float *drow, *dbrick, *dout;
    CK(cudaMalloc(&drow, h.size() * 4)); CK(cudaMalloc(&dbrick, h.size() * 4)); CK(cudaMalloc(&dout, TRAJ * 4 * 4));
    CK(cudaMemcpy(drow, h.data(), h.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dbrick, hb.data(), h.size() * 4, cudaMemcpyHostToDevice));
    cudaChannelFormatDesc fd = cudaCreateChannelDesc<float>();
    cudaArray_t arr; CK(cudaMallocArray(&arr, &fd, N, N, cudaArraySurfaceLoadStore));
    CK(cudaMemcpy2DToArray(arr, 0, 0, h.data(), N * 4, N * 4, N, cudaMemcpyHostToDevice));
    cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
    cudaTextureDesc td = {}; td.addressMode[0] = td.addressMode[1] = cudaAddressModeBorder; td.filterMode = cudaFilterModePoint; td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
    cudaTextureObject_t tex; CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
    cudaSurfaceObject_t surf; CK(cudaCreateSurfaceObject(&surf, &rd));
    cudaArray_t arrg; CK(cudaMallocArray(&arrg, &fd, N, N, cudaArrayTextureGather));
    CK(cudaMemcpy2DToArray(arrg, 0, 0, h.data(), N * 4, N * 4, N, cudaMemcpyHostToDevice));
    cudaResourceDesc rdg = {}; rdg.resType = cudaResourceTypeArray; rdg.res.array.array = arrg;
    cudaTextureDesc tdg = {}; tdg.addressMode[0] = tdg.addressMode[1] = cudaAddressModeClamp; tdg.filterMode = cudaFilterModePoint; tdg.readMode = cudaReadModeElementType; tdg.normalizedCoords = 0;
    cudaTextureObject_t texg; CK(cudaCreateTextureObject(&texg, &rdg, &tdg, nullptr));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char* names[5] = {"LDG bricked 8x4 x-major", "LDG row-major", "TEX point", "SULD", "control: coalesced LDG"};
    for (float slope : {1.0f, 0.5f, 0.05f}) {
        printf("slope %.2f\n", slope);
        for (int mode = 0; mode < 5; ++mode) {
            CK(cudaMemset(dout, 0, TRAJ * 4 * 4));
            float best = 1e9f;
            for (int rep = 0; rep < 6; ++rep) {
                cudaEventRecord(e0);
                if (mode == 0) k<0><<<TRAJ, 256>>>(dbrick, drow, tex, surf, slope, dout);
                if (mode == 1) k<1><<<TRAJ, 256>>>(dbrick, drow, tex, surf, slope, dout + TRAJ);
                if (mode == 2) k<2><<<TRAJ, 256>>>(dbrick, drow, tex, surf, slope, dout + 2 * TRAJ);
                if (mode == 3) k<3><<<TRAJ, 256>>>(dbrick, drow, tex, surf, slope, dout + 3 * TRAJ);
                if (mode == 4) k<4><<<TRAJ, 256>>>(dbrick, drow, tex, surf, slope, dout);
                cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
            }
            std::vector<float> o(TRAJ); CK(cudaMemcpy(o.data(), dout + (mode & 3) * TRAJ, TRAJ * 4, cudaMemcpyDeviceToHost));
            printf("  %-26s %8.2f us   %.1f Gsamples/s   check %.4f\n", names[mode], best * 1e3, (double)TRAJ * S / best * 1e-6, o[5] / 5.0);
        }
        const char* names2[2] = {"2 samples/lane: 2 x LDG", "2 samples/lane: 1 x TLD4"};
        for (int pair = 0; pair < 2; ++pair) {
            CK(cudaMemset(dout, 0, TRAJ * 4 * 4));
            float best = 1e9f;
            for (int rep = 0; rep < 6; ++rep) {
                cudaEventRecord(e0);
                if (pair == 0) k2<0><<<TRAJ, 256>>>(dbrick, texg, slope, dout);
                else k2<1><<<TRAJ, 256>>>(dbrick, texg, slope, dout);
                cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
            }
            std::vector<float> o(TRAJ); CK(cudaMemcpy(o.data(), dout, TRAJ * 4, cudaMemcpyDeviceToHost));
            printf("  %-26s %8.2f us   %.1f Gsamples/s   check %.4f\n", names2[pair], best * 1e3, (double)TRAJ * S / best * 1e-6, o[5] / 30.0);
        }
    }
    return 0;
}
