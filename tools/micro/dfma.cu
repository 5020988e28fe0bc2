// Micro-benchmark: FP64 fused multiply-add latency and per-SM throughput (one CTA on one SM), next to FP32.
// Settles what bounds k_coef (k_gram.cuh): the FP64 pipe or the dependent chains.  (tools/, not part of the library)
#include <cuda_runtime.h>
#include <cstdio>
template <typename T, int CH>
__global__ void k(T* out, long long* cyc, int iters, T a, T b) {
    T x[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) x[c] = (T)(threadIdx.x + c);
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CH; ++c) x[c] = x[c] * a + b;          // contracted to one FMA per chain and iteration
    }
    const long long t1 = clock64();
    T s = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) s += x[c];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
template <typename T, int CH>
void run(const char* name, int threads) {
    T* out; long long* cyc; cudaMalloc(&out, 1024 * sizeof(T)); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    k<T, CH><<<1, threads>>>(out, cyc, iters, (T)1.0000001, (T)1e-9);
    k<T, CH><<<1, threads>>>(out, cyc, iters, (T)1.0000001, (T)1e-9);
    long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_iter = (double)h / iters;
    printf("%-5s threads %4d chains %d: %8.2f cycles / iteration  -> %.2f FMA / clock / SM, %.1f cycles per dependent FMA if latency-bound\n", name, threads, CH,
           per_iter, (double)threads * CH / per_iter, per_iter / 1.0);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<double, 1>("fp64", 32); run<double, 1>("fp64", 640); run<double, 4>("fp64", 640); run<double, 8>("fp64", 1024);
    run<float, 1>("fp32", 32); run<float, 8>("fp32", 1024);
    return 0;
}
