// k_gram.cuh — the recompute of the inverse-direction vectors (lmcma.cpp:373-390, 449-463) for shapes whose
// direction rows fit neither the registers nor the shared memory of one SM (C4: m = 77 rows of n = 1500).
// The factor-major sweep of k_update.cuh would then stream every row-step through L2 from a single SM
// (~1 ms).  Every vector of the sweep is a linear combination of the L basis rows
//     b_i = v_i (i < first_stale, unchanged)   |   b_i = pc_i (i >= first_stale, to be recomputed),
// so the whole recurrence can be carried out on L-vectors of coefficients once the Gram matrix G = B B^T is known:
//     v_j . x  =  coef(v_j)^T G coef(x).
//   k_gram     G in FP64 from the FP32 rows, many CTAs (the only pass over the long rows besides the last one)
//   k_coef     the recurrence x <- K x - Lj_j (v_j . x) v_j on coefficient vectors, FP64, one CTA per instance,
//              same operation order as the reference; |v_j|^2 = c_j^T G c_j, Nj / Lj closed forms
//   k_combine  v_i = sum_k C[i][k] b_k for the recomputed rows, FP64 accumulation, many CTAs; also refreshes the
//              sequence-ordered mirror k_sample streams from
// All cancellation (|pc| ~ 10^3 collapsing to |v| ~ 10 once the paths are collinear) happens in FP64 here, so this
// path is also the more accurate one.
#pragma once
#include "lmcma_common.cuh"

namespace lmcma {

constexpr int GRAM_TILE = 16;
constexpr int GRAM_KC = 64;        // columns staged per step

__device__ __forceinline__ const float* basis_row(const OptDev& o, int b, int i, int first_stale) {
    const int slot = o.t[(size_t)b * o.m + i];
    return (i < first_stale ? o.V : o.P) + ((size_t)b * o.m + slot) * o.ns;
}

// grid = (tiles_a, tiles_b, B) with tiles_b >= tiles_a used (upper triangle incl. diagonal), 256 threads = 16 x 16 dots
__global__ void __launch_bounds__(256) k_gram(OptDev o) {
    __shared__ float As[GRAM_TILE][GRAM_KC + 1];
    __shared__ float Bs[GRAM_TILE][GRAM_KC + 1];
    const int b = blockIdx.z, ta = blockIdx.x, tb = blockIdx.y;
    if (tb < ta) return;
    const int2 hdr = o.gram_hdr[b];
    const int first_stale = hdr.x, live = hdr.y;
    const int a0 = ta * GRAM_TILE, b0 = tb * GRAM_TILE;
    if (a0 >= live || b0 >= live) return;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;       // G[a0 + ty][b0 + tx]
    const int lr = threadIdx.x >> 4, lc = (threadIdx.x & 15) * 4; // loader: 16 rows x 16 float4
    const float* arow = (a0 + lr < live) ? basis_row(o, b, a0 + lr, first_stale) : nullptr;
    const float* brow = (b0 + lr < live) ? basis_row(o, b, b0 + lr, first_stale) : nullptr;
    double acc0 = 0.0, acc1 = 0.0;
    for (int c0 = 0; c0 < o.ns; c0 += GRAM_KC) {
        float4 av = make_float4(0.f, 0.f, 0.f, 0.f), bv = av;
        if (arow && c0 + lc < o.ns) av = *reinterpret_cast<const float4*>(arow + c0 + lc);
        if (brow && c0 + lc < o.ns) bv = *reinterpret_cast<const float4*>(brow + c0 + lc);
        __syncthreads();
        As[lr][lc] = av.x; As[lr][lc + 1] = av.y; As[lr][lc + 2] = av.z; As[lr][lc + 3] = av.w;
        Bs[lr][lc] = bv.x; Bs[lr][lc + 1] = bv.y; Bs[lr][lc + 2] = bv.z; Bs[lr][lc + 3] = bv.w;
        __syncthreads();
#pragma unroll 8
        for (int c = 0; c < GRAM_KC; c += 2) {
            acc0 = fma((double)As[ty][c], (double)Bs[tx][c], acc0);
            acc1 = fma((double)As[ty][c + 1], (double)Bs[tx][c + 1], acc1);
        }
    }
    const int ia = a0 + ty, ib = b0 + tx;
    if (ia < live && ib < live) {
        double* G = o.G + (size_t)b * o.m * o.m;
        const double g = acc0 + acc1;
        G[(size_t)ia * o.m + ib] = g;
        G[(size_t)ib * o.m + ia] = g;
    }
}

// one CTA per instance; dynamic shared memory: G, C, W as L x LP doubles (LP odd: conflict-free column access)
__global__ void __launch_bounds__(1024) k_coef(OptDev o) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int b = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int m = o.m;
    const int2 hdr = o.gram_hdr[b];
    const int first_stale = hdr.x, L = hdr.y;
    const int LP = m | 1;
    double* Gs = reinterpret_cast<double*>(smem_raw);             // [L][LP]
    double* Cs = Gs + (size_t)m * LP;                             // [L][LP]: row i = coefficients of the current x_i
    double* Ws = Cs + (size_t)m * LP;                             // [L][LP]: row j = G coef(v_j)
    double* lj_s = Ws + (size_t)m * LP;                           // [L]
    double* nv_s = lj_s + m;                                      // [L]
    const double* G = o.G + (size_t)b * m * m;
    const int* order = o.t + (size_t)b * m;
    for (int e = tid; e < L * L; e += nthr) {
        const int i = e / L, k = e - i * L;
        Gs[i * LP + k] = G[(size_t)i * m + k];
        Cs[i * LP + k] = (i == k) ? 1.0 : 0.0;
    }
    for (int j = tid; j < L; j += nthr) lj_s[j] = (j < first_stale) ? o.Lj[(size_t)b * m + order[j]] : 0.0;
    __syncthreads();
    const double Kd = o.K, r = o.c1 / (1.0 - o.c1), am = o.M;
    // 8 lanes per row (aligned groups inside a warp): the length-L dot products and updates are split 8 ways and folded
    // with three shuffles, so that a step of the recurrence is ~L/8 dependent FP64 FMAs instead of L
    const int sub = tid & 7, grp = tid >> 3, ngrp = nthr >> 3;
    auto quad_sum = [&](double v) {
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        return v;
    };
    auto finalize = [&](int j) {       // row j has had all its factors applied: W_j = G c_j, |v_j|^2, Lj_j (block-collective)
        const double* cj = Cs + j * LP;
        for (int k0 = 0; k0 < L; k0 += ngrp) {
            const int k = k0 + grp;
            double s0 = 0.0;
            if (k < L) { const double* gk = Gs + k * LP; for (int l = sub; l <= j; l += 8) s0 = fma(gk[l], cj[l], s0); }
            s0 = quad_sum(s0);
            if (k < L && sub == 0) Ws[j * LP + k] = s0;
        }
        __syncthreads();
        if (warp == 0) {
            double s0 = 0.0;
            for (int k = lane; k <= j; k += 32) s0 = fma(cj[k], Ws[j * LP + k], s0);
#pragma unroll
            for (int ofs = 16; ofs > 0; ofs >>= 1) s0 += __shfl_xor_sync(0xffffffffu, s0, ofs);
            if (lane == 0) {
                const double nv = s0 > 0.0 ? s0 : 0.0;
                const double t = sqrt(1.0 + r * nv);
                nv_s[j] = nv;
                lj_s[j] = r / (am * t * (t + 1.0));               // cancellation-free form of lmcma.cpp:388-389
            }
        }
        __syncthreads();
    };
    for (int e = tid; e < first_stale * L; e += nthr) {             // final rows: coef(v_j) = e_j, W_j = row j of G
        const int j = e / L, k = e - j * L;
        Ws[j * LP + k] = Gs[j * LP + k];
    }
    __syncthreads();
    if (first_stale == 0) finalize(0);                              // v_0 = pc_0
    for (int j = 0; j + 1 < L; ++j) {
        // factor j on every pending row i > j: x <- K x - Lj_j (v_j . x) v_j (lmcma.cpp:455-461), 8 lanes per row
        const double lj = lj_s[j];
        const bool fresh = j >= first_stale;
        const double* wj = Ws + j * LP;
        const double* cj = Cs + j * LP;
        const int ibeg = max(j + 1, first_stale);
        for (int i0 = ibeg; i0 < L; i0 += ngrp) {
            const int i = i0 + grp;
            const bool on = i < L;
            double* ci = Cs + (on ? i : 0) * LP;
            double d = 0.0;                                          // support of c_i: {0 .. j} and i
            if (on) {
                for (int k = sub; k <= j; k += 8) d = fma(ci[k], wj[k], d);
                if (sub == 0) d = fma(ci[i], wj[i], d);
            }
            const double e = lj * quad_sum(d);
            if (on) {
                if (fresh) { for (int k = sub; k <= j; k += 8) ci[k] = Kd * ci[k] - e * cj[k]; }
                else { for (int k = sub; k < j; k += 8) ci[k] *= Kd; if (sub == 0) ci[j] = Kd * ci[j] - e; }
                if (sub == 1) ci[i] *= Kd;
            }
        }
        __syncthreads();
        if (j + 1 >= first_stale) finalize(j + 1);                   // row j + 1 is final now
    }
    // ---- outputs: coefficients of the recomputed rows, Nj / Lj (lmcma.cpp:386-389) by slot and by position ----
    double* Cf = o.Cf + (size_t)b * m * m;
    for (int e = tid; e < L * L; e += nthr) { const int i = e / L, k = e - i * L; if (i >= first_stale) Cf[(size_t)i * m + k] = (k <= i) ? Cs[i * LP + k] : 0.0; }
    for (int i = first_stale + tid; i < L; i += nthr) {
        const int slot = order[i];
        const double nv = nv_s[i], t = sqrt(1.0 + r * nv);
        const double nj = am * r / (t + 1.0);
        o.Nj[(size_t)b * m + slot] = nj; o.Lj[(size_t)b * m + slot] = lj_s[i];
        o.Njf[(size_t)b * m + slot] = (float)nj; o.Njs[(size_t)b * m + i] = (float)nj;
    }
}

// v_i = sum_{k <= i} C[i][k] b_k for the recomputed rows: grid = (ceil(nq / 128), ceil(m / 8), B), 128 threads,
// each thread one float4 column of 8 rows; writes V (slot-indexed) and both halves of the mirror
__global__ void __launch_bounds__(128) k_combine(OptDev o) {
    __shared__ double cs[8][128];                                   // coefficients of this CTA's 8 rows (m <= 128)
    __shared__ const float* rowp[128];                              // basis row pointers (one dependent load chain, not L)
    const int b = blockIdx.z, nq = o.ns >> 2, q = blockIdx.x * 128 + threadIdx.x;
    const int2 hdr = o.gram_hdr[b];
    const int first_stale = hdr.x, L = hdr.y;
    const int i0 = first_stale + blockIdx.y * 8;
    if (i0 >= L) return;
    const int rows = min(8, L - i0), kmax = i0 + rows - 1;
    const double* Cf = o.Cf + (size_t)b * o.m * o.m;
    for (int e = threadIdx.x; e < 8 * 128; e += 128) { const int r2 = e >> 7, k = e & 127; cs[r2][k] = (r2 < rows && k <= i0 + r2) ? Cf[(size_t)(i0 + r2) * o.m + k] : 0.0; }
    if (threadIdx.x <= kmax) rowp[threadIdx.x] = basis_row(o, b, threadIdx.x, first_stale);
    __syncthreads();
    if (q >= nq) return;
    double acc[8][4];
#pragma unroll
    for (int r2 = 0; r2 < 8; ++r2) { acc[r2][0] = acc[r2][1] = acc[r2][2] = acc[r2][3] = 0.0; }
    for (int k0 = 0; k0 <= kmax; k0 += 4) {                         // 4 independent row loads in flight
        float4 bk[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) bk[u] = (k0 + u <= kmax) ? __ldg(reinterpret_cast<const float4*>(rowp[k0 + u]) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const double x = bk[u].x, y = bk[u].y, z = bk[u].z, w = bk[u].w;
            const int k = min(k0 + u, 127);
#pragma unroll
            for (int r2 = 0; r2 < 8; ++r2) {
                const double c = cs[r2][k];                         // 0 beyond a row's support
                acc[r2][0] = fma(c, x, acc[r2][0]); acc[r2][1] = fma(c, y, acc[r2][1]);
                acc[r2][2] = fma(c, z, acc[r2][2]); acc[r2][3] = fma(c, w, acc[r2][3]);
            }
        }
    }
    for (int r2 = 0; r2 < rows; ++r2) {
        const int i = i0 + r2, slot = o.t[(size_t)b * o.m + i];
        const float4 v = make_float4((float)acc[r2][0], (float)acc[r2][1], (float)acc[r2][2], (float)acc[r2][3]);
        // V[slot] is not a basis row of any CTA (pending rows are read from P), so it can be written in place
        reinterpret_cast<float4*>(o.V + ((size_t)b * o.m + slot) * o.ns)[q] = v;
        float4* mir = reinterpret_cast<float4*>(o.VPs + ((size_t)b * o.m + i) * 2 * o.ns);
        mir[q] = v;
        mir[nq + q] = reinterpret_cast<const float4*>(o.P + ((size_t)b * o.m + slot) * o.ns)[q];
    }
}

}  // namespace lmcma
