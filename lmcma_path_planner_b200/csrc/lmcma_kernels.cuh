// lmcma_kernels.cuh — hand-written sm_100a kernels of the LM-CMA trajectory-optimisation hot path.
//
// One generation = k_cost -> k_rank -> k_recombine -> k_update -> k_sample, all on one stream and
// replayed from a CUDA graph (lmcma_capi.cu).  Everything is batched over B independent optimiser
// instances (gridDim.y or gridDim.x = B).  FP32 on CUDA cores for the bulk data, FP64 for the
// handful of per-instance scalars whose closed forms cancel (sigma, s, Nj, Lj, xmean).
//
// Reference lines each kernel stands in for are cited at the kernel.  Row stride `ns` is n rounded
// up to a multiple of 4 floats so that every row is 16-byte aligned (float4 / bulk-copy granularity);
// the padding lanes are kept at exactly 0.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "lmcma_layout.hpp"

namespace lmcma {

constexpr int KMAX_SUBSTEPS = 65536;   // cap on sub-steps per segment (cost model)

struct Scalars {          // per-instance scalar state
    double sigma;         // LMCMA::sigma
    double s;             // LMCMA::s
    double best_f;        // CMABase::BestF
    long long counteval;  // CMABase::counteval
    int itr;              // CMABase::itr
    int live;             // LMCMA::iterator_sz
    int best_local;       // 1 if best_x holds the row that achieved best_f (split mode: may live on a peer)
    int pad;
};

struct OptDev {
    int n, ns, lambda, mu, m, B;
    int pop_offset, pop_count;     // rows of lambda owned by this handle
    int rng_mode, record_z;
    unsigned long long seed;
    // population
    float* X;          // B x pop_count x ns
    float* Z;          // B x pop_count x ns (INJECT / record_z) or null
    float* fit;        // B x lambda   fitness as evaluated / told (all rows, global order)
    float* fit_sorted; // B x lambda
    float* prev_fit;   // B x lambda   previous generation (any order)
    int* rank;         // B x lambda   (only [pop_offset, +pop_count) written in split mode)
    int* arindex;      // B x lambda
    int* ncoll;        // B x pop_count
    int* nsamp;        // B x pop_count
    // distribution state
    double* xmean;     // B x ns
    float* pc;         // B x ns
    float* V;          // B x m x ns   (slot-indexed)
    float* P;          // B x m x ns
    double* Nj;        // B x m
    double* Lj;        // B x m
    float* Njf;        // B x m   FP32 copies read by k_sample
    int* t;            // B x m   slot order, oldest -> newest
    int* vec;          // B x m   generation stamp per slot
    Scalars* sc;       // B
    float* best_x;     // B x ns
    const float* lo;   // n or null
    const float* hi;   // n or null
    const float* w;    // mu recombination weights
    float* partial;    // B x RS x ns weighted partial sums of (x - xmean)
    int RS;
    unsigned long long* S_count;   // B : #{(i,j): prev_j < cur_i}
    double c1, cc, cs, target, K, M, mueff;
    double pc_coef;    // sqrt(cc (2 - cc) mueff), lmcma.cpp:328
};

struct MapDev {
    int dims, nx, ny, nz;
    unsigned nbx, nby;         // bricks per row / per column (lmcma_layout.hpp)
    int storage;               // 0 = F32 sign-tagged reciprocal clearance, 1 = U8 quantised distance
    const float* g32;
    const unsigned char* q8;
    const float* lut;          // 256 sign-tagged reciprocal clearances (U8)
    float g_coll;              // 1 / c_min
};

struct CostArgs {
    int W;
    float w_len, w_clr, w_col;
    const float* X;            // candidates
    long long ld;              // row stride (floats)
    long long inst_rows;       // rows per instance (gridDim.x)
    const float* ends;         // per instance: start[3], goal[3]   (stride 6 floats)
    int ends_per_instance;     // 1: ends[b], 0: ends[0] for every instance
    float* f;                  // outputs, indexed [b * f_stride + f_offset + row]
    long long f_stride;
    int f_offset;
    int* ncoll;                // indexed [b * inst_rows + row] (nullable)
    int* nsamp;
    long long* cells;          // trace mode (nullable)
    long long max_cells;
};

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float canon_fitness(float f) { return (f != f) ? __int_as_float(0x7f800000) : f; }

// mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP) wrappers
__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(void* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(void* bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(void* bar, unsigned parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, void* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Philox4x32-10 (Salmon et al. 2011): counter-based, so any rank can regenerate any offspring row.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}
__device__ __forceinline__ float4 philox_normal4(unsigned q, unsigned row, unsigned gen, unsigned inst, unsigned long long seed) {
    const uint4 r = philox4x32_10(make_uint4(q, row, gen, inst), make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
    const float s = 2.3283064365386963e-10f;   // 2^-32
    const float u0 = r.x * s + 1.1641532182693481e-10f, u1 = r.y * s + 1.1641532182693481e-10f;
    const float u2 = r.z * s + 1.1641532182693481e-10f, u3 = r.w * s + 1.1641532182693481e-10f;
    const float ra = sqrtf(-2.0f * logf(u0)), rb = sqrtf(-2.0f * logf(u2));
    float s0, c0, s1, c1;
    sincospif(2.0f * u1, &s0, &c0);
    sincospif(2.0f * u3, &s1, &c1);
    return make_float4(ra * c0, ra * s0, rb * c1, rb * s1);
}

// ------------------------------------------------------------------------------------------------
// k_sample — LMCMA::sample + computeAz + applyBoundaries (lmcma.cpp:301-311, 431-447, 220-230) for
// all offspring at once.  One CTA per tile of (blockDim/32)*RB offspring rows of one instance; the
// live (v_j, pc_j) pairs are streamed through shared memory in sequence order by 1-D bulk async
// copies (double-buffered, mbarrier-tracked); each warp keeps RB rows of z and Az in registers,
// lanes own float4 column slots (lane + 32*i), dots are warp-shuffle reductions against the ORIGINAL
// z (lmcma.cpp:441-443) and the M*Az + d*pc_j recurrence runs in the reference's order.
// ------------------------------------------------------------------------------------------------
template <int NV, int RB, int MAXT>
__global__ void __launch_bounds__(MAXT) k_sample(OptDev o, int kc /* pairs per stage */) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int ns = o.ns, nq = ns >> 2;
    const Scalars sc = o.sc[b];
    const int live = sc.live;
    const int* order = o.t + (size_t)b * o.m;
    float* stage_base = reinterpret_cast<float*>(smem_raw);
    const size_t stage_floats = (size_t)kc * 2 * ns;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(stage_base + 2 * stage_floats);
    float* nj_s = reinterpret_cast<float*>(bars + 2);             // Nj of the live pairs, in sequence order
    const int nchunks = (live + kc - 1) / kc;

    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_barrier_init();
    }
    for (int k = threadIdx.x; k < live; k += blockDim.x) nj_s[k] = o.Njf[(size_t)b * o.m + order[k]];
    __syncthreads();
    auto issue = [&](int chunk) {   // thread 0 only
        const int st = chunk & 1;
        const int k0 = chunk * kc, cnt = min(kc, live - k0);
        float* dst = stage_base + st * stage_floats;
        mbar_expect_tx(&bars[st], (unsigned)(cnt * 2 * ns * sizeof(float)));
        for (int k = 0; k < cnt; ++k) {
            const int slot = order[k0 + k];
            bulk_g2s(dst + (size_t)(2 * k) * ns, o.V + ((size_t)b * o.m + slot) * ns, ns * sizeof(float), &bars[st]);
            bulk_g2s(dst + (size_t)(2 * k + 1) * ns, o.P + ((size_t)b * o.m + slot) * ns, ns * sizeof(float), &bars[st]);
        }
    };
    if (threadIdx.x == 0) {
        if (nchunks > 0) issue(0);
        if (nchunks > 1) issue(1);
    }

    // ---- load / generate z for this warp's RB rows ----
    const int row0 = (blockIdx.x * nwarps + warp) * RB;       // local row index
    float4 z[RB][NV], az[RB][NV];
#pragma unroll
    for (int r = 0; r < RB; ++r) {
        const int row = row0 + r;
        const bool rv = row < o.pop_count;
        const size_t roff = ((size_t)b * o.pop_count + (rv ? row : 0)) * ns;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int q = lane + 32 * i;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (rv && q < nq) {
                if (o.rng_mode == 0) {
                    v = philox_normal4((unsigned)q, (unsigned)(o.pop_offset + row), (unsigned)sc.itr, (unsigned)b, o.seed);
                    const int e = q * 4;
                    if (e + 1 >= o.n) v.y = 0.f;
                    if (e + 2 >= o.n) v.z = 0.f;
                    if (e + 3 >= o.n) v.w = 0.f;
                    if (o.Z) reinterpret_cast<float4*>(o.Z + roff)[q] = v;
                } else {
                    v = reinterpret_cast<const float4*>(o.Z + roff)[q];
                }
            }
            z[r][i] = v;
            az[r][i] = v;
        }
    }

    // ---- stream the pairs, four at a time: their dots are mutually independent (all against the original z),
    //      so the 4*RB warp reductions overlap; only the M*Az + d*pc_j recurrence is ordered ----
    const float Mf = (float)o.M;
    for (int c = 0; c < nchunks; ++c) {
        const int st = c & 1;
        mbar_wait(&bars[st], (unsigned)((c >> 1) & 1));
        const float* sb = stage_base + st * stage_floats;
        const int k0 = c * kc, cnt = min(kc, live - k0);
        for (int k = 0; k < cnt; k += 4) {
            const int gcnt = min(4, cnt - k);
            float d[4][RB];
#pragma unroll
            for (int g = 0; g < 4; ++g)
#pragma unroll
                for (int r = 0; r < RB; ++r) d[g][r] = 0.f;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int q = lane + 32 * i;
                if (q < nq) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        if (g < gcnt) {
                            const float4 v = reinterpret_cast<const float4*>(sb + (size_t)(2 * (k + g)) * ns)[q];
#pragma unroll
                            for (int r = 0; r < RB; ++r)
                                d[g][r] += fmaf(v.x, z[r][i].x, v.y * z[r][i].y) + fmaf(v.z, z[r][i].z, v.w * z[r][i].w);
                        }
                    }
                }
            }
#pragma unroll
            for (int ofs = 16; ofs > 0; ofs >>= 1)
#pragma unroll
                for (int g = 0; g < 4; ++g)
#pragma unroll
                    for (int r = 0; r < RB; ++r) d[g][r] += __shfl_xor_sync(0xffffffffu, d[g][r], ofs);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const float nj = (g < gcnt) ? nj_s[k0 + k + g] : 0.f;
#pragma unroll
                for (int r = 0; r < RB; ++r) d[g][r] *= nj;
            }
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int q = lane + 32 * i;
                if (q < nq) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        if (g < gcnt) {
                            const float4 p = reinterpret_cast<const float4*>(sb + (size_t)(2 * (k + g) + 1) * ns)[q];
#pragma unroll
                            for (int r = 0; r < RB; ++r) {
                                az[r][i].x = fmaf(Mf, az[r][i].x, d[g][r] * p.x);
                                az[r][i].y = fmaf(Mf, az[r][i].y, d[g][r] * p.y);
                                az[r][i].z = fmaf(Mf, az[r][i].z, d[g][r] * p.z);
                                az[r][i].w = fmaf(Mf, az[r][i].w, d[g][r] * p.w);
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();                      // every warp is done with stage st
        if (threadIdx.x == 0 && c + 2 < nchunks) issue(c + 2);
    }

    // ---- x = xmean + sigma * Az, clamp lo then hi (lmcma.cpp:307-310, 222-229) ----
    const double* xm = o.xmean + (size_t)b * ns;
#pragma unroll
    for (int r = 0; r < RB; ++r) {
        const int row = row0 + r;
        if (row >= o.pop_count) continue;
        float* xrow = o.X + ((size_t)b * o.pop_count + row) * ns;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int q = lane + 32 * i;
            if (q >= nq) continue;
            const double2 m01 = reinterpret_cast<const double2*>(xm)[2 * q];
            const double2 m23 = reinterpret_cast<const double2*>(xm)[2 * q + 1];
            float4 x;
            x.x = (float)(m01.x + sc.sigma * (double)az[r][i].x);
            x.y = (float)(m01.y + sc.sigma * (double)az[r][i].y);
            x.z = (float)(m23.x + sc.sigma * (double)az[r][i].z);
            x.w = (float)(m23.y + sc.sigma * (double)az[r][i].w);
            const int e = q * 4;
            if (o.lo) {
                if (e < o.n) x.x = fmaxf(x.x, o.lo[e]);
                if (e + 1 < o.n) x.y = fmaxf(x.y, o.lo[e + 1]);
                if (e + 2 < o.n) x.z = fmaxf(x.z, o.lo[e + 2]);
                if (e + 3 < o.n) x.w = fmaxf(x.w, o.lo[e + 3]);
            }
            if (o.hi) {
                if (e < o.n) x.x = fminf(x.x, o.hi[e]);
                if (e + 1 < o.n) x.y = fminf(x.y, o.hi[e + 1]);
                if (e + 2 < o.n) x.z = fminf(x.z, o.hi[e + 2]);
                if (e + 3 < o.n) x.w = fminf(x.w, o.hi[e + 3]);
            }
            if (e + 1 >= o.n) x.y = 0.f;
            if (e + 2 >= o.n) x.z = 0.f;
            if (e + 3 >= o.n) x.w = 0.f;
            reinterpret_cast<float4*>(xrow)[q] = x;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_cost — batched trajectory cost (DESIGN.md "cost model"; the reference's per-state pieces are
// ValidityChecker::isValid/clearance planner.cpp:591-631, ClearanceObjective::stateCost :655-669,
// weights :677-690).  One CTA per trajectory.  Phase 1: per-segment records {A, B-A, 1/K, len/K} and
// sub-step counts into shared memory + block scan.  Phase 2: the flattened sample sequence is cut into
// 32-sample blocks, a contiguous run of blocks per warp, consecutive samples on consecutive lanes: they
// are <= 1 cell apart, and the map is stored in 128-byte bricks (lmcma_layout.hpp), so one warp load
// touches a handful of lines instead of 32 (the L1 wavefront rate, not DRAM, bounds a row-major
// gather).  Phase 3: block reduction.
// The index path (t = k * (1/K); q = A + t*d; rint; bounds test) uses explicitly rounded FP32
// mul/add so that no FMA contraction can change a cell index relative to the CPU oracle.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int substeps_of(float linf) {
    if (linf >= 1.0f) return linf <= (float)KMAX_SUBSTEPS ? (int)ceilf(linf) : KMAX_SUBSTEPS;
    return 1;   // also NaN
}

template <int DIMS, int STORAGE, bool TRACE>
__global__ void __launch_bounds__(256) k_cost(MapDev mp, CostArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int W = a.W, NSEG = W + 1;
    float4* segA = reinterpret_cast<float4*>(smem_raw);          // 2-D {Ax, Ay, dx, dy}   3-D {Ax, Ay, Az, dx}
    float4* segB = segA + NSEG;                                  // 2-D {invK, scale, -, -} 3-D {dy, dz, invK, scale}
    int* off = reinterpret_cast<int*>(segB + NSEG);              // NSEG + 1 exclusive sample offsets
    float* lut = reinterpret_cast<float*>(off + NSEG + 1);       // 256 (U8 only)
    __shared__ float red_f[2][8];
    __shared__ int red_i[8];
    __shared__ int warp_tot[8];

    const int row = blockIdx.x, b = blockIdx.y;
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
    const float* x = a.X + ((size_t)b * a.inst_rows + row) * a.ld;
    const float* en = a.ends + (a.ends_per_instance ? (size_t)b * 6 : 0);
    if (STORAGE == 1) for (int i = tid; i < 256; i += nthr) lut[i] = mp.lut[i];

    // ---- phase 1: per-segment records, sub-step counts, lengths; block scan of the sample counts ----
    const int spt = (NSEG + nthr - 1) / nthr;                    // consecutive segments per thread
    const int s_begin = min(NSEG, tid * spt), s_end = min(NSEG, s_begin + spt);
    float len_acc = 0.f; int my_cnt = 0;
    for (int s = s_begin; s < s_end; ++s) {
        float A[3], D[3]; float linf = 0.f, l2 = 0.f; bool bad = false;
#pragma unroll
        for (int c = 0; c < DIMS; ++c) {
            A[c] = (s == 0) ? en[c] : x[c * W + s - 1];
            const float Bc = (s == W) ? en[3 + c] : x[c * W + s];
            D[c] = __fsub_rn(Bc, A[c]);
            const float ad = fabsf(D[c]);
            bad |= (ad != ad);
            if (ad > linf) linf = ad;
            l2 = fmaf(D[c], D[c], l2);
        }
        if (bad) linf = __int_as_float(0x7fc00000);
        const int K = substeps_of(linf);
        const float invK = __frcp_rn((float)K);                  // == 1.0f / (float)K, IEEE round-to-nearest
        const float len = sqrtf(l2);
        len_acc += len;
        // A segment with a non-finite end point has no sample inside the map (NaN / inf coordinates fail the
        // oracle's bounds test for every k): encode that as a finite far-away anchor so that the sample loop
        // can use integer conversion + unsigned bounds tests (NaN would convert to 0).
        bool finite = true;
#pragma unroll
        for (int c = 0; c < DIMS; ++c) finite = finite && (fabsf(A[c]) < 3.0e38f) && (fabsf(D[c]) < 3.0e38f);
        if (!finite) {
#pragma unroll
            for (int c = 0; c < DIMS; ++c) { A[c] = -1.0e9f; D[c] = 0.f; }
        }
        if (DIMS == 2) { segA[s] = make_float4(A[0], A[1], D[0], D[1]); segB[s] = make_float4(invK, len * invK, 0.f, 0.f); }
        else { segA[s] = make_float4(A[0], A[1], A[2], D[0]); segB[s] = make_float4(D[1], D[2], invK, len * invK); }
        off[s + 1] = K + 1;                                      // samples of this segment (rewritten below)
        my_cnt += K + 1;
    }
    int incl = my_cnt;                                           // block-wide exclusive scan of my_cnt
#pragma unroll
    for (int o2 = 1; o2 < 32; o2 <<= 1) {
        const int nb = __shfl_up_sync(0xffffffffu, incl, o2);
        if (lane >= o2) incl += nb;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int base = incl - my_cnt, T = 0;
    for (int w2 = 0; w2 < nwarps; ++w2) { const int v = warp_tot[w2]; if (w2 < warp) base += v; T += v; }
    if (tid == 0) off[0] = 0;
    for (int s = s_begin; s < s_end; ++s) { base += off[s + 1]; off[s + 1] = base; }
    __syncthreads();

    // ---- phase 2: consecutive samples on consecutive lanes (<= 1 cell apart -> few lines per warp load) ----
    const int nblk = (T + 31) >> 5;
    const int blk0 = (int)(((long long)warp * nblk) / nwarps), blk1 = (int)(((long long)(warp + 1) * nblk) / nwarps);
    float clr_acc = 0.f; int coll = 0;
    if (blk0 < blk1) {
        int lo = 0, hi = NSEG - 1;                               // warp-uniform: last s with off[s] <= first sample
        const int tfirst = blk0 << 5;
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (off[mid] <= tfirst) lo = mid; else hi = mid - 1; }
        int s_warp = lo;
        const unsigned nxm1 = (unsigned)(mp.nx - 1), nym1 = (unsigned)(mp.ny - 1), nzm1 = (unsigned)(mp.nz - 1);
        const int last = NSEG - 1;
        for (int blk = blk0; blk < blk1; ++blk) {
            const int t = (blk << 5) + lane;
            const bool valid = t < T;
            const int tt = valid ? t : T - 1;
            int s = s_warp;
            int cur = off[s], nxt = off[s + 1];                  // warp-uniform (broadcast) loads
            while (tt >= nxt) { ++s; cur = nxt; nxt = off[s + 1]; }
            s_warp = __shfl_sync(0xffffffffu, s, 31);
            const int K = nxt - cur - 1, k = tt - cur;
            const float4 ra = segA[s];
            float invK, scale, dx, dy, dz = 0.f, az = 0.f;
            if (DIMS == 2) { const float2 rb = *reinterpret_cast<const float2*>(&segB[s]); invK = rb.x; scale = rb.y; dx = ra.z; dy = ra.w; }
            else { const float4 rb = segB[s]; dx = ra.w; dy = rb.x; dz = rb.y; invK = rb.z; scale = rb.w; az = ra.z; }
            const float tk = __fmul_rn((float)k, invK);
            // round-half-even conversion == (int)rintf(q); saturates for huge |q| (-> fails the unsigned test)
            const int ix = __float2int_rn(__fadd_rn(ra.x, __fmul_rn(tk, dx)));
            const int iy = __float2int_rn(__fadd_rn(ra.y, __fmul_rn(tk, dy)));
            bool inb = ((unsigned)ix <= nxm1) && ((unsigned)iy <= nym1);
            int iz = 0;
            if (DIMS == 3) {
                iz = __float2int_rn(__fadd_rn(az, __fmul_rn(tk, dz)));
                inb = inb && ((unsigned)iz <= nzm1);
            }
            const unsigned adr = inb ? brick_offset<DIMS, STORAGE>((unsigned)ix, (unsigned)iy, (unsigned)iz, mp.nbx, mp.nby) : 0u;
            float g = (STORAGE == 0) ? __ldg(mp.g32 + adr) : lut[__ldg(mp.q8 + adr)];   // branch-free: cell 0 when outside
            g = inb ? g : -mp.g_coll;
            float wgt = (k == 0 || k == K) ? 0.5f : 1.0f;
            wgt = valid ? wgt : 0.f;
            clr_acc = fmaf(fabsf(g) * wgt, scale, clr_acc);
            coll += (valid && g < 0.f && (k < K || s == last)) ? 1 : 0;
            if (TRACE) {
                if (valid && t < a.max_cells) a.cells[t] = inb ? ((long long)iz * mp.ny + iy) * mp.nx + ix : -1;
            }
        }
    }

    // ---- phase 3: block reduction ----
    len_acc = warp_sum(len_acc);
    clr_acc = warp_sum(clr_acc);
    coll = warp_sum_i(coll);
    if (lane == 0) { red_f[0][warp] = len_acc; red_f[1][warp] = clr_acc; red_i[warp] = coll; }
    __syncthreads();
    if (tid == 0) {
        float L = 0.f, C = 0.f; int NC = 0;
        for (int w2 = 0; w2 < nwarps; ++w2) { L += red_f[0][w2]; C += red_f[1][w2]; NC += red_i[w2]; }
        const float f = fmaf(a.w_col, (float)NC, fmaf(a.w_clr, C, a.w_len * L));
        a.f[(size_t)b * a.f_stride + a.f_offset + row] = f;
        if (a.ncoll) a.ncoll[(size_t)b * a.inst_rows + row] = NC;
        if (a.nsamp) a.nsamp[(size_t)b * a.inst_rows + row] = T;
    }
}

// ------------------------------------------------------------------------------------------------
// k_rank — myqsort/compare (lmcma.cpp:84-104, 315) and the 2*lambda merged ranking of the step-size
// rule (lmcma.cpp:393-411) as rank-by-counting: rank_i = #{j: f_j < f_i or (f_j == f_i and j < i)}
// reproduces the stable ascending order (ties keep the lower id, -0 == +0); the merged ranking only
// enters through S = #{(i,j): prev_j < cur_i} (see k_update).  NaN fitness ranks as +inf (the
// reference's comparator is undefined for NaN).  grid = (ceil(pop_count/32), B): one candidate per lane,
// the 8 warps of a CTA split the comparison range.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_rank(OptDev o, const float* __restrict__ f_all) {
    __shared__ int part_c[8][32];
    __shared__ unsigned long long part_p[8][32];
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int il = blockIdx.x * 32 + lane;                 // local row: one candidate per lane
    const int i = o.pop_offset + il;                       // global candidate id
    const bool valid = il < o.pop_count;
    const float* cur = f_all + (size_t)b * o.lambda;
    const float* prev = o.prev_fit + (size_t)b * o.lambda;
    const float ki = valid ? canon_fitness(cur[i]) : 0.f;
    // the 8 warps split the j range; every lane of a warp reads the same f_j (broadcast loads)
    const int j0 = (int)(((long long)warp * o.lambda) / 8), j1 = (int)(((long long)(warp + 1) * o.lambda) / 8);
    int c_lt = 0; unsigned long long p_lt = 0;
    int j = j0;
    for (; j + 4 <= j1; j += 4) {
        float kj[4], pj[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { kj[u] = canon_fitness(cur[j + u]); pj[u] = prev[j + u]; }
        int pl = 0;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            c_lt += (kj[u] < ki) || (kj[u] == ki && (j + u) < i);
            pl += pj[u] < ki;
        }
        p_lt += pl;
    }
    for (; j < j1; ++j) {
        const float kj = canon_fitness(cur[j]);
        c_lt += (kj < ki) || (kj == ki && j < i);
        p_lt += prev[j] < ki;
    }
    part_c[warp][lane] = c_lt;
    part_p[warp][lane] = p_lt;
    __syncthreads();
    if (warp == 0) {
        int c = 0; unsigned long long p = 0;
#pragma unroll
        for (int w2 = 0; w2 < 8; ++w2) { c += part_c[w2][lane]; p += part_p[w2][lane]; }
        if (!valid) p = 0;
        if (valid) {
            o.rank[(size_t)b * o.lambda + i] = c;
            o.arindex[(size_t)b * o.lambda + c] = i;
            o.fit_sorted[(size_t)b * o.lambda + c] = ki;
        }
        // warp-reduce p (<= 32 * lambda) -> one integer atomic per CTA (order-independent, deterministic)
        unsigned lo32 = (unsigned)p, hi32 = (unsigned)(p >> 32);
#pragma unroll
        for (int ofs = 16; ofs > 0; ofs >>= 1) {
            const unsigned l2 = __shfl_xor_sync(0xffffffffu, lo32, ofs), h2 = __shfl_xor_sync(0xffffffffu, hi32, ofs);
            const unsigned long long a = (((unsigned long long)hi32 << 32) | lo32) + (((unsigned long long)h2 << 32) | l2);
            lo32 = (unsigned)a; hi32 = (unsigned)(a >> 32);
        }
        if (lane == 0) atomicAdd(o.S_count + b, ((unsigned long long)hi32 << 32) | lo32);
    }
}

// ------------------------------------------------------------------------------------------------
// k_recombine — the weighted recombination of LMCMA::update (lmcma.cpp:316-326) as partial sums of
// w_{rank(i)} * (x_i - xmean) over the rows this handle owns (candidate order, fixed split -> run-to-run
// deterministic).  Summing differences keeps the FP32 sum accurate relative to the mean SHIFT, which
// is what the evolution path needs (lmcma.cpp:327-329).  grid = (ceil(nq/128), RS, B), block 512 =
// 128 float4 columns x 4 row groups, 4 independent row loads in flight per thread.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) k_recombine(OptDev o) {
    __shared__ float4 red[3][128];
    const int b = blockIdx.z, rs = blockIdx.y;
    const int tq = threadIdx.x & 127, grp = threadIdx.x >> 7;     // 128 float4 columns x 4 row groups
    const int q = blockIdx.x * 128 + tq, nq = o.ns >> 2;
    const int rows_per = (o.pop_count + o.RS - 1) / o.RS;
    const int r0 = rs * rows_per, r1 = min(o.pop_count, r0 + rows_per);
    const int* rk = o.rank + (size_t)b * o.lambda + o.pop_offset;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q < nq) {
        const double* xm = o.xmean + (size_t)b * o.ns + 4 * q;
        const float4 m4 = make_float4((float)xm[0], (float)xm[1], (float)xm[2], (float)xm[3]);
        for (int rbase = r0 + grp; rbase < r1; rbase += 16) {     // 4 rows in flight per thread
            float w[4]; float4 x[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = rbase + 4 * u;
                const int rnk = (r < r1) ? rk[r] : o.mu;
                w[u] = (rnk < o.mu) ? o.w[rnk] : 0.f;
                x[u] = (rnk < o.mu) ? reinterpret_cast<const float4*>(o.X + ((size_t)b * o.pop_count + r) * o.ns)[q] : m4;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                acc.x = fmaf(w[u], x[u].x - m4.x, acc.x);
                acc.y = fmaf(w[u], x[u].y - m4.y, acc.y);
                acc.z = fmaf(w[u], x[u].z - m4.z, acc.z);
                acc.w = fmaf(w[u], x[u].w - m4.w, acc.w);
            }
        }
    }
    if (grp > 0) red[grp - 1][tq] = acc;
    __syncthreads();
    if (grp == 0 && q < nq) {
#pragma unroll
        for (int g = 0; g < 3; ++g) { const float4 t = red[g][tq]; acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w; }
        reinterpret_cast<float4*>(o.partial + ((size_t)b * o.RS + rs) * o.ns)[q] = acc;
    }
}

// split mode: fold the RS local partials and the local S count into the all-gather payload
__global__ void __launch_bounds__(128) k_pack_payload(OptDev o, float* __restrict__ payload) {
    const int b = blockIdx.y, e = blockIdx.x * 128 + threadIdx.x;
    float* pay = payload + (size_t)b * (o.ns + 4);
    if (e < o.ns) {
        float acc = 0.f;
        for (int rs = 0; rs < o.RS; ++rs) acc += o.partial[((size_t)b * o.RS + rs) * o.ns + e];
        pay[e] = acc;
    }
    if (e == 0) {
        const unsigned long long S = o.S_count[b];
        pay[o.ns] = __uint_as_float((unsigned)S);
        pay[o.ns + 1] = __uint_as_float((unsigned)(S >> 32));
        pay[o.ns + 2] = 0.f; pay[o.ns + 3] = 0.f;
        o.S_count[b] = 0ull;
    }
}

// ------------------------------------------------------------------------------------------------
// k_update — the rest of LMCMA::update (lmcma.cpp:316-424): mean + evolution path, slot bookkeeping,
// recomputation of the inverse-direction vectors (invAz, lmcma.cpp:449-463) and the population-success
// step size.  One CTA per instance.  The triangular recompute is run factor-major: step j applies
// factor j to every still-pending row i > j (same per-row operation order as the reference's row-major
// loops, lmcma.cpp:375-390), so its depth is `live` block barriers instead of live^2/2 serial dots.
// slices: n_slices partial-sum slices per instance; slice k of instance b starts at
//         slices + k * slice_stride + b * inst_stride.  S: see s_src.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_update(OptDev o, const float* __restrict__ slices, int n_slices,
                                                 long long slice_stride, long long inst_stride,
                                                 const float* __restrict__ f_all, int payload_mode, int cap_rows) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int m = o.m, ns = o.ns;
    float* rows_s = reinterpret_cast<float*>(smem_raw);           // cap_rows x ns: the pending rows live here
    float* lj_s = rows_s + (size_t)cap_rows * ns;                 // m: Lj in sequence order
    float* nv_s = lj_s + m;                                       // m: |v|^2 of the recomputed rows
    int* order = reinterpret_cast<int*>(nv_s + m);                // m: slot order
    int* stamp = order + m;                                       // m: generation stamp per SLOT
    __shared__ int sh_first_stale, sh_live, sh_slot_new;
    __shared__ double sh_sigma_old;
    const int b = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
    Scalars* scp = o.sc + b;
    int* tg = o.t + (size_t)b * m;
    int* vg = o.vec + (size_t)b * m;

    // ---- slot bookkeeping (lmcma.cpp:331-364): data-independent integer logic.  t[] / vec[] are staged in
    //      shared memory with one coalesced load so the serial part never waits on HBM ----
    for (int i = tid; i < m; i += nthr) { order[i] = tg[i]; stamp[i] = vg[i]; }
    __syncthreads();
    if (tid == 0) {
        const int itr = scp->itr;
        int first_stale = 1;
        if (itr < m) {
            order[itr] = itr;
        } else {
            int gap_min = stamp[order[1]] - stamp[order[0]];
            for (int j = 1; j < m - 1; ++j) {
                const int gap = stamp[order[j + 1]] - stamp[order[j]];
                if (gap < gap_min) { gap_min = gap; first_stale = j + 1; }
            }
            if (gap_min >= m /* maxsteps = nvectors, lmcma.cpp:267 */) first_stale = 0;
            if (first_stale != m - 1) {
                const int recycled = order[first_stale];
                for (int j = first_stale; j < m - 1; ++j) order[j] = order[j + 1];
                order[m - 1] = recycled;
            }
        }
        const int live = min(itr + 1, m);
        const int slot_new = order[live - 1];
        stamp[slot_new] = itr;
        if (first_stale == 1) first_stale = 0;                   // lmcma.cpp:373-374
        sh_first_stale = first_stale; sh_live = live; sh_slot_new = slot_new;
        sh_sigma_old = scp->sigma;
    }
    __syncthreads();
    const int first_stale = sh_first_stale, live = sh_live, slot_new = sh_slot_new;
    const double sigma_old = sh_sigma_old;
    double* Njd = o.Nj + (size_t)b * m;
    double* Ljd = o.Lj + (size_t)b * m;
    float* Njf = o.Njf + (size_t)b * m;
    for (int i = tid; i < m; i += nthr) {
        const int slot = order[i];
        tg[i] = slot; vg[i] = stamp[i];
        lj_s[i] = (i < live) ? (float)Ljd[slot] : 0.f;            // rows >= first_stale are rewritten below
    }

    // ---- mean, evolution path, new pc_j (lmcma.cpp:316-329, 365-366) ----
    {
        const double coef = o.pc_coef / sigma_old;               // sqrt(cc (2 - cc) mueff) / sigma
        double* xm = o.xmean + (size_t)b * ns;
        float* pc = o.pc + (size_t)b * ns;
        float* pnew = o.P + ((size_t)b * m + slot_new) * ns;
        for (int e = tid; e < ns; e += nthr) {
            const float* sp = slices + (size_t)b * inst_stride + e;
            float d = 0.f;
            int k = 0;
            for (; k + 8 <= n_slices; k += 8) {                   // 8 independent loads in flight
                float v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = sp[(size_t)(k + u) * slice_stride];
#pragma unroll
                for (int u = 0; u < 8; ++u) d += v[u];
            }
            for (; k < n_slices; ++k) d += sp[(size_t)k * slice_stride];
            const double xold_f = (double)(float)xm[e];          // the partials are relative to float(xmean)
            const double shift = (xold_f - xm[e]) + (double)d;    // new mean - old mean
            xm[e] = xm[e] + shift;
            const float pcn = (float)((1.0 - o.cc) * (double)pc[e] + coef * shift);
            pc[e] = pcn;
            pnew[e] = pcn;
        }
    }
    __syncthreads();

    // ---- recompute v from the first stale position (lmcma.cpp:373-390) ----
    float* Vb = o.V + (size_t)b * m * ns;
    const float* Pb = o.P + (size_t)b * m * ns;
    const float Kf = (float)o.K;
    const int nq = ns >> 2;
    // row i of the sequence: pending rows (i >= first_stale) sit in shared memory while they fit, the
    // untouched older rows (and any overflow) are addressed in place in HBM/L2
    auto row_ptr = [&](int i) -> float* {
        const int p = i - first_stale;
        return (p >= 0 && p < cap_rows) ? rows_s + (size_t)p * ns : Vb + (size_t)order[i] * ns;
    };
    for (int i = first_stale + warp; i < live; i += nwarps) {      // pending rows start as pc_j
        const float4* src = reinterpret_cast<const float4*>(Pb + (size_t)order[i] * ns);
        float4* dst = reinterpret_cast<float4*>(row_ptr(i));
        for (int q = lane; q < nq; q += 32) dst[q] = src[q];
    }
    // Closed forms of lmcma.cpp:386-389 with t = sqrt(1 + c1/(1-c1) |v|^2), rewritten without cancellation:
    //   Nj = (sqrt(1-c1)/|v|^2)(t - 1)            = sqrt(1-c1) * r / (t + 1)
    //   Lj = (1/(sqrt(1-c1)|v|^2))(1 - 1/t)       = r / (sqrt(1-c1) * t * (t + 1)),   r = c1/(1-c1)
    // (identical in exact arithmetic; finite where the reference divides 0/0 for a zero vector).  The FP32
    // value of Lj feeds the next factor step; the FP64 state is filled in after the loop, off the critical path.
    const float r_f = (float)(o.c1 / (1.0 - o.c1)), a_f = (float)o.M;
    auto finalize = [&](int i) {       // warp-collective: |v|^2 and Lj of sequence row i
        const float4* v = reinterpret_cast<const float4*>(row_ptr(i));
        float nvf = 0.f;
        for (int q = lane; q < nq; q += 32) { const float4 x = v[q]; nvf += fmaf(x.x, x.x, x.y * x.y) + fmaf(x.z, x.z, x.w * x.w); }
        nvf = warp_sum(nvf);
        if (lane == 0) {
            const float t = sqrtf(fmaf(r_f, nvf, 1.0f));
            nv_s[i] = nvf;
            lj_s[i] = r_f / (a_f * t * (t + 1.0f));
        }
    };
    __syncthreads();
    if (first_stale == 0 && warp == 0 && live > 0) finalize(0);   // row 0 has no factors
    for (int j = 0; j + 1 < live; ++j) {
        __syncthreads();                                           // row j (and its Lj) is final
        const float4* vj = reinterpret_cast<const float4*>(row_ptr(j));
        const float lj = lj_s[j];
        const int i_begin = max(j + 1, first_stale);
        for (int i = i_begin + warp; i < live; i += nwarps) {
            float4* vi = reinterpret_cast<float4*>(row_ptr(i));
            float d = 0.f;
            for (int q = lane; q < nq; q += 32) {
                const float4 a4 = vj[q], c4 = vi[q];
                d += fmaf(a4.x, c4.x, a4.y * c4.y) + fmaf(a4.z, c4.z, a4.w * c4.w);
            }
            d = lj * warp_sum(d);
            for (int q = lane; q < nq; q += 32) {
                const float4 a4 = vj[q]; float4 c4 = vi[q];
                c4.x = fmaf(Kf, c4.x, -d * a4.x);
                c4.y = fmaf(Kf, c4.y, -d * a4.y);
                c4.z = fmaf(Kf, c4.z, -d * a4.z);
                c4.w = fmaf(Kf, c4.w, -d * a4.w);
                vi[q] = c4;
            }
            if (i == j + 1) { __syncwarp(); finalize(i); }
        }
    }
    __syncthreads();
    for (int i = first_stale + tid; i < live; i += nthr) {        // FP64 scalar state of the recomputed rows
        const int slot = order[i];
        const double nv = (double)nv_s[i], r = o.c1 / (1.0 - o.c1), a = o.M;
        const double t = sqrt(1.0 + r * nv);
        const double nj = a * r / (t + 1.0), lj = r / (a * t * (t + 1.0));
        Njd[slot] = nj; Ljd[slot] = lj; Njf[slot] = (float)nj;
    }
    for (int i = first_stale + warp; i < live && i - first_stale < cap_rows; i += nwarps) {   // write back
        const float4* src = reinterpret_cast<const float4*>(rows_s + (size_t)(i - first_stale) * ns);
        float4* dst = reinterpret_cast<float4*>(Vb + (size_t)order[i] * ns);
        for (int q = lane; q < nq; q += 32) dst[q] = src[q];
    }

    // ---- population-success step size (lmcma.cpp:393-419), bookkeeping (lmcma.cpp:420-423, 192-194) ----
    if (tid == 0) {
        unsigned long long S = 0;
        if (payload_mode) {
            for (int k = 0; k < n_slices; ++k) {
                const float* pay = slices + (size_t)k * slice_stride + (size_t)b * inst_stride + ns;
                S += ((unsigned long long)__float_as_uint(pay[1]) << 32) | __float_as_uint(pay[0]);
            }
        } else {
            S = o.S_count[b];
            o.S_count[b] = 0ull;
        }
        const int itr = scp->itr;
        if (itr > 0) {
            const double lam = (double)o.lambda;
            const unsigned long long L = (unsigned long long)o.lambda;
            const unsigned long long sum_cur = L * (L - 1ull) / 2ull + S;      // ranks of this generation in the merged order
            const unsigned long long sum_prev = L * (2ull * L - 1ull) - sum_cur;
            const double mean_cur = (double)sum_cur / lam, mean_prev = (double)sum_prev / lam;
            const double success = (mean_prev - mean_cur) / lam;
            const double snew = (1.0 - o.cs) * scp->s + o.cs * (success - o.target);
            scp->s = snew;
            scp->sigma = sigma_old * exp(snew);
        }
        scp->itr = itr + 1;
        scp->live = live;
        scp->counteval += o.lambda;
    }
    // best-so-far (strict improvement, or the very first evaluation: lmcma.cpp:192)
    {
        const float* fa = f_all + (size_t)b * o.lambda;
        __shared__ int sh_best_row; __shared__ int sh_take;
        // the rank-0 candidate: first occurrence of the minimum in evaluation order (block arg-min)
        __shared__ float am_v[32]; __shared__ int am_i[32];
        float bf = __int_as_float(0x7f800000); int bi = 0x7fffffff;
        for (int j = tid; j < o.lambda; j += nthr) { const float v = canon_fitness(fa[j]); if (v < bf || (v == bf && j < bi)) { bf = v; bi = j; } }
#pragma unroll
        for (int ofs = 16; ofs > 0; ofs >>= 1) {
            const float v2 = __shfl_xor_sync(0xffffffffu, bf, ofs); const int i2 = __shfl_xor_sync(0xffffffffu, bi, ofs);
            if (v2 < bf || (v2 == bf && i2 < bi)) { bf = v2; bi = i2; }
        }
        if (lane == 0) { am_v[warp] = bf; am_i[warp] = bi; }
        __syncthreads();
        if (tid == 0) {
            for (int w2 = 1; w2 < nwarps; ++w2) if (am_v[w2] < bf || (am_v[w2] == bf && am_i[w2] < bi)) { bf = am_v[w2]; bi = am_i[w2]; }
            if (bi == 0x7fffffff) bi = 0;
            const bool take = ((double)bf < scp->best_f) || (scp->counteval == o.lambda);
            sh_take = take ? 1 : 0; sh_best_row = bi;
            if (take) { scp->best_f = (double)bf; scp->best_local = (bi >= o.pop_offset && bi < o.pop_offset + o.pop_count) ? 1 : 0; }
        }
        __syncthreads();
        if (sh_take && sh_best_row >= o.pop_offset && sh_best_row < o.pop_offset + o.pop_count) {
            const float* src = o.X + ((size_t)b * o.pop_count + (sh_best_row - o.pop_offset)) * ns;
            float* dst = o.best_x + (size_t)b * ns;
            for (int e = tid; e < ns; e += nthr) dst[e] = src[e];
        }
        for (int j = tid; j < o.lambda; j += nthr) o.prev_fit[(size_t)b * o.lambda + j] = canon_fitness(fa[j]);
    }
}

}  // namespace lmcma
