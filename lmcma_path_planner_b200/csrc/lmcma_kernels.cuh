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
};

struct MapDev {
    int dims, nx, ny, nz;
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
    const int nchunks = (live + kc - 1) / kc;

    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_barrier_init();
    }
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

    // ---- stream the pairs ----
    const float Mf = (float)o.M;
    const float* njf = o.Njf + (size_t)b * o.m;
    for (int c = 0; c < nchunks; ++c) {
        const int st = c & 1;
        mbar_wait(&bars[st], (unsigned)((c >> 1) & 1));
        const float* sb = stage_base + st * stage_floats;
        const int k0 = c * kc, cnt = min(kc, live - k0);
        for (int k = 0; k < cnt; ++k) {
            const float4* v4 = reinterpret_cast<const float4*>(sb + (size_t)(2 * k) * ns);
            const float4* p4 = reinterpret_cast<const float4*>(sb + (size_t)(2 * k + 1) * ns);
            float d[RB];
#pragma unroll
            for (int r = 0; r < RB; ++r) d[r] = 0.f;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int q = lane + 32 * i;
                if (q < nq) {
                    const float4 v = v4[q];
#pragma unroll
                    for (int r = 0; r < RB; ++r) {
                        d[r] = fmaf(v.x, z[r][i].x, d[r]);
                        d[r] = fmaf(v.y, z[r][i].y, d[r]);
                        d[r] = fmaf(v.z, z[r][i].z, d[r]);
                        d[r] = fmaf(v.w, z[r][i].w, d[r]);
                    }
                }
            }
            const float nj = njf[order[k0 + k]];
#pragma unroll
            for (int r = 0; r < RB; ++r) d[r] = nj * warp_sum(d[r]);
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int q = lane + 32 * i;
                if (q < nq) {
                    const float4 p = p4[q];
#pragma unroll
                    for (int r = 0; r < RB; ++r) {
                        az[r][i].x = fmaf(Mf, az[r][i].x, d[r] * p.x);
                        az[r][i].y = fmaf(Mf, az[r][i].y, d[r] * p.y);
                        az[r][i].z = fmaf(Mf, az[r][i].z, d[r] * p.z);
                        az[r][i].w = fmaf(Mf, az[r][i].w, d[r] * p.w);
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
// weights :677-690).  One CTA per trajectory.  Phase 1: waypoints -> shared memory, per-segment
// sub-step counts + block scan.  Phase 2: the flattened sample sequence is cut into blockDim equal
// contiguous ranges; every thread walks its range along the poly-line (consecutive samples are <= 1
// cell apart, so its loads stay in the same / neighbouring sectors).  Phase 3: block reduction.
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
    const int W = a.W, NP = W + 2, NSEG = W + 1;
    float* pts = reinterpret_cast<float*>(smem_raw);            // DIMS x NP
    int* off = reinterpret_cast<int*>(pts + DIMS * NP);         // NSEG + 1 exclusive offsets
    float* lut = reinterpret_cast<float*>(off + NSEG + 1);      // 256 (U8 only)
    __shared__ float red_f[2][8];
    __shared__ int red_i[8];
    __shared__ int carry_s;

    const int row = blockIdx.x, b = blockIdx.y;
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const float* x = a.X + ((size_t)b * a.inst_rows + row) * a.ld;
    const float* en = a.ends + (a.ends_per_instance ? (size_t)b * 6 : 0);

    for (int i = tid; i < DIMS * W; i += nthr) {
        const int d = i / W, w = i - d * W;
        pts[d * NP + 1 + w] = x[i];
    }
    if (tid < DIMS) { pts[tid * NP] = en[tid]; pts[tid * NP + NP - 1] = en[3 + tid]; }
    if (STORAGE == 1) for (int i = tid; i < 256; i += nthr) lut[i] = mp.lut[i];
    __syncthreads();

    // ---- phase 1: sub-step counts, lengths ----
    float len_acc = 0.f;
    for (int s = tid; s < NSEG; s += nthr) {
        float linf = 0.f, l2 = 0.f; bool bad = false;
#pragma unroll
        for (int c = 0; c < DIMS; ++c) {
            const float d = __fsub_rn(pts[c * NP + s + 1], pts[c * NP + s]);
            const float ad = fabsf(d);
            bad |= (ad != ad);
            if (ad > linf) linf = ad;
            l2 = fmaf(d, d, l2);
        }
        if (bad) linf = __int_as_float(0x7fc00000);
        off[s + 1] = substeps_of(linf) + 1;     // samples of this segment
        len_acc += sqrtf(l2);
    }
    if (tid == 0) { off[0] = 0; carry_s = 0; }
    __syncthreads();
    if (warp == 0) {                              // inclusive scan of off[1..NSEG] by one warp
        int carry = 0;
        for (int base = 1; base <= NSEG; base += 32) {
            const int i = base + lane;
            int v = (i <= NSEG) ? off[i] : 0;
#pragma unroll
            for (int o2 = 1; o2 < 32; o2 <<= 1) {
                const int nb = __shfl_up_sync(0xffffffffu, v, o2);
                if (lane >= o2) v += nb;
            }
            if (i <= NSEG) off[i] = v + carry;
            carry += __shfl_sync(0xffffffffu, v, 31);
        }
        if (lane == 0) carry_s = carry;
    }
    __syncthreads();
    const int T = carry_s;

    // ---- phase 2: walk my contiguous sample range ----
    const int t0 = (int)(((long long)tid * T) / nthr), t1 = (int)(((long long)(tid + 1) * T) / nthr);
    float clr_acc = 0.f; int coll = 0;
    if (t0 < t1) {
        int lo = 0, hi = NSEG - 1;               // last s with off[s] <= t0
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (off[mid] <= t0) lo = mid; else hi = mid - 1; }
        int s = lo, k = t0 - off[s], t = t0;
        const float nxm1 = (float)(mp.nx - 1), nym1 = (float)(mp.ny - 1), nzm1 = (float)(mp.nz - 1);
        while (t < t1) {
            const int K = off[s + 1] - off[s] - 1;
            float A[DIMS], D[DIMS]; float l2 = 0.f;
#pragma unroll
            for (int c = 0; c < DIMS; ++c) {
                A[c] = pts[c * NP + s];
                D[c] = __fsub_rn(pts[c * NP + s + 1], A[c]);
                l2 = fmaf(D[c], D[c], l2);
            }
            const float invK = __frcp_rn((float)K);
            const float scale = sqrtf(l2) * invK;
            const int kend = min(K, k + (t1 - t) - 1);
            const bool last_seg = (s == NSEG - 1);
            float seg_acc = 0.f;
#pragma unroll 4
            for (int kk = k; kk <= kend; ++kk) {
                const float tk = __fmul_rn((float)kk, invK);
                const float rx = rintf(__fadd_rn(A[0], __fmul_rn(tk, D[0])));
                const float ry = rintf(__fadd_rn(A[1], __fmul_rn(tk, D[1])));
                bool inb = (rx >= 0.f) && (rx <= nxm1) && (ry >= 0.f) && (ry <= nym1);
                long long idx = (long long)(int)ry * mp.nx + (int)rx;
                if (DIMS == 3) {
                    const float rz = rintf(__fadd_rn(A[DIMS - 1], __fmul_rn(tk, D[DIMS - 1])));
                    inb = inb && (rz >= 0.f) && (rz <= nzm1);
                    idx += (long long)(int)rz * mp.nx * mp.ny;
                }
                float g = -mp.g_coll;
                if (inb) {
                    if (STORAGE == 0) g = __ldg(mp.g32 + idx);
                    else g = lut[__ldg(mp.q8 + idx)];
                }
                const float wgt = (kk == 0 || kk == K) ? 0.5f : 1.0f;
                seg_acc = fmaf(fabsf(g), wgt, seg_acc);
                coll += (g < 0.f) && (kk < K || last_seg);
                if (TRACE) { const long long tt = t + (kk - k); if (tt < a.max_cells) a.cells[tt] = inb ? idx : -1; }
            }
            clr_acc = fmaf(seg_acc, scale, clr_acc);
            t += kend - k + 1;
            k = 0; ++s;
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
        for (int w2 = 0; w2 < (nthr + 31) / 32; ++w2) { L += red_f[0][w2]; C += red_f[1][w2]; NC += red_i[w2]; }
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
// reference's comparator is undefined for NaN).  grid = (ceil(pop_count/256), B).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_rank(OptDev o, const float* __restrict__ f_all) {
    __shared__ float cur_s[1024];
    __shared__ float prev_s[1024];
    __shared__ unsigned long long red[8];
    const int b = blockIdx.y, tid = threadIdx.x;
    const int il = blockIdx.x * 256 + tid;                 // local row
    const int i = o.pop_offset + il;                       // global candidate id
    const bool valid = il < o.pop_count;
    const float* cur = f_all + (size_t)b * o.lambda;
    const float* prev = o.prev_fit + (size_t)b * o.lambda;
    const float ki = valid ? canon_fitness(cur[i]) : 0.f;
    int c_lt = 0; unsigned long long p_lt = 0;
    for (int base = 0; base < o.lambda; base += 1024) {
        const int cnt = min(1024, o.lambda - base);
        __syncthreads();
        for (int j = tid; j < cnt; j += 256) { cur_s[j] = canon_fitness(cur[base + j]); prev_s[j] = prev[base + j]; }
        __syncthreads();
        if (valid) {
            int pl = 0;
#pragma unroll 4
            for (int j = 0; j < cnt; ++j) {
                const float kj = cur_s[j];
                c_lt += (kj < ki) || (kj == ki && (base + j) < i);
                pl += prev_s[j] < ki;
            }
            p_lt += pl;
        }
    }
    if (valid) {
        o.rank[(size_t)b * o.lambda + i] = c_lt;
        o.arindex[(size_t)b * o.lambda + c_lt] = i;
        o.fit_sorted[(size_t)b * o.lambda + c_lt] = ki;
    }
    // block-reduce p_lt -> one integer atomic per CTA (order-independent, deterministic)
    unsigned lo32 = (unsigned)p_lt, hi32 = (unsigned)(p_lt >> 32);
    // p_lt <= lambda <= 2^31 per thread: reduce as 64-bit via two shuffles
#pragma unroll
    for (int ofs = 16; ofs > 0; ofs >>= 1) {
        const unsigned l2 = __shfl_xor_sync(0xffffffffu, lo32, ofs), h2 = __shfl_xor_sync(0xffffffffu, hi32, ofs);
        unsigned long long a = ((unsigned long long)hi32 << 32) | lo32, c = ((unsigned long long)h2 << 32) | l2;
        a += c; lo32 = (unsigned)a; hi32 = (unsigned)(a >> 32);
    }
    if ((tid & 31) == 0) red[tid >> 5] = ((unsigned long long)hi32 << 32) | lo32;
    __syncthreads();
    if (tid == 0) {
        unsigned long long tot = 0;
        for (int w2 = 0; w2 < 8; ++w2) tot += red[w2];
        atomicAdd(o.S_count + b, tot);
    }
}

// ------------------------------------------------------------------------------------------------
// k_recombine — the weighted recombination of LMCMA::update (lmcma.cpp:316-326) as partial sums of
// w_{rank(i)} * (x_i - xmean) over the rows this handle owns (candidate order, fixed split -> run-to-run
// deterministic).  Summing differences keeps the FP32 sum accurate relative to the mean SHIFT, which
// is what the evolution path needs (lmcma.cpp:327-329).  grid = (ceil(nq/128), RS, B), block 128.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_recombine(OptDev o) {
    const int b = blockIdx.z, rs = blockIdx.y;
    const int q = blockIdx.x * 128 + threadIdx.x, nq = o.ns >> 2;
    if (q >= nq) return;
    const int rows_per = (o.pop_count + o.RS - 1) / o.RS;
    const int r0 = rs * rows_per, r1 = min(o.pop_count, r0 + rows_per);
    const double* xm = o.xmean + (size_t)b * o.ns + 4 * q;
    const float4 m4 = make_float4((float)xm[0], (float)xm[1], (float)xm[2], (float)xm[3]);
    const int* rk = o.rank + (size_t)b * o.lambda + o.pop_offset;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = r0; r < r1; ++r) {
        const int rnk = rk[r];
        if (rnk < o.mu) {
            const float w = o.w[rnk];
            const float4 x = reinterpret_cast<const float4*>(o.X + ((size_t)b * o.pop_count + r) * o.ns)[q];
            acc.x = fmaf(w, x.x - m4.x, acc.x);
            acc.y = fmaf(w, x.y - m4.y, acc.y);
            acc.z = fmaf(w, x.z - m4.z, acc.z);
            acc.w = fmaf(w, x.w - m4.w, acc.w);
        }
    }
    reinterpret_cast<float4*>(o.partial + ((size_t)b * o.RS + rs) * o.ns)[q] = acc;
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
__global__ void __launch_bounds__(512) k_update(OptDev o, const float* __restrict__ slices, int n_slices,
                                                long long slice_stride, long long inst_stride,
                                                const float* __restrict__ f_all, int payload_mode) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    int* order = reinterpret_cast<int*>(smem_raw);               // m
    __shared__ int sh_first_stale, sh_live, sh_slot_new;
    __shared__ double sh_sigma_old;
    const int b = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
    const int m = o.m, ns = o.ns;
    Scalars* scp = o.sc + b;
    int* tg = o.t + (size_t)b * m;
    int* vg = o.vec + (size_t)b * m;

    // ---- slot bookkeeping (lmcma.cpp:331-364), data-independent integer logic, one thread ----
    if (tid == 0) {
        const int itr = scp->itr;
        int first_stale = 1;
        if (itr < m) {
            tg[itr] = itr;
        } else {
            int gap_min = vg[tg[1]] - vg[tg[0]];
            for (int j = 1; j < m - 1; ++j) {
                const int gap = vg[tg[j + 1]] - vg[tg[j]];
                if (gap < gap_min) { gap_min = gap; first_stale = j + 1; }
            }
            if (gap_min >= m /* maxsteps = nvectors, lmcma.cpp:267 */) first_stale = 0;
            if (first_stale != m - 1) {
                const int recycled = tg[first_stale];
                for (int j = first_stale; j < m - 1; ++j) tg[j] = tg[j + 1];
                tg[m - 1] = recycled;
            }
        }
        const int live = min(itr + 1, m);
        const int slot_new = tg[live - 1];
        vg[slot_new] = itr;
        if (first_stale == 1) first_stale = 0;                   // lmcma.cpp:373-374
        sh_first_stale = first_stale; sh_live = live; sh_slot_new = slot_new;
        sh_sigma_old = scp->sigma;
    }
    __syncthreads();
    for (int i = tid; i < m; i += nthr) order[i] = tg[i];
    const int first_stale = sh_first_stale, live = sh_live, slot_new = sh_slot_new;
    const double sigma_old = sh_sigma_old;

    // ---- mean, evolution path, new pc_j (lmcma.cpp:316-329, 365-366) ----
    {
        const double coef = sqrt(o.cc * (2.0 - o.cc) * o.mueff) / sigma_old;
        double* xm = o.xmean + (size_t)b * ns;
        float* pc = o.pc + (size_t)b * ns;
        float* pnew = o.P + ((size_t)b * m + slot_new) * ns;
        for (int e = tid; e < ns; e += nthr) {
            float d = 0.f;
            for (int k = 0; k < n_slices; ++k) d += slices[(size_t)k * slice_stride + (size_t)b * inst_stride + e];
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
    double* Njd = o.Nj + (size_t)b * m;
    double* Ljd = o.Lj + (size_t)b * m;
    float* Njf = o.Njf + (size_t)b * m;
    const float Kf = (float)o.K;
    const int nq = ns >> 2;
    for (int i = first_stale + warp; i < live; i += nwarps) {      // pending rows start as pc_j
        const float4* src = reinterpret_cast<const float4*>(Pb + (size_t)order[i] * ns);
        float4* dst = reinterpret_cast<float4*>(Vb + (size_t)order[i] * ns);
        for (int q = lane; q < nq; q += 32) dst[q] = src[q];
    }
    auto finalize = [&](int slot) {    // warp-collective: |v|^2 and the two closed forms, in FP64
        const float4* v = reinterpret_cast<const float4*>(Vb + (size_t)slot * ns);
        float nvf = 0.f;
        for (int q = lane; q < nq; q += 32) { const float4 x = v[q]; nvf = fmaf(x.x, x.x, fmaf(x.y, x.y, fmaf(x.z, x.z, fmaf(x.w, x.w, nvf)))); }
        nvf = warp_sum(nvf);
        if (lane == 0) {
            const double nv = (double)nvf, c1 = o.c1;
            const double nj = (sqrt(1.0 - c1) / nv) * (sqrt(1.0 + (c1 / (1.0 - c1)) * nv) - 1.0);
            const double lj = (1.0 / (sqrt(1.0 - c1) * nv)) * (1.0 - (1.0 / sqrt(1.0 + (c1 / (1.0 - c1)) * nv)));
            Njd[slot] = nj; Ljd[slot] = lj; Njf[slot] = (float)nj;
        }
    };
    __syncthreads();
    if (first_stale == 0 && warp == 0 && live > 0) finalize(order[0]);   // row 0 has no factors
    for (int j = 0; j + 1 < live; ++j) {
        __syncthreads();                                           // row j (and its Lj) is final
        const int slot_j = order[j];
        const float4* vj = reinterpret_cast<const float4*>(Vb + (size_t)slot_j * ns);
        const float lj = (float)Ljd[slot_j];
        const int i_begin = max(j + 1, first_stale);
        for (int i = i_begin + warp; i < live; i += nwarps) {
            float4* vi = reinterpret_cast<float4*>(Vb + (size_t)order[i] * ns);
            float d = 0.f;
            for (int q = lane; q < nq; q += 32) {
                const float4 a4 = vj[q], c4 = vi[q];
                d = fmaf(a4.x, c4.x, fmaf(a4.y, c4.y, fmaf(a4.z, c4.z, fmaf(a4.w, c4.w, d))));
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
            if (i == j + 1) { __syncwarp(); finalize(order[i]); }
        }
    }
    __syncthreads();

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
        __shared__ float am_v[16]; __shared__ int am_i[16];
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
