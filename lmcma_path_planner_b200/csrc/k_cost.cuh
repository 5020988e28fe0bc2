// k_cost.cuh — batched trajectory cost kernel.
#pragma once
#include "lmcma_common.cuh"

namespace lmcma {

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
// 7 warps per CTA, 7 CTAs per SM: 40 registers per thread with every trajectory of a 1024-wide population resident at once
__global__ void __launch_bounds__(224, 7) k_cost(MapDev mp, CostArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int W = a.W, NSEG = W + 1;
    float4* segA = reinterpret_cast<float4*>(smem_raw);          // 2-D {Ax, Ay, dx, dy}   3-D {Ax, Ay, Az, dx}
    float4* segB = segA + NSEG;                                  // 2-D {invK, scale, K, first sample} 3-D {dy, dz, invK, scale}
    int2* segC = reinterpret_cast<int2*>(segB + NSEG);           // 3-D {K, first sample} (NSEG entries, unused in 2-D)
    int* off = reinterpret_cast<int*>(segC + NSEG);              // NSEG + 1 exclusive sample offsets + 32 x T padding
    float* lut = reinterpret_cast<float*>(off + NSEG + 1 + 32);  // 256 (U8 only)
    __shared__ float red_f[2][8];
    __shared__ int red_i[8];
    __shared__ int warp_tot[8];

    griddep_launch_dependents();          // k_rank may become resident as this grid drains; it waits for completion
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
        if (DIMS == 2) { segA[s] = make_float4(A[0], A[1], D[0], D[1]); segB[s] = make_float4(invK, len * invK, __int_as_float(K), 0.f); }
        else { segA[s] = make_float4(A[0], A[1], A[2], D[0]); segB[s] = make_float4(D[1], D[2], invK, len * invK); segC[s].x = K; }
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
    for (int s = s_begin; s < s_end; ++s) {
        if (DIMS == 2) segB[s].w = __int_as_float(base); else segC[s].y = base;   // first sample of the segment
        base += off[s + 1]; off[s + 1] = base;
    }
    if (tid < 32) off[NSEG + 1 + tid] = T;                       // padding read by the 32-wide end-offset loads
    __syncthreads();

    // ---- phase 2: consecutive samples on consecutive lanes (<= 1 cell apart -> few lines per warp load).  The
    //      segment of every lane's sample comes from ONE warp-wide step: the 32 next segment-end offsets are
    //      loaded one per lane, the ends that fall inside this 32-sample block are OR-reduced into a bit mask
    //      (redux.sync), and a lane's segment is the warp's first segment + popc(mask bits below the lane).
    //      The loop is issue-slot bound (the map is L2-resident), so it is kept to ~50 instructions per block:
    //      no per-lane search, bounds / tail handling hoisted to warp-uniform rare paths ----
    const int nblk = (T + 31) >> 5;
    const int blk0 = (int)(((long long)warp * nblk) / nwarps), blk1 = (int)(((long long)(warp + 1) * nblk) / nwarps);
    float clr_acc = 0.f; int coll = 0;
    if (blk0 < blk1) {
        int lo = 0, hi = NSEG - 1;                               // warp-uniform: last s with off[s] <= first sample
        const int tfirst = blk0 << 5;
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (off[mid] <= tfirst) lo = mid; else hi = mid - 1; }
        int s_warp = lo;
        const unsigned nxm1 = (unsigned)(mp.nx - 1), nym1 = (unsigned)(mp.ny - 1), nzm1 = (unsigned)(mp.nz - 1);
        const unsigned nbx = mp.nbx, nby = mp.nby;
        const int last = NSEG - 1;
        const float ng_coll = -mp.g_coll;
        const float* __restrict__ g32 = mp.g32;
        const unsigned char* __restrict__ q8 = mp.q8;
        // one 32-sample block; TAIL = the trajectory's last block, where lanes past the last sample repeat it with weight 0
        auto block = [&](const int blk, const bool TAIL) {
            const int t0 = blk << 5;
            // segment ends at t0 + 1 .. t0 + 32 -> bits 0 .. 31 (off[] is padded with T beyond NSEG; shl clamps >= 32 to 0)
            const unsigned rel1 = (unsigned)(off[s_warp + 1 + lane] - t0 - 1);
            unsigned bit;
            asm("shl.b32 %0, 1, %1;" : "=r"(bit) : "r"(rel1));
            const unsigned mask = __reduce_or_sync(0xffffffffu, bit);
            const int tl = TAIL ? min(lane, T - 1 - t0) : lane;
            const int s = s_warp + __popc(mask & ((1u << tl) - 1u));   // ends at or before this lane's sample
            s_warp += __popc(mask);
            const float4 ra = segA[s];
            const float4 rb = segB[s];
            float invK, scale, dx, dy, dz = 0.f, az = 0.f; int K, cur;
            if (DIMS == 2) { invK = rb.x; scale = rb.y; K = __float_as_int(rb.z); cur = __float_as_int(rb.w); dx = ra.z; dy = ra.w; }
            else { const int2 rc = segC[s]; dx = ra.w; dy = rb.x; dz = rb.y; invK = rb.z; scale = rb.w; az = ra.z; K = rc.x; cur = rc.y; }
            const int k = t0 + tl - cur;
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
            unsigned adr = brick_offset<DIMS, STORAGE>((unsigned)ix, (unsigned)iy, (unsigned)iz, nbx, nby);
            const bool all_in = __all_sync(0xffffffffu, inb);    // the box bounds keep candidates inside: almost always true
            if (!all_in) adr = inb ? adr : 0u;
            float g = (STORAGE == 0) ? __ldg(g32 + adr) : lut[__ldg(q8 + adr)];
            if (!all_in) g = inb ? g : ng_coll;                   // a sample outside the map is a collision
            const bool endpt = (k == 0) || (k == K);
            float sw = endpt ? 0.5f * scale : scale;             // trapezoid weight x len/K
            bool counted = (k < K) || (s == last);               // every distinct poly-line sample once
            if (TAIL) { const bool valid = tl == lane; sw = valid ? sw : 0.f; counted = counted && valid; }
            clr_acc = fmaf(fabsf(g), sw, clr_acc);
            coll += (counted && g < 0.f) ? 1 : 0;
            if (TRACE) {
                if ((!TAIL || tl == lane) && (long long)(t0 + lane) < a.max_cells) a.cells[t0 + lane] = inb ? ((long long)iz * mp.ny + iy) * mp.nx + ix : -1;
            }
        };
        const int blk_full_end = min(blk1, nblk - 1);            // the trajectory's last block may be partial
        for (int blk = blk0; blk < blk_full_end; ++blk) block(blk, false);
        if (blk1 == nblk) block(nblk - 1, true);
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

}  // namespace lmcma
