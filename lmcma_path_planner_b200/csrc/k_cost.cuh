// k_cost.cuh — batched trajectory cost kernel.
#pragma once
#include "lmcma_common.cuh"

namespace lmcma {

// ------------------------------------------------------------------------------------------------
// k_cost — batched trajectory cost (DESIGN.md "cost model"; the reference's per-state pieces are
// ValidityChecker::isValid/clearance planner.cpp:591-631, ClearanceObjective::stateCost :655-669,
// weights :677-690).  One CTA per trajectory.
// Phase 1: one thread per segment: record {A, B-A, 1/K, len/K, first sample} into shared memory, block scan of the
//   sample counts, and the two END samples of the segment (k = 0, k = K) fetched here — see below.
// Phase 2: the flattened sample sequence is cut into 32-sample blocks, a contiguous run of blocks per warp,
//   consecutive samples on consecutive lanes: they are <= 1 cell apart, and the map is stored in 128-byte bricks
//   (lmcma_layout.hpp), so one warp load touches a handful of lines instead of 32 (the L1 wavefront rate, not
//   DRAM, bounds a row-major gather).  Every sample is accumulated with the INTERIOR trapezoid weight len/K and
//   counted as a potential collision; phase 1's end samples then take back half a weight at k = 0 and k = K and the
//   doubly counted joint (k = K of every segment but the last) — that keeps weights, end-point tests and K itself
//   out of the 32-sample loop.  The loop is software-pipelined: the loads of two blocks are in flight while the
//   previous two are accumulated (the kernel was stalled on the load-to-use latency, not on issue slots).
//   Trajectories whose waypoints all lie inside the map (always, for candidates clamped by the optimiser's box
//   bounds) take a loop without bounds tests; others take the CHECK variant (out of the map = collision).
// Phase 3: block reduction.
// The index path (t = k * (1/K); q = A + t*d; rint; bounds test) uses explicitly rounded FP32
// mul/add so that no FMA contraction can change a cell index relative to the CPU oracle.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int substeps_of(float linf) {
    if (linf >= 1.0f) return linf <= (float)KMAX_SUBSTEPS ? (int)ceilf(linf) : KMAX_SUBSTEPS;
    return 1;   // also NaN
}

constexpr int COST_MAX_WARPS = 8;

// what a map load returns before it is turned into the sign-tagged reciprocal clearance
template <int STORAGE> struct RawCell { typedef float type; };
template <> struct RawCell<1> { typedef unsigned char type; };

template <int DIMS, int STORAGE, bool TRACE>
// 8 warps per CTA, 7 CTAs per SM: 32 registers per thread, every trajectory of a 1024-wide population resident at once
__global__ void __launch_bounds__(COST_MAX_WARPS * 32, 7) k_cost(MapDev mp, CostArgs a) {
    typedef typename RawCell<STORAGE>::type raw_t;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int W = a.W, NSEG = W + 1;
    float4* segA = reinterpret_cast<float4*>(smem_raw);          // 2-D {Ax, Ay, dx, dy}   3-D {Ax, Ay, Az, dx}
    float4* segB = segA + NSEG;                                  // 2-D {invK, len/K, first sample, -} 3-D {dy, dz, invK, len/K}
    int* segC = reinterpret_cast<int*>(segB + NSEG);             // 3-D first sample (NSEG entries, unused in 2-D)
    int* off = segC + NSEG;                                      // NSEG + 1 exclusive sample offsets + 32 x T padding
    float* lut = reinterpret_cast<float*>(off + NSEG + 1 + 32);  // 256 (U8 only)
    __shared__ float red_f[3][COST_MAX_WARPS];
    __shared__ int red_i[COST_MAX_WARPS];
    __shared__ int warp_tot[COST_MAX_WARPS];

    griddep_launch_dependents();          // k_rank may become resident as this grid drains; it waits for completion
    const int row = blockIdx.x, b = blockIdx.y;
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
    const float* x = a.X + ((size_t)b * a.inst_rows + row) * a.ld;
    const float* en = a.ends + (a.ends_per_instance ? (size_t)b * 6 : 0);
    if (STORAGE == 1) for (int i = tid; i < 256; i += nthr) lut[i] = mp.lut[i];

    const unsigned nxm1 = (unsigned)(mp.nx - 1), nym1 = (unsigned)(mp.ny - 1), nzm1 = (unsigned)(mp.nz - 1);
    const float fx1 = (float)(mp.nx - 1), fy1 = (float)(mp.ny - 1), fz1 = (float)(mp.nz - 1);
    const unsigned py = mp.py, pz = mp.pz;                       // brick pitches (lmcma_layout.hpp), formed on the host
    const float ng_coll = -mp.g_coll;
    const float* __restrict__ g32 = mp.g32;
    const unsigned char* __restrict__ q8 = mp.q8;
    const int last = NSEG - 1;

    // sample k * invK of a segment -> cell; returns false outside the map.  Round-half-even conversion == (int)rintf(q);
    // it saturates for huge |q| (-> fails the unsigned test); explicitly rounded mul / add (see the header)
    auto cell_of = [&](float tk, float ax, float ay, float az, float dx, float dy, float dz, int& ix, int& iy, int& iz) {
        ix = __float2int_rn(__fadd_rn(ax, __fmul_rn(tk, dx)));
        iy = __float2int_rn(__fadd_rn(ay, __fmul_rn(tk, dy)));
        bool inb = ((unsigned)ix <= nxm1) && ((unsigned)iy <= nym1);
        iz = 0;
        if (DIMS == 3) {
            iz = __float2int_rn(__fadd_rn(az, __fmul_rn(tk, dz)));
            inb = inb && ((unsigned)iz <= nzm1);
        }
        return inb;
    };
    auto load_raw = [&](unsigned adr) -> raw_t {
        if (STORAGE == 0) return (raw_t)__ldg(g32 + adr);
        return (raw_t)__ldg(q8 + adr);
    };
    auto value_of = [&](raw_t r) -> float { return STORAGE == 0 ? (float)r : lut[(int)r]; };   // lut is read after a barrier

    // ---- phase 1: per-segment records, sub-step counts, lengths, end samples; block scan of the sample counts ----
    const int spt = (NSEG + nthr - 1) / nthr;                    // consecutive segments per thread
    const int s_begin = min(NSEG, tid * spt), s_end = min(NSEG, s_begin + spt);
    float len_acc = 0.f; int my_cnt = 0;
    bool safe = true;                                            // every waypoint of my segments inside the map
    // end samples of the thread's most recent segment, taken back from the sums once their loads have landed
    raw_t e0 = 0, eK = 0; bool e0_out = false, eK_out = false, eK_joint = false; float e_half = 0.f;
    float corr_clr = 0.f; int corr_coll = 0;
    auto take_back = [&]() {
        const float g0 = e0_out ? ng_coll : value_of(e0), gK = eK_out ? ng_coll : value_of(eK);
        corr_clr = fmaf(e_half, fabsf(g0) + fabsf(gK), corr_clr);
        corr_coll += (eK_joint && gK < 0.f) ? 1 : 0;
    };
    if (STORAGE == 1) __syncthreads();                           // the lut
    for (int s = s_begin; s < s_end; ++s) {
        if (s > s_begin) take_back();
        float A[3], D[3]; float linf = 0.f, l2 = 0.f; bool bad = false;
        A[2] = 0.f; D[2] = 0.f;
#pragma unroll
        for (int c = 0; c < DIMS; ++c) {
            A[c] = (s == 0) ? en[c] : x[c * W + s - 1];
            const float Bc = (s == W) ? en[3 + c] : x[c * W + s];
            const float hi_c = c == 0 ? fx1 : (c == 1 ? fy1 : fz1);
            safe = safe && (A[c] >= 0.f) && (A[c] <= hi_c) && (Bc >= 0.f) && (Bc <= hi_c);
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
        const float scale = len * invK;
        if (DIMS == 2) { segA[s] = make_float4(A[0], A[1], D[0], D[1]); segB[s] = make_float4(invK, scale, 0.f, 0.f); }
        else { segA[s] = make_float4(A[0], A[1], A[2], D[0]); segB[s] = make_float4(D[1], D[2], invK, scale); }
        off[s + 1] = K + 1;                                      // samples of this segment (rewritten below)
        my_cnt += K + 1;
        // the two end samples (k = 0 and k = K), by the same arithmetic as the sample loop
        int ix, iy, iz;
        e0_out = !cell_of(__fmul_rn(0.f, invK), A[0], A[1], A[2], D[0], D[1], D[2], ix, iy, iz);
        e0 = load_raw(e0_out ? 0u : brick_offset_p<DIMS, STORAGE>((unsigned)ix, (unsigned)iy, (unsigned)iz, py, pz));
        eK_out = !cell_of(__fmul_rn((float)K, invK), A[0], A[1], A[2], D[0], D[1], D[2], ix, iy, iz);
        eK = load_raw(eK_out ? 0u : brick_offset_p<DIMS, STORAGE>((unsigned)ix, (unsigned)iy, (unsigned)iz, py, pz));
        e_half = 0.5f * scale;
        eK_joint = s != last;
    }
    int incl = my_cnt;                                           // block-wide exclusive scan of my_cnt
#pragma unroll
    for (int o2 = 1; o2 < 32; o2 <<= 1) {
        const int nb = __shfl_up_sync(0xffffffffu, incl, o2);
        if (lane >= o2) incl += nb;
    }
    if (lane == 31) warp_tot[warp] = incl;
    const bool all_safe = __syncthreads_and(safe ? 1 : 0) != 0;
    int base = incl - my_cnt, T = 0;
    for (int w2 = 0; w2 < nwarps; ++w2) { const int v = warp_tot[w2]; if (w2 < warp) base += v; T += v; }
    if (tid == 0) off[0] = 0;
    for (int s = s_begin; s < s_end; ++s) {
        if (DIMS == 2) segB[s].z = __int_as_float(base); else segC[s] = base;   // first sample of the segment
        base += off[s + 1]; off[s + 1] = base;
    }
    if (tid < 32) off[NSEG + 1 + tid] = T;                       // padding read by the 32-wide end-offset loads
    __syncthreads();

    // ---- phase 2: consecutive samples on consecutive lanes (<= 1 cell apart -> few lines per warp load).  The
    //      segment of every lane's sample comes from ONE warp-wide step: the 32 next segment-end offsets are
    //      loaded one per lane, the ends that fall inside this 32-sample block are OR-reduced into a bit mask
    //      (redux.sync), and a lane's segment is the warp's first segment + popc(mask bits below the lane) ----
    const unsigned nblk = (unsigned)(T + 31) >> 5;
    const int blk0 = (int)(((unsigned)warp * nblk) / (unsigned)nwarps), blk1 = (int)(((unsigned)(warp + 1) * nblk) / (unsigned)nwarps);
    float clr_acc = 0.f; int coll = 0;
    if (blk0 < blk1) {
        int lo = 0, hi = NSEG - 1;                               // warp-uniform: last s with off[s] <= first sample
        const int tfirst = blk0 << 5;
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (off[mid] <= tfirst) lo = mid; else hi = mid - 1; }
        int s_warp = lo;
        const unsigned below = (1u << lane) - 1u;
        // one 32-sample block: address + load issue.  TAIL = the trajectory's last block, where the lanes past the last
        // sample contribute nothing (weight 0, not a collision).  CHECK = bounds tests (a sample outside the map is a collision)
        auto fetch = [&](const int blk, const bool TAIL, const bool CHECK, raw_t& raw, float& wgt, bool& outside) {
            const int t0 = blk << 5;
            // segment ends at t0 + 1 .. t0 + 32 -> bits 0 .. 31 (off[] is padded with T beyond NSEG; shl clamps >= 32 to 0)
            const unsigned rel1 = (unsigned)(off[s_warp + 1 + lane] - t0 - 1);
            unsigned bit;
            asm("shl.b32 %0, 1, %1;" : "=r"(bit) : "r"(rel1));
            const unsigned mask = __reduce_or_sync(0xffffffffu, bit);
            const bool dead = TAIL && (t0 + lane >= T);
            const int tl = TAIL ? min(lane, T - 1 - t0) : lane;
            const int s = s_warp + __popc(mask & (TAIL ? ((1u << tl) - 1u) : below));   // ends at or before this lane's sample
            s_warp += __popc(mask);
            const float4 ra = segA[s];
            const float4 rb = segB[s];
            float invK, scale, ax = ra.x, ay = ra.y, az = 0.f, dx, dy, dz = 0.f; int cur;
            if (DIMS == 2) { invK = rb.x; scale = rb.y; cur = __float_as_int(rb.z); dx = ra.z; dy = ra.w; }
            else { dx = ra.w; dy = rb.x; dz = rb.y; invK = rb.z; scale = rb.w; az = ra.z; cur = segC[s]; }
            const int k = t0 + tl - cur;
            int ix, iy, iz;
            const bool inb = cell_of(__fmul_rn((float)k, invK), ax, ay, az, dx, dy, dz, ix, iy, iz);
            unsigned adr = brick_offset_p<DIMS, STORAGE>((unsigned)ix, (unsigned)iy, (unsigned)iz, py, pz);
            if (CHECK) adr = inb ? adr : 0u;
            raw = load_raw(adr);
            outside = CHECK && !inb;
            wgt = dead ? 0.f : scale;
            if (TRACE) {
                if (!dead && (long long)(t0 + lane) < a.max_cells) a.cells[t0 + lane] = inb ? ((long long)iz * mp.ny + iy) * mp.nx + ix : -1;
            }
        };
        // |g| x weight into the clearance sum; a collision is the sign bit of g (obstacle cells and out-of-map samples
        // are negative, never -0)
        auto accumulate = [&](raw_t raw, float wgt, bool outside, const bool CHECK) {
            float g = value_of(raw);
            if (CHECK) g = outside ? ng_coll : g;
            clr_acc = fmaf(fabsf(g), wgt, clr_acc);
            coll += (int)(__float_as_uint(g) >> 31);
        };
        auto run = [&](const bool CHECK) {
            const int blk_full_end = min(blk1, (int)nblk - 1);   // the trajectory's last block may be partial
            const raw_t idle = (raw_t)(STORAGE == 1 ? 1 : 0);     // a free cell: nothing in flight yet (weight 0, no collision)
            raw_t p0 = idle, p1 = idle; float w0 = 0.f, w1 = 0.f; bool o0 = false, o1 = false;   // the two blocks in flight
            int blk = blk0;
            for (; blk + 1 < blk_full_end; blk += 2) {
                raw_t r0, r1; float v0, v1; bool q0, q1;
                fetch(blk, false, CHECK, r0, v0, q0);
                fetch(blk + 1, false, CHECK, r1, v1, q1);
                accumulate(p0, w0, o0, CHECK);
                accumulate(p1, w1, o1, CHECK);
                p0 = r0; w0 = v0; o0 = q0; p1 = r1; w1 = v1; o1 = q1;
            }
            // at most one full block and the partial last block remain
            raw_t r0 = idle, r1 = idle; float v0 = 0.f, v1 = 0.f; bool q0 = false, q1 = false;
            if (blk < blk_full_end) fetch(blk, false, CHECK, r0, v0, q0);
            if (blk1 == (int)nblk) {
                fetch((int)nblk - 1, true, CHECK, r1, v1, q1);
                if ((int)((nblk - 1) << 5) + lane >= T) { r1 = idle; q1 = false; }   // lanes past the last sample
            }
            accumulate(p0, w0, o0, CHECK);
            accumulate(p1, w1, o1, CHECK);
            accumulate(r0, v0, q0, CHECK);
            accumulate(r1, v1, q1, CHECK);
        };
        if (all_safe) run(false); else run(true);
    }
    if (s_begin < s_end) take_back();

    // ---- phase 3: block reduction ----
    len_acc = warp_sum(len_acc);
    clr_acc = warp_sum(clr_acc);
    corr_clr = warp_sum(corr_clr);
    coll = warp_sum_i(coll - corr_coll);
    if (lane == 0) { red_f[0][warp] = len_acc; red_f[1][warp] = clr_acc; red_f[2][warp] = corr_clr; red_i[warp] = coll; }
    __syncthreads();
    if (tid == 0) {
        float L = 0.f, C = 0.f, Cc = 0.f; int NC = 0;
        for (int w2 = 0; w2 < nwarps; ++w2) { L += red_f[0][w2]; C += red_f[1][w2]; Cc += red_f[2][w2]; NC += red_i[w2]; }
        const float f = fmaf(a.w_col, (float)NC, fmaf(a.w_clr, C - Cc, a.w_len * L));
        a.f[(size_t)b * a.f_stride + a.f_offset + row] = f;
        if (a.ncoll) a.ncoll[(size_t)b * a.inst_rows + row] = NC;
        if (a.nsamp) a.nsamp[(size_t)b * a.inst_rows + row] = T;
    }
}

}  // namespace lmcma
