// lmcma_capi.cu — the C ABI declared in include/lmcma_b200.h: library / device queries, the optimiser entry points, launch
// configuration of the sampler / rank / update kernels, the per-generation CUDA graphs, the split-population stages.
// No CPU fallback: every compute entry point needs a CUDA device.  Cost map + evaluator: lmcma_capi_map.cu; state
// access + host-side reference pieces: lmcma_capi_state.cu; shared declarations: lmcma_internal.cuh.
#include "lmcma_internal.cuh"
#include "k_rank.cuh"
#include "k_sample.cuh"
#include "k_update.cuh"
#include "k_prior.cuh"
#include "k_gram.cuh"

using namespace lmcma;
using namespace lmcma_capi;

namespace lmcma_capi {
thread_local std::string g_err;
std::atomic<long long> g_launches{0};
DeviceProps g_props[64];
}  // namespace lmcma_capi
namespace lmcma {
// error reporting for the other translation units of the library (lmcma_ingest.cpp)
int set_error(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    lmcma_capi::g_err = buf;
    return code;
}
}  // namespace lmcma

namespace {
template <int NV, int RB, int MAXT, bool SPEC = false>
int launch_sample_t(lmcma_b200_opt* o, const OptDev& d, bool pdl, cudaStream_t st) {
    auto kern = k_sample<NV, RB, MAXT, SPEC>;
    if (o->smp_smem > 48 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)o->smp_smem));
    const int rows_per_cta = (o->smp_threads / 32) * RB;
    const int ctas = (o->d.pop_count + rows_per_cta - 1) / rows_per_cta;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(ctas, o->d.B); cfg.blockDim = dim3(o->smp_threads); cfg.dynamicSmemBytes = o->smp_smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    CU(cudaLaunchKernelEx(&cfg, kern, d, o->smp_kc, o->smp_stages));
    g_launches++;
    return 0;
}

template <int RBW, int MAXT>
int launch_sample_wide_t(lmcma_b200_opt* o, const OptDev& d, bool pdl, cudaStream_t st, int progressive) {
    auto kern = k_sample_wide<RBW, MAXT>;
    if (o->smp_smem > 48 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)o->smp_smem));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    const int rows_per_cta = o->smp_R * RBW;
    cfg.gridDim = dim3((o->d.pop_count + rows_per_cta - 1) / rows_per_cta, o->d.B); cfg.blockDim = dim3(32 * o->smp_R * o->smp_CW);
    cfg.dynamicSmemBytes = o->smp_smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    CU(cudaLaunchKernelEx(&cfg, kern, d, o->smp_kc, o->smp_stages, o->smp_R, o->smp_CW, o->smp_qpw, progressive));
    g_launches++;
    return 0;
}
template <int QW, int RL>
int launch_sample_rows_t(lmcma_b200_opt* o, const OptDev& d, bool pdl, cudaStream_t st) {
    auto kern = k_sample_rows<QW, RL>;
    if (o->smp_rows_smem > 48 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)o->smp_rows_smem));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((d.pop_count + 32 * RL - 1) / (32 * RL), d.B); cfg.blockDim = dim3(512);
    cfg.dynamicSmemBytes = o->smp_rows_smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    CU(cudaLaunchKernelEx(&cfg, kern, d, o->smp_rows_kc, o->smp_rows_stages, o->smp_rows_region));
    g_launches++;
    return 0;
}
int launch_sample_rows(lmcma_b200_opt* o, const OptDev& d, bool pdl, cudaStream_t st) {
    const bool two = o->smp_rows_rl == 2;
    switch (o->smp_rows_qw) {
        case 4: return two ? launch_sample_rows_t<4, 2>(o, d, pdl, st) : launch_sample_rows_t<4, 1>(o, d, pdl, st);
        case 7: return two ? launch_sample_rows_t<7, 2>(o, d, pdl, st) : launch_sample_rows_t<7, 1>(o, d, pdl, st);
        default: return two ? launch_sample_rows_t<8, 2>(o, d, pdl, st) : launch_sample_rows_t<8, 1>(o, d, pdl, st);
    }
}

int launch_sample_wide(lmcma_b200_opt* o, const OptDev& d, bool pdl, cudaStream_t st, int progressive) {
    switch (o->smp_RBW) {
        case 1: return launch_sample_wide_t<1, 1024>(o, d, pdl, st, progressive);
        case 2: return launch_sample_wide_t<2, 1024>(o, d, pdl, st, progressive);
        default: return launch_sample_wide_t<4, 512>(o, d, pdl, st, progressive);
    }
}

// the sequence-ordered (v, pc) mirror k_sample streams from is maintained by k_update; after create or a state
// setter it is rebuilt from the slot-indexed arrays
int ensure_mirror(lmcma_b200_opt* o, cudaStream_t st) {
    if (!o->mirror_dirty) return 0;
    k_pack_pairs<<<dim3(o->d.m, o->d.B), 128, 0, st>>>(o->d);
    g_launches++;
    CU(cudaGetLastError());
    o->mirror_dirty = false;
    return 0;
}

// pdl: launched as a programmatic dependent of the kernel enqueued just before it on `st` (k_update)
// progressive: that k_update was launched with UpdateArgs::progressive (only meaningful with pdl)
int launch_sample(lmcma_b200_opt* o, cudaStream_t st, bool pdl = false, bool progressive = false, int progressive_mode = 1) {
    if (o->mirror_dirty) { int rc = ensure_mirror(o, st); if (rc) return rc; pdl = false; }
    if (o->d_Lf) {   // smoothness prior: z <- L z for the whole population before computeAz (lmcma.cpp:216-217)
        const OptDev& d = o->d;
        if (o->cfg.rng == LMCMA_B200_RNG_PHILOX) {
            k_gauss<<<dim3((d.ns / 4 + 127) / 128, d.pop_count, d.B), 128, 0, st>>>(d);
            g_launches++;
        }
        const int rows = d.B * d.pop_count;
        k_prior<<<dim3((d.ns + 63) / 64, (rows + 63) / 64), 256, 0, st>>>(d.Z, o->d_Lf, d.Zc, rows, d.n, d.ns);
        g_launches++;
        CU(cudaGetLastError());
        pdl = false;
    }
    // the host mirror of X (lmcma_b200_ask_all_view) is written by the samplers of the host-buffer protocol only: the fused
    // on-device generations (lmcma_b200_run) and the split-population stages keep the candidates on the device
    OptDev d = o->d;
    if (!o->mirror_on || o->mirror_suppressed) d.Xh = nullptr;
    o->xh_fresh = d.Xh != nullptr;
    if (o->smp_rows) return launch_sample_rows(o, d, pdl, st);
    if (o->smp_wide) return launch_sample_wide(o, d, pdl, st, (pdl && progressive && o->progressive) ? progressive_mode : 0);
    switch (o->smp_nv) {
        case 1: return launch_sample_t<1, 4, 512>(o, d, pdl, st);
        case 2: return launch_sample_t<2, 4, 512>(o, d, pdl, st);
        case 4: return o->tune.sample_spec ? launch_sample_t<4, 2, 512, true>(o, d, pdl, st) : launch_sample_t<4, 2, 512>(o, d, pdl, st);
        case 8: return launch_sample_t<8, 1, 512>(o, d, pdl, st);
        case 12: return launch_sample_t<12, 1, 256>(o, d, pdl, st);
        case 16: return launch_sample_t<16, 1, 256>(o, d, pdl, st);
    }
    return fail(LMCMA_B200_ERR_ARG, "unsupported n for k_sample (nv=%d)", o->smp_nv);
}

int configure_sample(lmcma_b200_opt* o) {
    const int nq = o->d.ns / 4;
    const int need = (nq + 31) / 32;
    static const int nvs[] = {1, 2, 4, 8, 12, 16};
    o->smp_nv = 0;
    for (int v : nvs)
        if (v >= need) { o->smp_nv = v; break; }
    if (!o->smp_nv) return fail(LMCMA_B200_ERR_ARG, "n = %d too large (max 2048)", o->d.n);
    o->smp_rb = o->smp_nv <= 2 ? 4 : (o->smp_nv <= 4 ? 2 : 1);
    const int maxt = o->smp_nv >= 12 ? 256 : 512;
    // as many warps per CTA as keeps >= ~1 CTA per SM
    int threads = maxt;
    const int forced = o->tune.sample_threads;
    if (forced >= 32 && forced <= maxt && forced % 32 == 0) threads = forced;
    else
        while (threads > 64) {
            const int rows_per_cta = (threads / 32) * o->smp_rb;
            const long long ctas = (long long)((o->d.pop_count + rows_per_cta - 1) / rows_per_cta) * o->d.B;
            if (ctas * 4 >= (long long)o->props->sm_count * 3) break;   // bigger tiles re-read the pairs less often
            threads >>= 1;
        }
    o->smp_threads = threads;
    const size_t pair_bytes = (size_t)2 * o->d.ns * sizeof(float);
    // pairs per stage: one group of 8 when it fits (else 4 / 2 / 1); as many stages as the live pairs need, within
    // ~160 KB, so that for the common shapes every pair is requested up front and nothing is re-issued
    const size_t budget = (size_t)o->tune.sample_smem_kb * 1024;
    int kc = (int)(budget / 2 / pair_bytes);
    kc = kc >= 8 ? 8 : (kc >= 4 ? 4 : (kc >= 2 ? 2 : 1));
    // long rows (C4: 12 KB per pair): two stages of a FULL group of 8 still fit one SM (192 KB), and a 4-pair stage runs the
    // 8-wide group code half empty
    if (kc == 4 && 2 * 8 * pair_bytes + 24 * 1024 <= o->props->smem_optin && o->tune.sample_smem_kb_set == 0) kc = 8;
    o->smp_kc = kc;
    const int max_chunks = (o->d.m + kc - 1) / kc;
    o->smp_stages = (int)std::max<size_t>(2, std::min<size_t>(std::min(max_chunks, SAMPLE_MAX_STAGES), budget / (kc * pair_bytes)));   // >= 2 even beyond the budget
    o->smp_smem = (size_t)o->smp_stages * kc * pair_bytes + SAMPLE_MAX_STAGES * 8 + (size_t)(2 * o->d.m + 16) * sizeof(float);
    // one large population: split every row over CW column-warps (k_sample_wide) for occupancy
    o->smp_wide = false;
    if (o->d.pop_count >= 256 && nq > 32 && !o->tune.sample_narrow) {
        o->smp_wide = true;
        o->smp_CW = (nq + 31) / 32;
        o->smp_qpw = (nq + o->smp_CW - 1) / o->smp_CW;
        // rows per warp: 2 (4 for very large populations); row-groups per CTA so that the grid still covers the SMs
        o->smp_RBW = o->tune.sample_rbw ? o->tune.sample_rbw : (o->d.pop_count >= 1024 ? 4 : 2);
        if (o->smp_RBW != 1 && o->smp_RBW != 2) o->smp_RBW = 4;
        int R = std::max(1, std::min(15, (o->smp_RBW == 4 ? 16 : 32) / o->smp_CW));
        const int forced_R = o->tune.sample_r;
        if (forced_R >= 1 && forced_R <= R) R = forced_R;
        else
            while (R > 1 && (long long)((o->d.pop_count + R * o->smp_RBW - 1) / (R * o->smp_RBW)) * o->d.B * 4 < (long long)o->props->sm_count * 3) R >>= 1;
        o->smp_R = R;
        o->smp_smem += (size_t)o->smp_stages * ((kc + SAMPLE_G - 1) / SAMPLE_G) * o->smp_R * o->smp_CW * o->smp_RBW * SAMPLE_G * sizeof(float);
    }
    if (o->smp_smem > o->props->smem_optin) return fail(LMCMA_B200_ERR_ARG, "k_sample needs %zu B shared memory", o->smp_smem);
    // many rows (batched queries, or one very large population): rows on lanes, column slices on warps (k_sample_rows)
    o->smp_rows = false;
    const long long total_rows = (long long)o->d.B * o->d.pop_count;
    // one population: from lambda = 8192 on it beats the wide sampler even though the progressive hand-over / overlapped
    // generation (wide sampler only) is given up (fused generation 0.271 -> 0.253 ms at 8192, 1.71 -> 1.41 ms at 65536)
    const bool many = total_rows >= 4096 && o->d.pop_count >= 32 &&
                      (o->d.B > 1 || o->tune.sample_rows == 1 || (o->d.pop_count >= 8192 && o->d.pop_count == o->d.lambda));
    if (o->tune.sample_rows != 0 && many && nq <= 128 && o->d.m <= 256 && !o->d_Lf) {
        const int NW = 16;
        const int qper = (nq + NW - 1) / NW;
        o->smp_rows_qw = qper <= 4 ? 4 : (qper <= 7 ? 7 : 8);
        // two rows per lane when an instance has them and the grid still fills the SMs
        o->smp_rows_rl = (o->d.pop_count >= 64 && total_rows / 64 >= 2LL * o->props->sm_count) ? 2 : 1;
        const int rows = 32 * o->smp_rows_rl, kc2 = 8;
        const int chunks = (o->d.m + kc2 - 1) / kc2;
        const size_t stage_bytes = (size_t)kc2 * pair_bytes;
        const size_t fixed2 = SAMPLE_MAX_STAGES * 8 + (size_t)((o->d.m + kc2 + 3) & ~3) * 4 + (size_t)chunks * kc2 * rows * 4 + (size_t)NW * kc2 * rows * 4;
        const size_t tile_bytes = (size_t)rows * (o->d.ns + 4) * 4;
        if (fixed2 + std::max(2 * stage_bytes, tile_bytes) + 1024 <= o->props->smem_optin) {
            const size_t room = o->props->smem_optin - 1024 - fixed2;
            int stages = (int)std::min<size_t>(std::min(chunks, SAMPLE_MAX_STAGES), room / stage_bytes);
            stages = std::max(stages, 2);
            const size_t region = std::max((size_t)stages * stage_bytes, tile_bytes);
            o->smp_rows_kc = kc2; o->smp_rows_stages = stages; o->smp_rows_region = (int)(region / 4);
            o->smp_rows_smem = region + fixed2;
            o->smp_rows = true;
            o->smp_wide = false;
        }
    }
    return 0;
}

// ranks + recombination partial sums (k_rank.cuh); RANK_PACK is the split-population stage
// pdl: launched as a programmatic dependent of the kernel enqueued just before it on `st` (k_cost)
int launch_rank(lmcma_b200_opt* o, const float* f_all, int mode, float* payload, cudaStream_t st, bool pdl = false) {
    if (o->d.tile_sorted) {                                       // lambda > 4096, unsplit: sort the fitness tiles first (k_rank.cuh)
        cudaLaunchConfig_t c0;
        memset(&c0, 0, sizeof(c0));
        c0.gridDim = dim3((o->d.lambda + TELL_FTILE - 1) / TELL_FTILE, o->d.B); c0.blockDim = dim3(1024); c0.stream = st;
        cudaLaunchAttribute a0[1];
        a0[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        a0[0].val.programmaticStreamSerializationAllowed = 1;
        c0.attrs = a0; c0.numAttrs = pdl ? 1 : 0;
        CU(cudaLaunchKernelEx(&c0, k_rank_tiles, o->d, f_all));
        g_launches++;
        pdl = false;                                              // k_rank itself follows in stream order
    }
    auto kern = k_rank<1024>;
    if (o->rank_smem > 48 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)o->rank_smem));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(o->d.RS, o->d.B); cfg.blockDim = dim3(o->rank_threads); cfg.dynamicSmemBytes = o->rank_smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    CU(cudaLaunchKernelEx(&cfg, kern, o->d, f_all, mode, payload));
    g_launches++;
    return 0;
}

template <int NVB, int RMAX, bool SMEM, bool OVERLAP = false, int WARPS = UPD_WARPS>
int launch_update_t(lmcma_b200_opt* o, const UpdateArgs& a, bool pdl, cudaStream_t st) {
    auto kern = k_update<NVB, RMAX, SMEM, OVERLAP, WARPS>;
    if (o->upd_smem > 48 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)o->upd_smem));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(o->d.B); cfg.blockDim = dim3(32 * WARPS); cfg.dynamicSmemBytes = o->upd_smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    CU(cudaLaunchKernelEx(&cfg, kern, o->d, a));
    g_launches++;
    return 0;
}

// the serial part of update() (k_update.cuh).  pdl: launched as a programmatic dependent of the kernel enqueued just
// before it on `st` (k_rank), so that its prologue overlaps that kernel
int launch_update(lmcma_b200_opt* o, const UpdateArgs& a_in, bool pdl, cudaStream_t st) {
    UpdateArgs a = a_in;
    if (a.sweep_warps <= 0) a.sweep_warps = o->upd_sweep_warps;     // callers that build UpdateArgs from scratch
    if (a.phase == 0) o->spec_valid = false;                        // a whole update: whatever the speculative pass left is stale
    if (o->upd_gram) {
        int rc = o->upd_nvb == 4 ? launch_update_t<4, -1, false>(o, a, pdl, st) : launch_update_t<16, -1, false>(o, a, pdl, st);
        if (rc) return rc;
        const OptDev& d = o->d;
        const int tiles = (d.m + GRAM_TILE - 1) / GRAM_TILE;
        k_gram<<<dim3(tiles, tiles, d.B * GRAM_KS), 256, 0, st>>>(d);
        const int coef_threads = std::min(1024, (8 * d.m + 31) & ~31);    // 8 lanes per basis row (m <= 128)
        auto coef = d.m <= 40 ? k_coef<5> : (d.m <= 80 ? k_coef<10> : (d.m <= 96 ? k_coef<12> : k_coef<16>));   // elements per lane: ceil(m / 8)
        if (o->coef_smem > 48 * 1024) CU(cudaFuncSetAttribute(coef, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)o->coef_smem));
        coef<<<d.B, coef_threads, o->coef_smem, st>>>(d);
        k_combine<<<dim3((d.ns / 4 + 127) / 128, (d.m + COMBINE_ROWS - 1) / COMBINE_ROWS, d.B), 128, 0, st>>>(d);
        g_launches += 3;
        CU(cudaGetLastError());
        return 0;
    }
    if (o->upd_nvb == 4) {
        if (a.overlap) {                                         // the overlapped generation (ensure_graph): register sweep only
            if (o->upd_rmax == 3) return launch_update_t<4, 3, true, true>(o, a, pdl, st);
            if (o->upd_rmax == 5) return launch_update_t<4, 5, true, true>(o, a, pdl, st);
            return fail(LMCMA_B200_ERR_STATE, "overlapped generation without the register sweep");
        }
        if (o->upd_warps == 8) {                                 // batched instances: two CTAs of 8 warps per SM (configure_update)
            if (o->upd_rmax == 3) return launch_update_t<4, 3, true, false, 8>(o, a, pdl, st);
            if (o->upd_rmax == 5) return launch_update_t<4, 5, true, false, 8>(o, a, pdl, st);
        }
        if (o->upd_rmax == 3) return launch_update_t<4, 3, true>(o, a, pdl, st);
        if (o->upd_rmax == 5) return launch_update_t<4, 5, true>(o, a, pdl, st);
        return o->upd_rows_in_smem ? launch_update_t<4, 0, true>(o, a, pdl, st) : launch_update_t<4, 0, false>(o, a, pdl, st);
    }
    return o->upd_rows_in_smem ? launch_update_t<16, 0, true>(o, a, pdl, st) : launch_update_t<16, 0, false>(o, a, pdl, st);
}

UpdateArgs update_args_local(lmcma_b200_opt* o) {
    UpdateArgs a;
    memset(&a, 0, sizeof(a));
    a.f_all = o->d.fit;
    a.slices = o->d.partial; a.n_slices = o->d.RS; a.slice_stride = o->d.ns; a.inst_stride = (long long)o->d.RS * o->d.ns;
    a.blocked = o->tune.update_blocked;
    a.sweep_warps = o->upd_sweep_warps;
    a.no_dry = o->tune.update_dry ? 0 : 1;
    a.spec = o->d_spec; a.spec_stride = (long long)o->spec_stride;
    return a;
}

int configure_update(lmcma_b200_opt* o) {
    const int nq = o->d.ns / 4;
    if (nq > 512) return fail(LMCMA_B200_ERR_ARG, "n = %d too large (max 2048)", o->d.n);
    o->upd_nvb = nq <= 128 ? 4 : 16;
    // k_rank: small populations (lambda <= 256) take CTAs of 256 threads and tiles of lambda values — the choice depends on
    // lambda alone, so that a batched instance and the same instance run alone sum their partial sums in the same order
    o->rank_threads = o->d.lambda <= 256 ? 256 : 1024;
    o->d.rank_ftile = o->d.lambda <= 256 ? ((o->d.lambda + 3) & ~3) : TELL_FTILE;
    o->rank_smem = (size_t)2 * o->d.rank_ftile * 4 + (size_t)3 * TELL_MAX_ROWS * 4 + (size_t)7 * 128 * 16;
    const size_t fixed = (((size_t)o->d.m * 36 + 8 + 127) & ~(size_t)127) + (size_t)2048 * (UPD_GROUPS - 1);
    // + the block Gram entries, and the stages of the newest row's multi-warp chain (k_update.cuh: chain_part)
    const size_t rows = (size_t)o->d.m * o->d.ns * sizeof(float) + (size_t)o->d.m * (UPD_BLK + 1) * sizeof(float) +
                        (o->upd_nvb == 4 ? (size_t)(2 * 4 * UPD_BLK * 32 + 16) * sizeof(float) + 16 : 0);
    const size_t budget = o->props->smem_optin - 2048;            // static shared + slack
    o->upd_rows_in_smem = fixed + rows <= budget;
    o->upd_smem = fixed + (o->upd_rows_in_smem ? rows : 0);
    if (o->upd_smem > budget) return fail(LMCMA_B200_ERR_ARG, "k_update needs %zu B shared memory (m = %d too large)", o->upd_smem, o->d.m);
    // rows that fit neither the registers nor the shared memory of one SM: Gram-matrix recompute (k_gram.cuh)
    o->coef_smem = (size_t)2 * o->d.m * (o->d.m | 1) * sizeof(double) + (size_t)2 * o->d.m * sizeof(double);
    o->upd_gram = (!o->upd_rows_in_smem || o->tune.update_gram) && o->d.m <= 128 && o->coef_smem <= budget && !o->tune.update_streaming;
    if (o->upd_gram) { o->upd_rows_in_smem = false; o->upd_smem = fixed; }
    // pending rows in registers when they fit: m <= 8 warps x RMAX rows of <= 128 float4 columns
    o->upd_rmax = 0;
    o->upd_warps = UPD_WARPS;
    if (o->upd_nvb == 4 && o->upd_rows_in_smem && !o->upd_gram && !o->tune.update_streaming) {
        // batched instances (more than one wave of one-CTA-per-SM launches: two 8-warp CTAs per SM take ~1.3x the time of one
        // 16-warp CTA, so a second wave is what they beat) whose rows fit 8 warps x 5: CTAs of 8 warps, two per
        // SM — the sweep of one instance is a chain of dependent steps that leaves its SM two thirds idle (k_update.cuh)
        const bool two_per_sm = 2 * o->upd_smem + 4096 <= o->props->smem_optin && o->d.m <= 8 * 5;
        if (o->tune.update_warps == 8 ? two_per_sm : (o->tune.update_warps == 0 && two_per_sm && o->d.B > o->props->sm_count)) o->upd_warps = 8;
        o->upd_sweep_warps = o->upd_warps;
        const int forced_sw = o->tune.update_sweep_warps;
        if (forced_sw >= 1 && forced_sw <= o->upd_warps) o->upd_sweep_warps = forced_sw;
        if (o->d.m <= o->upd_sweep_warps * 3) o->upd_rmax = 3;
        else if (o->d.m <= o->upd_sweep_warps * 5) o->upd_rmax = 5;
    }
    return 0;
}

int cost_args_for(lmcma_b200_opt* o, CostArgs* a) {
    if (!o->map) return fail(LMCMA_B200_ERR_STATE, "no cost attached (call lmcma_b200_attach_cost)");
    memset(a, 0, sizeof(*a));
    a->W = o->obj.waypoints; a->w_len = o->obj.w_len; a->w_clr = o->obj.w_clr; a->w_col = o->obj.w_col;
    a->X = o->d.X; a->ld = o->d.ns; a->inst_rows = o->d.pop_count;
    a->ends = o->d_ends; a->ends_per_instance = 1;
    a->f = o->d.fit; a->f_stride = o->d.lambda; a->f_offset = o->d.pop_offset;
    a->ncoll = o->d.ncoll; a->nsamp = o->d.nsamp;
    return 0;
}

// generate / stage deviates on the host for HANSEN mode, then sample
int host_rng_and_sample(lmcma_b200_opt* o, cudaStream_t st, bool progressive = false) {
    if (o->cfg.rng == LMCMA_B200_RNG_HANSEN) {
        const size_t rows = (size_t)o->d.pop_count;
        o->z_host.assign(rows * o->d.ns, 0.f);
        // the reference draws row by row, N deviates per offspring (lmcma.cpp:303-305, 212-215)
        for (size_t r = 0; r < rows; ++r)
            for (int k = 0; k < o->d.n; ++k) o->z_host[r * o->d.ns + k] = (float)o->hansen.gauss();
        CU(cudaMemcpyAsync(o->d.Z, o->z_host.data(), o->z_host.size() * sizeof(float), cudaMemcpyHostToDevice, st));
        CU(cudaStreamSynchronize(st));   // z_host is pageable and reused
    }
    return launch_sample(o, st, o->cfg.rng != LMCMA_B200_RNG_HANSEN, progressive);
}

// update() + sample() after the fitness of a generation is in d.fit
int generation_tail(lmcma_b200_opt* o, cudaStream_t st) {
    int rc;
    if ((rc = launch_rank(o, o->d.fit, RANK_PLAIN, nullptr, st))) return rc;
    // device RNG: k_sample follows k_update directly and can take its outputs as they become final
    const bool prog = o->progressive && o->cfg.rng == LMCMA_B200_RNG_PHILOX && !o->mirror_dirty;
    UpdateArgs ua = update_args_local(o);
    ua.progressive = prog ? 1 : 0;
    if ((rc = launch_update(o, ua, true, st))) return rc;
    o->x_cache_valid = false;
    o->sample_idx = 0;
    if (o->cfg.rng == LMCMA_B200_RNG_INJECT && !o->pending_z) { o->needs_sample = true; return 0; }
    o->pending_z = false;
    return host_rng_and_sample(o, st, prog);
}

// Do the two branches of a forked CUDA graph really run concurrently on this device, in this process?  (Not under a
// profiler / sanitizer that serialises kernels, not when someone else holds the SMs.)  Probed once per device with two
// one-thread kernels that shake hands within 5 ms; the second of two launches counts (the first pays module loading).
int probe_coschedule(lmcma_b200_opt* o) {
    if (o->props->cosched >= 0) return o->props->cosched;
    int result = 0;
    int* d = nullptr;
    cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
    cudaStream_t st = o->own_stream;
    if (cudaMalloc(&d, 4 * sizeof(int)) != cudaSuccess) { cudaGetLastError(); return 0; }
    bool ok = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    if (ok) {
        ok = cudaEventRecord(o->ev_fork, st) == cudaSuccess && cudaStreamWaitEvent(o->side_stream, o->ev_fork, 0) == cudaSuccess;
        if (ok) { k_probe<<<1, 1, 0, o->side_stream>>>(d, d + 2, 0, 5000000ll); ok = cudaEventRecord(o->ev_join, o->side_stream) == cudaSuccess; }
        if (ok) { k_probe<<<1, 1, 0, st>>>(d, d + 2, 1, 5000000ll); ok = cudaStreamWaitEvent(st, o->ev_join, 0) == cudaSuccess; }
        cudaError_t e = cudaStreamEndCapture(st, &graph);
        ok = ok && e == cudaSuccess && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
    }
    if (ok) {
        for (int attempt = 0; attempt < 2 && ok; ++attempt) {
            int h[4] = {0, 0, 0, 0};
            ok = cudaMemcpyAsync(d, h, sizeof(h), cudaMemcpyHostToDevice, st) == cudaSuccess && cudaGraphLaunch(exec, st) == cudaSuccess &&
                 cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, st) == cudaSuccess && cudaStreamSynchronize(st) == cudaSuccess;
            result = ok && h[2] == 1 && h[3] == 1;
        }
        g_launches += 4;
    }
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
    cudaFree(d);
    cudaGetLastError();
    o->props->cosched = result;
    return result;
}

// A kernel of the overlapped generation gave up waiting for its partner (OptDev::err): report it ONCE, and keep the
// handle usable on the linear PDL graph from here on.  The generation in flight is void: the optimiser state must be
// restored by the caller (set_* from a checkpoint) or the handle recreated.
}  // namespace
int lmcma_capi::check_lost(lmcma_b200_opt* o) {
    if (!o->err_host || *reinterpret_cast<volatile int*>(o->err_host) == 0) return 0;
    const int code = *reinterpret_cast<volatile int*>(o->err_host);
    *reinterpret_cast<volatile int*>(o->err_host) = 0;
    o->overlap = false;
    o->props->cosched = 0;
    if (o->graph_exec) { cudaGraphExecDestroy(o->graph_exec); o->graph_exec = nullptr; }
    if (o->graph_exec_multi) { cudaGraphExecDestroy(o->graph_exec_multi); o->graph_exec_multi = nullptr; }
    if (o->tell_graph) { cudaGraphExecDestroy(o->tell_graph); o->tell_graph = nullptr; }
    if (o->tell_graph_resume) { cudaGraphExecDestroy(o->tell_graph_resume); o->tell_graph_resume = nullptr; }
    o->spec_valid = false;
    o->tell_graph_failed = true;
    o->mirror_dirty = true;
    return fail(LMCMA_B200_ERR_CUDA, "overlapped generation lost co-scheduling (%s): the branches of the CUDA graph did not run "
                "concurrently (profiler / sanitizer / SMs held elsewhere).  This handle now uses the linear graph; the generation in "
                "flight is void - restore the state or recreate the handle (LMCMA_B200_OVERLAP=0 avoids the overlapped graph)",
                code == LOST_UPDATE_WAITING_FOR_RANK ? "k_update timed out waiting for k_rank" : "k_sample timed out waiting for k_update");
}
namespace {

int ensure_graph(lmcma_b200_opt* o) {
    if (o->graph_exec && o->graph_built_for == o->stream) return 0;
    if (o->graph_exec) { cudaGraphExecDestroy(o->graph_exec); o->graph_exec = nullptr; }
    if (o->graph_exec_multi) { cudaGraphExecDestroy(o->graph_exec_multi); o->graph_exec_multi = nullptr; }
    CostArgs ca;
    int rc = cost_args_for(o, &ca);
    if (rc) return rc;
    cudaStream_t st = o->stream;
    if ((rc = ensure_mirror(o, st))) return rc;
    const long long before = g_launches.load();
    if (o->tune.graph_dbg && !o->graph_dbg && cudaMalloc(&o->graph_dbg, 64 * sizeof(long long)) == cudaSuccess) { cudaMemset(o->graph_dbg, 0, 64 * sizeof(long long)); cudaDeviceSynchronize(); }
    // `reps` generations in ONE graph: between two graph launches the device idles for the launch latency of the next one
    // (~5 us of a 52 us generation); inside a graph a generation follows the previous one as a plain dependency
    auto capture = [&](int reps, cudaGraphExec_t* out) -> int {
        cudaGraph_t graph = nullptr;
        CU(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        o->mirror_suppressed = true;                             // fused generations keep the candidates on the device
        UpdateArgs ua = update_args_local(o);
        ua.progressive = o->progressive ? 1 : 0;                 // ensure_mirror above: the mirror is clean
        ua.dbg = o->graph_dbg;
        int rc2 = 0;
        for (int g = 0; g < reps && !rc2; ++g) {
            if (o->overlap) {
                // side branch: k_update starts with the generation; main branch: k_gate (waits until k_update holds its SM) ->
                // k_cost -> k_rank -> k_sample (released by k_rank, follows k_update's flags); join before the generation ends
                ua.overlap = 1;
                if (cudaEventRecord(o->ev_fork, st) != cudaSuccess || cudaStreamWaitEvent(o->side_stream, o->ev_fork, 0) != cudaSuccess) rc2 = fail(LMCMA_B200_ERR_CUDA, "graph fork failed");
                if (!rc2) rc2 = launch_update(o, ua, false, o->side_stream);
                if (!rc2 && cudaEventRecord(o->ev_join, o->side_stream) != cudaSuccess) rc2 = fail(LMCMA_B200_ERR_CUDA, "graph join record failed");
                if (!rc2) { k_gate<<<(o->d.B + 31) / 32, 32, 0, st>>>(o->d); g_launches++; }
                if (!rc2) rc2 = launch_cost(o->map->dev, ca, o->d.pop_count, o->d.B, o->cost_shape, false, st);
                if (!rc2) rc2 = launch_rank(o, o->d.fit, RANK_PLAIN | RANK_KEEP_FLAGS | (o->tune.rank_late ? 0 : 4), nullptr, st, true);
                if (!rc2) rc2 = launch_sample(o, st, true, true, 2);
                if (!rc2 && cudaStreamWaitEvent(st, o->ev_join, 0) != cudaSuccess) rc2 = fail(LMCMA_B200_ERR_CUDA, "graph join failed");
            } else {
                rc2 = launch_cost(o->map->dev, ca, o->d.pop_count, o->d.B, o->cost_shape, false, st);
                if (!rc2) rc2 = launch_rank(o, o->d.fit, RANK_PLAIN, nullptr, st, true);
                if (!rc2) rc2 = launch_update(o, ua, true, st);
                if (!rc2) rc2 = launch_sample(o, st, true, o->progressive);
            }
        }
        cudaError_t e = cudaStreamEndCapture(st, &graph);
        o->mirror_suppressed = false;
        if (!rc2 && e == cudaSuccess) {
            e = cudaGraphInstantiate(out, graph, 0);
            if (e != cudaSuccess) *out = nullptr;
        }
        if (graph) cudaGraphDestroy(graph);
        if (!rc2 && e != cudaSuccess) rc2 = fail(LMCMA_B200_ERR_CUDA, "graph capture / instantiate failed: %s", cudaGetErrorString(e));
        return rc2;
    };
    rc = capture(1, &o->graph_exec);
    if (!rc && o->tune.graph_unroll > 1 && capture(o->tune.graph_unroll, &o->graph_exec_multi) != 0) {   // optional: the single one still works
        cudaGetLastError();
        o->graph_exec_multi = nullptr;
    }
    g_launches.store(before);   // capture enqueues nothing
    if (rc && o->overlap) {
        // the forked graph could not be built here (capture / instantiation of the side branch): fall back to the linear one
        cudaGetLastError();
        if (o->graph_exec) { cudaGraphExecDestroy(o->graph_exec); o->graph_exec = nullptr; }
    if (o->graph_exec_multi) { cudaGraphExecDestroy(o->graph_exec_multi); o->graph_exec_multi = nullptr; }
        o->overlap = false;
        return ensure_graph(o);
    }
    if (rc) return rc;
    o->graph_built_for = st;
    return 0;
}

// tell_all of one query as a forked graph (see lmcma_b200_tell_all); on failure the caller takes the stream-ordered path
int ensure_tell_graph(lmcma_b200_opt* o) {
    if (o->tell_graph && o->tell_graph_for == o->stream) return ensure_mirror(o, o->stream);
    if (o->tell_graph) { cudaGraphExecDestroy(o->tell_graph); o->tell_graph = nullptr; }
    if (o->tell_graph_resume) { cudaGraphExecDestroy(o->tell_graph_resume); o->tell_graph_resume = nullptr; }
    o->spec_valid = false;
    cudaStream_t st = o->stream;
    const OptDev& d = o->d;
    int rc = ensure_mirror(o, st);
    if (rc) return rc;
    if (!o->f_pinned && cudaHostAlloc(&o->f_pinned, (size_t)d.B * d.lambda * sizeof(float), cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError(); o->f_pinned = nullptr; o->tell_graph_failed = true; return 0;
    }
    const long long before = g_launches.load();
    // the speculative pass pays only where something hides it: with the host mirror on, the candidates' trip across PCIe
    // (35 us behind the sampler).  Without the mirror tell_all returns when the sampler is done, and a 25 us kernel behind
    // k_update would be the longer branch of the graph (measured: tell_all 57 -> 69 us)
    const bool spec = o->d_spec != nullptr && o->mirror_on && !o->mirror_suppressed;
    o->tell_graph_spec = spec;
    if (o->tune.graph_dbg && !o->graph_dbg && cudaMalloc(&o->graph_dbg, 64 * sizeof(long long)) == cudaSuccess) { cudaMemset(o->graph_dbg, 0, 64 * sizeof(long long)); cudaDeviceSynchronize(); }
    // resume = 0: the whole update; resume = 1: k_update picks up the rows the last graph's speculative pass left.  With the
    // scratch both graphs END with that pass for the next generation, on the side branch behind k_update
    auto capture = [&](int resume, cudaGraphExec_t* out) -> int {
        cudaGraph_t graph = nullptr;
        CU(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        UpdateArgs ua = update_args_local(o);
        ua.progressive = 1;
        ua.overlap = 1;
        ua.phase = resume ? 2 : 0;
        ua.dbg = o->graph_dbg;                                   // LMCMA_B200_GRAPH_DBG: the timeline lmcma_b200_sync prints
        const bool valid_before = o->spec_valid;
        bool ok = cudaEventRecord(o->ev_fork, st) == cudaSuccess && cudaStreamWaitEvent(o->side_stream, o->ev_fork, 0) == cudaSuccess;
        if (ok) ok = launch_update(o, ua, false, o->side_stream) == 0;
        if (ok && spec) { UpdateArgs us = ua; us.phase = 1; us.dbg = nullptr; ok = launch_update(o, us, false, o->side_stream) == 0; }
        o->spec_valid = valid_before;                            // capture enqueues nothing
        if (ok) ok = cudaEventRecord(o->ev_join, o->side_stream) == cudaSuccess;
        if (ok) ok = cudaMemcpyAsync(d.fit, o->f_pinned, (size_t)d.B * d.lambda * sizeof(float), cudaMemcpyHostToDevice, st) == cudaSuccess;
        if (ok) { k_gate<<<(d.B + 31) / 32, 32, 0, st>>>(d); ok = cudaGetLastError() == cudaSuccess; }   // k_update has reset the flags and holds its SM
        // the sampler is released when k_rank starts: the pairs it needs first are final (resume: at once; else as the sweep goes)
        if (ok) ok = launch_rank(o, d.fit, RANK_PLAIN | RANK_KEEP_FLAGS | (o->tune.rank_late ? 0 : 4), nullptr, st, false) == 0;
        if (ok) ok = launch_sample(o, st, true, true, 2) == 0;
        if (ok) ok = cudaStreamWaitEvent(st, o->ev_join, 0) == cudaSuccess;
        cudaError_t e = cudaStreamEndCapture(st, &graph);
        if (ok && e == cudaSuccess && cudaGraphInstantiate(out, graph, 0) != cudaSuccess) *out = nullptr;
        if (graph) cudaGraphDestroy(graph);
        return 0;
    };
    rc = capture(0, &o->tell_graph);
    if (!rc && o->tell_graph && spec) rc = capture(1, &o->tell_graph_resume);
    g_launches.store(before);   // capture enqueues nothing
    if (rc) return rc;
    if (!o->tell_graph || (spec && !o->tell_graph_resume)) {
        cudaGetLastError();
        if (o->tell_graph) { cudaGraphExecDestroy(o->tell_graph); o->tell_graph = nullptr; }
        o->tell_graph_failed = true; return 0;
    }
    o->tell_graph_for = st;
    return 0;
}

}  // namespace

extern "C" {

int lmcma_b200_abi_version(void) { return LMCMA_B200_ABI_VERSION; }
const char* lmcma_b200_last_error(void) { return g_err.c_str(); }
int64_t lmcma_b200_launch_count(void) { return g_launches.load(); }

int lmcma_b200_device_count(int* count_out) {
    ARG(count_out, "null output");
    CU(cudaGetDeviceCount(count_out));
    return 0;
}
int lmcma_b200_device_info(int device, int* sm_count, int64_t* l2_bytes, int64_t* hbm_bytes, int* cc) {
    cudaDeviceProp dp;
    CU(cudaGetDeviceProperties(&dp, device));
    if (sm_count) *sm_count = dp.multiProcessorCount;
    if (l2_bytes) *l2_bytes = dp.l2CacheSize;
    if (hbm_bytes) *hbm_bytes = (int64_t)dp.totalGlobalMem;
    if (cc) *cc = dp.major * 10 + dp.minor;
    return 0;
}

// =================================================================================================
// optimiser
// =================================================================================================
int lmcma_b200_create(const lmcma_b200_config* cfg, const double* x0, const double* lo, const double* hi,
                      lmcma_b200_opt** out) {
    return lmcma_b200_create_with_prior(cfg, x0, lo, hi, nullptr, out);
}

static int create_body(lmcma_b200_opt* o, DeviceProps* props, const lmcma_b200_config* cfg, const double* x0, const double* lo,
                       const double* hi, const double* covariance);

int lmcma_b200_create_with_prior(const lmcma_b200_config* cfg, const double* x0, const double* lo, const double* hi,
                                 const double* covariance, lmcma_b200_opt** out) {
    ARG(cfg && out, "null pointer");
    ARG(cfg->n >= 1, "n must be >= 1");
    ARG(cfg->batch >= 1, "batch must be >= 1");
    ARG(cfg->rng >= 0 && cfg->rng <= 2, "unknown rng mode");
    ARG(cfg->sigma0 > 0.0, "sigma0 must be > 0");
    ARG(x0 || cfg->rng == LMCMA_B200_RNG_HANSEN, "x0 == NULL needs the HANSEN rng (uniform start, lmcma.cpp:161-163)");
    ARG(cfg->rng != LMCMA_B200_RNG_HANSEN || cfg->batch == 1, "HANSEN rng is a single serial stream: batch must be 1");
    DeviceProps* props;
    int rc = query_props(cfg->device, &props);
    if (rc) return rc;
    CU(cudaSetDevice(cfg->device));

    lmcma_b200_opt* o = new lmcma_b200_opt();
    rc = create_body(o, props, cfg, x0, lo, hi, covariance);   // any failure inside: everything allocated so far is released here
    if (rc) { const std::string keep = g_err; lmcma_b200_destroy(o); g_err = keep; return rc; }
    *out = o;
    return 0;
}

static int create_body(lmcma_b200_opt* o, DeviceProps* props, const lmcma_b200_config* cfg, const double* x0, const double* lo,
                       const double* hi, const double* covariance) {
    int rc = 0;
    o->tune = Tuning::from_env();
    o->cfg = *cfg;
    o->props = props;
    OptDev& d = o->d;
    d.n = cfg->n;
    d.ns = (cfg->n + 3) & ~3;
    d.lambda = cfg->lambda < 1 ? 4 + int(3 * std::log((double)cfg->n)) : cfg->lambda;   // lmcma.cpp:134-135
    d.mu = d.lambda / 2;                                                                  // lmcma.cpp:136
    d.m = cfg->m < 1 ? d.lambda : cfg->m;                                                 // lmcma.cpp:266
    d.B = cfg->batch;
    if (d.lambda < 2 || d.mu < 1 || d.m < 2) { return fail(LMCMA_B200_ERR_ARG, "need lambda >= 2 and m >= 2"); }
    d.pop_offset = cfg->pop_count < 1 ? 0 : cfg->pop_offset;
    d.pop_count = cfg->pop_count < 1 ? d.lambda : cfg->pop_count;
    if (d.pop_offset < 0 || d.pop_offset + d.pop_count > d.lambda) { return fail(LMCMA_B200_ERR_ARG, "population slice out of range"); }
    if (cfg->rng == LMCMA_B200_RNG_HANSEN && d.pop_count != d.lambda) { return fail(LMCMA_B200_ERR_ARG, "HANSEN rng cannot be split"); }
    d.rng_mode = cfg->rng;
    d.record_z = cfg->record_z;
    d.seed = (unsigned long long)cfg->seed;
    o->hansen.reseed(cfg->seed);

    // constants (lmcma.cpp:144-156, 238, 268-272)
    o->weights.resize(d.mu);
    double sw = 0;
    for (int i = 0; i < d.mu; ++i) { o->weights[i] = std::log(double(d.mu) + 0.5) - std::log(double(1 + i)); sw += o->weights[i]; }
    double mueff = 0;
    for (int i = 0; i < d.mu; ++i) { o->weights[i] /= sw; mueff += o->weights[i] * o->weights[i]; }
    d.mueff = 1.0 / mueff;
    d.c1 = 1.0 / (10 * std::log((double)(d.n + 1)));
    d.cc = 1.0 / d.m;
    d.cs = 0.3; d.target = 0.25;
    d.K = 1 / std::sqrt(1 - d.c1);
    d.M = std::sqrt(1 - d.c1);
    d.pc_coef = std::sqrt(d.cc * (2 - d.cc) * d.mueff);

    CU(cudaStreamCreateWithFlags(&o->own_stream, cudaStreamNonBlocking));
    o->stream = o->own_stream;
    CU(cudaEventCreate(&o->ev0));
    CU(cudaEventCreate(&o->ev1));

    const size_t B = d.B, ns = d.ns, lam = d.lambda, pc = d.pop_count, m = d.m;
    DM(d.X, B * pc * ns);
    DM(d.D, B * pc * ns);
    if (cfg->rng != LMCMA_B200_RNG_PHILOX || cfg->record_z || covariance) DM(d.Z, B * pc * ns);
    if (covariance) {   // CMABase::init factors the prior once (cholesky, lmcma.cpp:165-169)
        std::vector<double> Ld((size_t)d.n * d.n);
        if (!cholesky_lower(covariance, d.n, Ld.data())) { return fail(LMCMA_B200_ERR_ARG, "covariance prior is not symmetric positive definite"); }
        std::vector<float> Lf((size_t)d.n * ns, 0.f);
        for (int i = 0; i < d.n; ++i)
            for (int k = 0; k <= i; ++k) Lf[(size_t)i * ns + k] = (float)Ld[(size_t)i * d.n + k];
        DM(o->d_Lf, (size_t)d.n * ns);
        CU(cudaMemcpy(o->d_Lf, Lf.data(), Lf.size() * sizeof(float), cudaMemcpyHostToDevice));
        d.Lf = o->d_Lf;
        DM(d.Zc, B * pc * ns);
    }
    DM(d.fit, B * lam); DM(d.fit_sorted, B * lam); DM(d.prev_fit, B * lam);
    DM(d.rank, B * lam); DM(d.arindex, B * lam);
    if (d.lambda > TELL_FTILE && d.pop_count == d.lambda && o->tune.rank_sorted) {
        DM(d.prev_sorted, B * lam); DM(d.tile_sorted, B * lam); DM(d.tile_pos, B * lam);
    }
    DM(d.ncoll, B * pc); DM(d.nsamp, B * pc);
    DM(d.xmean, B * ns); DM(d.pc, B * ns);
    DM(d.V, B * m * ns); DM(d.P, B * m * ns);
    DM(d.Nj, B * m); DM(d.Lj, B * m); DM(d.Njf, B * m); DM(d.Njs, B * m);
    DM(d.VPs, B * m * 2 * ns);
    DM(d.t, B * m); DM(d.vec, B * m);
    DM(d.sc, B); DM(d.best_x, B * ns); DM(d.S_count, B); DM(d.done_count, B); DM(d.progress, B * (m + 2)); DM(d.rank_ticket, B); DM(d.resident, B);
    {   // row slices of k_tell's phase A: 32 rows per slice, at most 256 slices
        const int rows_per = std::max(32, (d.pop_count + 255) / 256);
        if (rows_per > TELL_MAX_ROWS) { return fail(LMCMA_B200_ERR_ARG, "population too large (max %d rows per handle)", 256 * TELL_MAX_ROWS); }
        d.RS = (d.pop_count + rows_per - 1) / rows_per;
    }
    if (o->tune.dbg) DM(d.dbg, 64);
    DM(d.partial, B * d.RS * ns);

    std::vector<float> wf(o->weights.begin(), o->weights.end());
    DM(o->d_w, d.mu);
    CU(cudaMemcpy(o->d_w, wf.data(), d.mu * sizeof(float), cudaMemcpyHostToDevice));
    d.w = o->d_w;
    if (lo) {   // padded to the row stride so that the kernels clamp whole float4 columns
        o->lo_f.assign(d.ns, -std::numeric_limits<float>::max());
        for (int k = 0; k < d.n; ++k) o->lo_f[k] = (float)lo[k];
        DM(o->d_lo, d.ns);
        CU(cudaMemcpy(o->d_lo, o->lo_f.data(), d.ns * sizeof(float), cudaMemcpyHostToDevice));
        d.lo = o->d_lo;
    }
    if (hi) {
        o->hi_f.assign(d.ns, std::numeric_limits<float>::max());
        for (int k = 0; k < d.n; ++k) o->hi_f[k] = (float)hi[k];
        DM(o->d_hi, d.ns);
        CU(cudaMemcpy(o->d_hi, o->hi_f.data(), d.ns * sizeof(float), cudaMemcpyHostToDevice));
        d.hi = o->d_hi;
    }
    // initial mean (lmcma.cpp:158-163)
    std::vector<double> xm(B * ns, 0.0);
    for (size_t b = 0; b < B; ++b)
        for (int k = 0; k < d.n; ++k) xm[b * ns + k] = x0 ? x0[b * d.n + k] : o->hansen.uniform();
    CU(cudaMemcpy(d.xmean, xm.data(), xm.size() * sizeof(double), cudaMemcpyHostToDevice));
    std::vector<Scalars> sc(B);
    for (auto& s : sc) {
        memset(&s, 0, sizeof(s));
        s.sigma = cfg->sigma0; s.s = 0.0; s.best_f = std::numeric_limits<double>::max();
    }
    CU(cudaMemcpy(d.sc, sc.data(), B * sizeof(Scalars), cudaMemcpyHostToDevice));

    rc = configure_sample(o);
    if (!rc) rc = configure_update(o);
    // one query with a wide sampler and the register sweep: k_sample consumes the direction pairs while k_update's sweep
    // is still producing them (k_update.cuh / k_sample.cuh "progressive")
    o->progressive = !rc && B == 1 && o->smp_wide && o->upd_rmax > 0 && !o->upd_gram && !o->d_Lf && o->smp_kc == 8 &&
                     o->smp_stages >= (m + o->smp_kc - 1) / o->smp_kc && cfg->rng == LMCMA_B200_RNG_PHILOX &&
                     o->tune.progressive != 0;
    o->overlap = o->progressive && o->tune.overlap != 0;
    if (o->overlap) {
        cudaError_t e2 = cudaStreamCreateWithFlags(&o->side_stream, cudaStreamNonBlocking);
        if (e2 == cudaSuccess) e2 = cudaEventCreateWithFlags(&o->ev_fork, cudaEventDisableTiming);
        if (e2 == cudaSuccess) e2 = cudaEventCreateWithFlags(&o->ev_join, cudaEventDisableTiming);
        if (e2 == cudaSuccess) e2 = cudaHostAlloc(&o->err_host, sizeof(int), cudaHostAllocMapped);
        if (e2 == cudaSuccess) { *o->err_host = 0; e2 = cudaHostGetDevicePointer(&d.err, o->err_host, 0); }
        if (e2 != cudaSuccess) { cudaGetLastError(); o->overlap = false; d.err = nullptr; }
        // the overlapped generation needs the two branches of a forked graph to run CONCURRENTLY: enabled only where a probe
        // has seen that happen (LMCMA_B200_OVERLAP=2 skips the probe)
        if (o->overlap && o->tune.overlap != 2 && probe_coschedule(o) != 1) o->overlap = false;
        if (o->overlap && o->tune.tell_spec != 0) {              // scratch of the speculative update (best effort)
            o->spec_stride = (spec_floats(m, d.ns) + 3) & ~(size_t)3;
            if (cudaMalloc(&o->d_spec, B * o->spec_stride * sizeof(float)) != cudaSuccess) { cudaGetLastError(); o->d_spec = nullptr; }
            else cudaMemset(o->d_spec, 0xff, B * o->spec_stride * sizeof(float));   // generation -1: matches nothing
        }
    }
    if (!rc && o->upd_gram) {
        rc = dmalloc(&d.G, B * GRAM_KS * m * m);
        if (!rc) rc = dmalloc(&d.Cf, B * m * m);
        if (!rc) rc = dmalloc(&d.gram_hdr, B);
    }
    if (rc) return rc;
    o->f_host.assign(B * lam, 0.f);
    // first population (LMCMA::init -> sample(), lmcma.cpp:298)
    if (cfg->rng == LMCMA_B200_RNG_INJECT) o->needs_sample = true;
    else {
        rc = host_rng_and_sample(o, o->stream);
        if (!rc) { cudaError_t e = cudaStreamSynchronize(o->stream); if (e != cudaSuccess) rc = fail(LMCMA_B200_ERR_CUDA, "first sample: %s", cudaGetErrorString(e)); }
        if (rc) return rc;
    }
    return 0;
}

int lmcma_b200_destroy(lmcma_b200_opt* o) {
    if (!o) return 0;
    cudaSetDevice(o->cfg.device);
    if (o->stream) cudaStreamSynchronize(o->stream);
    OptDev& d = o->d;
    void* ptrs[] = {d.prev_sorted, d.tile_sorted, d.tile_pos, d.X, d.D, d.Z, d.Zc, o->d_Lf, d.fit, d.fit_sorted, d.prev_fit, d.rank, d.arindex, d.ncoll, d.nsamp, d.xmean, d.pc, d.V, d.P,
                    d.Nj, d.Lj, d.Njf, d.Njs, d.VPs, d.dbg, d.G, d.Cf, d.gram_hdr, d.t, d.vec, d.sc, d.best_x, d.S_count, d.done_count, d.progress, d.rank_ticket, d.resident, d.partial, o->d_lo, o->d_hi, o->d_w, o->d_ends};
    for (void* p : ptrs) cudaFree(p);
    if (o->graph_exec) cudaGraphExecDestroy(o->graph_exec);
    if (o->graph_exec_multi) cudaGraphExecDestroy(o->graph_exec_multi);
    if (o->tell_graph) cudaGraphExecDestroy(o->tell_graph);
    if (o->tell_graph_resume) cudaGraphExecDestroy(o->tell_graph_resume);
    cudaFree(o->d_spec);
    if (o->f_pinned) cudaFreeHost(o->f_pinned);
    if (o->err_host) cudaFreeHost(o->err_host);
    if (o->x_mirror) { unregister_mirror(o->x_mirror); cudaFreeHost(o->x_mirror); }
    if (o->ev0) cudaEventDestroy(o->ev0);
    if (o->ev1) cudaEventDestroy(o->ev1);
    if (o->own_stream) cudaStreamDestroy(o->own_stream);
    if (o->side_stream) cudaStreamDestroy(o->side_stream);
    if (o->ev_fork) cudaEventDestroy(o->ev_fork);
    if (o->ev_join) cudaEventDestroy(o->ev_join);
    if (o->graph_dbg) cudaFree(o->graph_dbg);
    delete o;
    return 0;
}

int lmcma_b200_set_stream(lmcma_b200_opt* o, void* cuda_stream) {
    ARG(o, "null handle");
    CU(cudaStreamSynchronize(o->stream));
    o->stream = cuda_stream ? (cudaStream_t)cuda_stream : o->own_stream;
    return 0;
}

int lmcma_b200_shape(const lmcma_b200_opt* o, int32_t* out8) {
    ARG(o && out8, "null pointer");
    const OptDev& d = o->d;
    out8[0] = d.n; out8[1] = d.lambda; out8[2] = d.mu; out8[3] = d.m; out8[4] = d.B;
    out8[5] = d.pop_offset; out8[6] = d.pop_count; out8[7] = d.ns;
    return 0;
}

int lmcma_b200_inject_z(lmcma_b200_opt* o, const float* Z) {
    ARG(o && Z, "null pointer");
    if (o->cfg.rng != LMCMA_B200_RNG_INJECT) return fail(LMCMA_B200_ERR_STATE, "inject_z needs rng == INJECT");
    CU(cudaSetDevice(o->cfg.device));
    const OptDev& d = o->d;
    int rc = h2d_rows(d.Z, Z, (size_t)d.B * d.pop_count, d.n * sizeof(float), d.ns * sizeof(float), o->stream);
    if (rc) return rc;
    if (o->needs_sample) {
        o->needs_sample = false;
        rc = launch_sample(o, o->stream);
        if (rc) return rc;
        CU(cudaStreamSynchronize(o->stream));
    } else {
        o->pending_z = true;
    }
    return 0;
}

int lmcma_b200_resample(lmcma_b200_opt* o) {
    ARG(o, "null handle");
    if (o->needs_sample) return fail(LMCMA_B200_ERR_STATE, "no deviates yet: inject_z first");
    if (o->cfg.rng == LMCMA_B200_RNG_HANSEN) return fail(LMCMA_B200_ERR_STATE, "resample would advance the serial HANSEN stream");
    CU(cudaSetDevice(o->cfg.device));
    int rc = launch_sample(o, o->stream);
    if (rc) return rc;
    CU(cudaStreamSynchronize(o->stream));
    o->x_cache_valid = false;
    return 0;
}

int lmcma_b200_ask_all(lmcma_b200_opt* o, float* X) {
    ARG(o && X, "null pointer");
    if (o->needs_sample) return fail(LMCMA_B200_ERR_STATE, "no population yet: inject_z first");
    CU(cudaSetDevice(o->cfg.device));
    const OptDev& d = o->d;
    int rc = d2h_rows(X, d.X, (size_t)d.B * d.pop_count, d.n * sizeof(float), d.ns * sizeof(float), o->stream);
    return rc ? rc : check_lost(o);
}

int lmcma_b200_ask_all_view(lmcma_b200_opt* o, const float** X_view, int64_t* ld_out) {
    ARG(o && X_view, "null pointer");
    if (o->needs_sample) return fail(LMCMA_B200_ERR_STATE, "no population yet: inject_z first");
    CU(cudaSetDevice(o->cfg.device));
    OptDev& d = o->d;
    const size_t floats = (size_t)d.B * d.pop_count * d.ns;
    if (!o->x_mirror) {
        CU(cudaStreamSynchronize(o->stream));
        CU(cudaHostAlloc(&o->x_mirror, floats * sizeof(float), cudaHostAllocMapped));
        cudaError_t e = cudaHostGetDevicePointer(&d.Xh, o->x_mirror, 0);
        if (e != cudaSuccess) { cudaFreeHost(o->x_mirror); o->x_mirror = nullptr; d.Xh = nullptr; return fail(LMCMA_B200_ERR_CUDA, "cudaHostGetDevicePointer: %s", cudaGetErrorString(e)); }
        o->mirror_on = true;
        o->xh_fresh = false;
        register_mirror(o->x_mirror, floats * sizeof(float), d.X, d.ns, o->cfg.device);
        // the captured graphs carry the old kernel parameters (OptDev by value): rebuild them on next use
        if (o->tell_graph) { cudaGraphExecDestroy(o->tell_graph); o->tell_graph = nullptr; }
        if (o->tell_graph_resume) { cudaGraphExecDestroy(o->tell_graph_resume); o->tell_graph_resume = nullptr; }
        o->spec_valid = false;
    }
    if (!o->xh_fresh) {                                  // first use / after a fused run or a state setter: one ordinary copy
        CU(cudaMemcpyAsync(o->x_mirror, d.X, floats * sizeof(float), cudaMemcpyDeviceToHost, o->stream));
        o->xh_fresh = true;
    }
    CU(cudaStreamSynchronize(o->stream));
    *X_view = o->x_mirror;
    if (ld_out) *ld_out = d.ns;
    return check_lost(o);
}

int lmcma_b200_tell_all(lmcma_b200_opt* o, const float* f) {
    ARG(o && f, "null pointer");
    if (o->needs_sample) return fail(LMCMA_B200_ERR_STATE, "no population yet: inject_z first");
    if (o->d.pop_count != o->d.lambda) return fail(LMCMA_B200_ERR_STATE, "split-population handles use the mg_* entry points");
    CU(cudaSetDevice(o->cfg.device));
    const OptDev& d = o->d;
    int rc;
    if (o->overlap && !o->tell_graph_failed && o->tune.tell_overlap != 0) {
        // One query (the conditions of the overlapped generation): everything in update() that does not
        // depend on this generation's fitness — slot bookkeeping, the sweep over every pending row but the newest — runs on
        // a side branch while the fitness crosses PCIe and k_rank runs; k_update then waits for k_rank's tickets and
        // k_sample follows its hand-over flags.  Same kernels and arithmetic as the fused generation (bit-identical to the
        // serial order), replayed as one CUDA graph: launched kernel by kernel, the fork / join costs more host time than
        // the overlap saves.
        if ((rc = ensure_tell_graph(o)) == 0 && o->tell_graph) {
            memcpy(o->f_pinned, f, (size_t)d.B * d.lambda * sizeof(float));
            const bool resume = o->spec_valid && o->tell_graph_resume;   // the last tell_all graph left the next update's rows behind
            CU(cudaGraphLaunch(resume ? o->tell_graph_resume : o->tell_graph, o->stream));
            o->spec_valid = o->tell_graph_spec;                   // ... and so does this one
            g_launches += 4 + (o->d.tile_sorted ? 1 : 0) + (o->tell_graph_spec ? 1 : 0);   // k_update, k_gate, (k_rank_tiles,) k_rank, k_sample, (speculative k_update)
            o->xh_fresh = o->mirror_on;                          // the captured sampler writes the host mirror when it is on
            o->x_cache_valid = false;
            o->sample_idx = 0;
            o->pending_z = false;
            CU(cudaStreamSynchronize(o->stream));
            return check_lost(o);
        }
        if (rc) return rc;
    }
    CU(cudaMemcpyAsync(d.fit, f, (size_t)d.B * d.lambda * sizeof(float), cudaMemcpyHostToDevice, o->stream));
    rc = generation_tail(o, o->stream);
    if (rc) return rc;
    CU(cudaStreamSynchronize(o->stream));
    return 0;
}

int lmcma_b200_ask_one(lmcma_b200_opt* o, double* params, int32_t n) {
    ARG(o && params, "null pointer");
    ARG(n == o->d.n, "N mismatch");
    if (o->d.B != 1 || o->d.pop_count != o->d.lambda) return fail(LMCMA_B200_ERR_STATE, "ask_one needs batch == 1 and an unsplit population");
    if (!o->x_cache_valid) {
        o->x_cache.resize((size_t)o->d.lambda * o->d.n);
        int rc = lmcma_b200_ask_all(o, o->x_cache.data());
        if (rc) return rc;
        o->x_cache_valid = true;
    }
    for (int k = 0; k < n; ++k) params[k] = (double)o->x_cache[(size_t)o->sample_idx * n + k];
    return 0;
}

int lmcma_b200_tell_one(lmcma_b200_opt* o, const double* feedbacks, int32_t num) {
    ARG(o && (feedbacks || num == 0) && num >= 0, "null pointer");
    if (o->d.B != 1 || o->d.pop_count != o->d.lambda) return fail(LMCMA_B200_ERR_STATE, "tell_one needs batch == 1 and an unsplit population");
    double f = 0.0;                                    // lmcma.cpp:186-188
    for (int i = 0; i < num; ++i) f += feedbacks[i];
    // the device ranks FP32 values: keep huge penalties finite and ordered instead of letting them round to +-inf
    const double fmax = (double)std::numeric_limits<float>::max();
    o->f_host[o->sample_idx] = (float)(f > fmax ? fmax : (f < -fmax ? -fmax : f));
    if (++o->sample_idx % o->d.lambda == 0) {          // lmcma.cpp:199-204
        o->sample_idx = 0;
        return lmcma_b200_tell_all(o, o->f_host.data());
    }
    return 0;
}

int lmcma_b200_is_done(lmcma_b200_opt* o, int32_t* done) {
    ARG(o && done, "null pointer");
    CU(cudaSetDevice(o->cfg.device));
    std::vector<Scalars> sc(o->d.B);
    CU(cudaMemcpyAsync(sc.data(), o->d.sc, sc.size() * sizeof(Scalars), cudaMemcpyDeviceToHost, o->stream));
    CU(cudaStreamSynchronize(o->stream));
    for (int b = 0; b < o->d.B; ++b) done[b] = sc[b].sigma < 1e-20 ? 1 : 0;   // lmcma.cpp:426-429
    return 0;
}

int lmcma_b200_attach_cost(lmcma_b200_opt* o, lmcma_b200_map* map, const lmcma_b200_objective* obj,
                           const lmcma_b200_endpoints* ends) {
    ARG(o && ends, "null pointer");
    int rc = check_obj(map, obj);
    if (rc) return rc;
    ARG(map->device == o->cfg.device, "map and optimiser live on different devices");
    ARG(map->dev.dims * obj->waypoints == o->d.n, "n != dims * waypoints");
    CU(cudaSetDevice(o->cfg.device));
    o->map = map; o->obj = *obj;
    std::vector<float> e6((size_t)o->d.B * 6);
    for (int b = 0; b < o->d.B; ++b)
        for (int c = 0; c < 3; ++c) { e6[b * 6 + c] = ends[b].start[c]; e6[b * 6 + 3 + c] = ends[b].goal[c]; }
    if (!o->d_ends) DM(o->d_ends, (size_t)o->d.B * 6);
    CU(cudaMemcpy(o->d_ends, e6.data(), e6.size() * sizeof(float), cudaMemcpyHostToDevice));
    o->cost_shape = pick_cost_shape(obj->waypoints, ends[0].start, ends[0].goal, map->dev.dims, o->tune);
    if (o->graph_exec) { cudaGraphExecDestroy(o->graph_exec); o->graph_exec = nullptr; }
    if (o->graph_exec_multi) { cudaGraphExecDestroy(o->graph_exec_multi); o->graph_exec_multi = nullptr; }
    return 0;
}

int lmcma_b200_run(lmcma_b200_opt* o, int32_t generations) {
    ARG(o && generations >= 0, "bad argument");
    if (!o->map) return fail(LMCMA_B200_ERR_STATE, "no cost attached");
    if (o->needs_sample) return fail(LMCMA_B200_ERR_STATE, "no population yet: inject_z first");
    if (o->d.pop_count != o->d.lambda) return fail(LMCMA_B200_ERR_STATE, "split-population handles use the mg_* entry points");
    CU(cudaSetDevice(o->cfg.device));
    int rc = apply_l2_window(o->map, o->stream);
    if (rc) return rc;
    o->spec_valid = false;                                       // the fused generations advance the state past what tell_all's speculative pass saw
    if (o->cfg.rng == LMCMA_B200_RNG_PHILOX) {
        if ((rc = ensure_graph(o))) return rc;
        if ((rc = ensure_mirror(o, o->stream))) return rc;       // a state setter since the last run: the graph's sampler reads the mirror
        CU(cudaEventRecord(o->ev0, o->stream));                  // after the (host-side) graph build: last_run_ms is device time
        const int unroll = o->graph_exec_multi ? o->tune.graph_unroll : 0;
        for (int g = 0; g < generations;) {
            const bool multi = unroll > 1 && generations - g >= unroll;
            CU(cudaGraphLaunch(multi ? o->graph_exec_multi : o->graph_exec, o->stream));
            const int done = multi ? unroll : 1;
            g_launches += (long long)done * (4 + (o->d_Lf ? (o->cfg.rng == LMCMA_B200_RNG_PHILOX ? 2 : 1) : 0) + (o->upd_gram ? 3 : 0) + (o->overlap ? 1 : 0) + (o->d.tile_sorted ? 1 : 0));
            g += done;
        }
    } else {
        if (o->cfg.rng == LMCMA_B200_RNG_INJECT && generations > 1)
            return fail(LMCMA_B200_ERR_STATE, "INJECT rng: run one generation per injected Z");
        CostArgs ca;
        if ((rc = cost_args_for(o, &ca))) return rc;
        CU(cudaEventRecord(o->ev0, o->stream));
        for (int g = 0; g < generations; ++g) {
            if ((rc = launch_cost(o->map->dev, ca, o->d.pop_count, o->d.B, o->cost_shape, false, o->stream))) return rc;
            if ((rc = generation_tail(o, o->stream))) return rc;
        }
    }
    CU(cudaEventRecord(o->ev1, o->stream));
    o->have_run_timing = true;
    o->x_cache_valid = false;
    if (o->cfg.rng == LMCMA_B200_RNG_PHILOX) o->xh_fresh = false;    // the graph's sampler does not write the host mirror
    return 0;
}

int lmcma_b200_sync(lmcma_b200_opt* o) {
    ARG(o, "null handle");
    CU(cudaSetDevice(o->cfg.device));
    CU(cudaStreamSynchronize(o->stream));
    { int lost = check_lost(o); if (lost) return lost; }
    if (o->graph_dbg) {
        long long h[64];
        cudaMemcpy(h, o->graph_dbg, sizeof(h), cudaMemcpyDeviceToHost);
        fprintf(stderr, "fused generation, k_update (ns since its start): bookkeeping=%lld prologue=%lld sweep: warp 0 done=%lld, all done + ranks seen=%lld post start=%lld mean=%lld newest row=%lld tail=%lld end=%lld",
                h[1] - h[0], h[2] - h[0], h[7] - h[0], h[8] - h[0], h[3] - h[0], h[4] - h[0], h[9] - h[0], h[5] - h[0], h[6] - h[0]);
        fprintf(stderr, " chain blocks:");
        for (int i = 12; i < 21; ++i) fprintf(stderr, " %lld", (h[i] - h[0]) / 100);
        fprintf(stderr, " rows:");
        for (int i = 0; i < 40; ++i) fprintf(stderr, " %lld", (h[24 + i] - h[0]) / 100);
        if (o->d.dbg) {
            long long g[64];
            cudaMemcpy(g, o->d.dbg, sizeof(g), cudaMemcpyDeviceToHost);
            fprintf(stderr, " | k_sample CTA 0: start=%lld flags=%lld loop done=%lld end=%lld", g[32] - h[0], g[33] - h[0], g[37] - h[0], g[38] - h[0]);
            fprintf(stderr, " | k_rank CTA 0: start=%lld ranked=%lld compacted=%lld gathered=%lld stored=%lld", g[16] - h[0], g[17] - h[0], g[18] - h[0], g[19] - h[0], g[20] - h[0]);
            fprintf(stderr, " last slice stored=%lld", g[21] - h[0]);
            cudaMemset(o->d.dbg + 21, 0, sizeof(long long));
            cudaDeviceSynchronize();
        }
        fprintf(stderr, "\n");
    }
    return 0;
}

int lmcma_b200_last_run_ms(lmcma_b200_opt* o, float* ms) {
    ARG(o && ms, "null pointer");
    if (!o->have_run_timing) return fail(LMCMA_B200_ERR_STATE, "no run yet");
    CU(cudaEventSynchronize(o->ev1));
    CU(cudaEventElapsedTime(ms, o->ev0, o->ev1));
    return 0;
}

int lmcma_b200_profile_kernels(lmcma_b200_opt* o, int32_t generations, float* ms4) {
    ARG(o && ms4 && generations >= 1, "bad argument");
    if (!o->map) return fail(LMCMA_B200_ERR_STATE, "no cost attached");
    if (o->cfg.rng != LMCMA_B200_RNG_PHILOX) return fail(LMCMA_B200_ERR_STATE, "profile_kernels needs the PHILOX rng");
    CU(cudaSetDevice(o->cfg.device));
    CostArgs ca;
    int rc = cost_args_for(o, &ca);
    if (rc) return rc;
    cudaStream_t st = o->stream;
    std::vector<cudaEvent_t> ev((size_t)generations * 5);
    for (auto& e : ev) CU(cudaEventCreate(&e));
    for (int g = 0; g < generations && !rc; ++g) {
        cudaEvent_t* e = &ev[(size_t)g * 5];
        cudaEventRecord(e[0], st);
        rc = launch_cost(o->map->dev, ca, o->d.pop_count, o->d.B, o->cost_shape, false, st);
        cudaEventRecord(e[1], st);
        if (!rc) rc = launch_rank(o, o->d.fit, RANK_PLAIN, nullptr, st);
        cudaEventRecord(e[2], st);
        if (!rc) rc = launch_update(o, update_args_local(o), false, st);   // serialised here so that the events separate the kernels
        cudaEventRecord(e[3], st);
        if (!rc) rc = launch_sample(o, st);
        cudaEventRecord(e[4], st);
    }
    cudaError_t se = cudaStreamSynchronize(st);
    if (!rc && se != cudaSuccess) rc = fail(LMCMA_B200_ERR_CUDA, "profile run: %s", cudaGetErrorString(se));
    if (!rc && o->tune.cost_dbg) {     // debug: per-CTA timeline of k_cost (phase stamps), one extra launch
        const size_t ctas = (size_t)o->d.pop_count * o->d.B;
        long long* dbg = nullptr;
        if (cudaMalloc(&dbg, ctas * 8 * sizeof(long long)) == cudaSuccess) {
            cudaMemset(dbg, 0, ctas * 8 * sizeof(long long));
            cudaDeviceSynchronize();
            CostArgs cd = ca; cd.dbg = dbg; cd.cells = nullptr; cd.max_cells = 0;
            launch_cost(o->map->dev, cd, o->d.pop_count, o->d.B, o->cost_shape, true, st);   // the TRACE build carries the stamps (no cells requested)
            cudaStreamSynchronize(st);
            std::vector<long long> h(ctas * 8);
            cudaMemcpy(h.data(), dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
            long long t0 = h[0];
            for (size_t c = 0; c < ctas; ++c) t0 = std::min(t0, h[c * 8]);
            static const char* nm[] = {"start", "x loaded", "phase 1 done", "loop done", "end samples done", "end"};
            fprintf(stderr, "k_cost timeline over %zu CTAs (ns since the first CTA start; min / median / max):", ctas);
            for (int k = 0; k < 6; ++k) {
                std::vector<long long> v(ctas);
                for (size_t c = 0; c < ctas; ++c) v[c] = h[c * 8 + k] - t0;
                std::sort(v.begin(), v.end());
                fprintf(stderr, " %s=%lld/%lld/%lld", nm[k], v[0], v[ctas / 2], v[ctas - 1]);
            }
            // per-CTA durations of the phases
            static const char* dn[] = {"x load", "phase 1", "loop", "end samples", "reduce"};
            fprintf(stderr, " | per-CTA phase durations (median):");
            for (int k = 0; k < 5; ++k) {
                std::vector<long long> v(ctas);
                for (size_t c = 0; c < ctas; ++c) v[c] = h[c * 8 + k + 1] - h[c * 8 + k];
                std::sort(v.begin(), v.end());
                fprintf(stderr, " %s=%lld", dn[k], v[ctas / 2]);
            }
            fprintf(stderr, "\n");
            cudaFree(dbg);
        }
    }
    if (!rc && o->tune.update_dbg) {   // debug: timeline of k_update
        long long* dbg = nullptr;
        if (cudaMalloc(&dbg, 64 * sizeof(long long)) == cudaSuccess) {
            cudaMemset(dbg, 0, 64 * sizeof(long long));
            UpdateArgs ua = update_args_local(o);
            ua.dbg = dbg;
            launch_cost(o->map->dev, ca, o->d.pop_count, o->d.B, o->cost_shape, false, st);
            launch_rank(o, o->d.fit, RANK_PLAIN, nullptr, st);
            launch_update(o, ua, true, st);
            launch_sample(o, st, true);
            cudaStreamSynchronize(st);
            long long h[64];
            cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost);
            static const char* names[] = {"start", "bookkeeping", "prologue done", "k_rank done", "mean", "sweep", "end"};
            fprintf(stderr, "k_update timeline (ns since start; first_stale=%lld live=%lld):", h[10], h[11]);
            for (int k = 1; k < 7; ++k) fprintf(stderr, " %s=%lld", names[k], h[k] - h[0]);
            fprintf(stderr, "\n");
            cudaFree(dbg);
        }
    }
    if (!rc && o->d.dbg) {
        long long h[64];
        cudaMemcpy(h, o->d.dbg, sizeof(h), cudaMemcpyDeviceToHost);
        fprintf(stderr, "k_sample timeline (CTA 0, ns):");
        static const char* nm[] = {"start", "griddep", "order", "z ready", "first pairs", "loop done", "end"};
        for (int k = 1; k < 7; ++k) fprintf(stderr, " %s=%lld", nm[k], h[32 + k] - h[32]);
        fprintf(stderr, " | chunk waits:");
        for (int k = 7; k < 12; ++k) fprintf(stderr, " %lld", h[32 + k] - h[32]);
        fprintf(stderr, " dots done=%lld exchanged=%lld", h[32 + 15] - h[32], h[32 + 16] - h[32]);
        fprintf(stderr, "\n");
    }
    for (int k = 0; k < 4; ++k) ms4[k] = 0.f;
    if (!rc)
        for (int g = 0; g < generations; ++g)
            for (int k = 0; k < 4; ++k) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, ev[(size_t)g * 5 + k], ev[(size_t)g * 5 + k + 1]);
                ms4[k] += ms / generations;
            }
    for (auto& e : ev) cudaEventDestroy(e);
    o->x_cache_valid = false;
    return rc;
}

int lmcma_b200_best(lmcma_b200_opt* o, float* x_best, float* f_best) {
    ARG(o, "null handle");
    CU(cudaSetDevice(o->cfg.device));
    const OptDev& d = o->d;
    if (x_best) {
        int rc = d2h_rows(x_best, d.best_x, d.B, d.n * sizeof(float), d.ns * sizeof(float), o->stream);
        if (rc) return rc;
    }
    if (f_best) {
        std::vector<Scalars> sc(d.B);
        CU(cudaMemcpyAsync(sc.data(), d.sc, sc.size() * sizeof(Scalars), cudaMemcpyDeviceToHost, o->stream));
        CU(cudaStreamSynchronize(o->stream));
        for (int b = 0; b < d.B; ++b) f_best[b] = (float)sc[b].best_f;
    }
    return 0;
}

// ---- split-population mode ---------------------------------------------------------------------
int lmcma_b200_mg_payload_floats(lmcma_b200_opt* o, int32_t* floats_out) {
    ARG(o && floats_out, "null pointer");
    *floats_out = o->d.ns + 4;
    return 0;
}

int lmcma_b200_mg_evaluate(lmcma_b200_opt* o, float* f_local_dev, void* stream) {
    ARG(o && f_local_dev, "null pointer");
    if (o->d.B != 1) return fail(LMCMA_B200_ERR_STATE, "split-population mode needs batch == 1");
    if (o->needs_sample) return fail(LMCMA_B200_ERR_STATE, "no population yet");
    CU(cudaSetDevice(o->cfg.device));
    CostArgs ca;
    int rc = cost_args_for(o, &ca);
    if (rc) return rc;
    ca.f = f_local_dev; ca.f_stride = o->d.pop_count; ca.f_offset = 0;
    cudaStream_t st = stream ? (cudaStream_t)stream : o->stream;
    if ((rc = apply_l2_window(o->map, st))) return rc;
    return launch_cost(o->map->dev, ca, o->d.pop_count, o->d.B, o->cost_shape, false, st);
}

int lmcma_b200_mg_rank(lmcma_b200_opt* o, const float* f_all_dev, float* payload_dev, void* stream) {
    ARG(o && f_all_dev && payload_dev, "null pointer");
    if (o->d.B != 1) return fail(LMCMA_B200_ERR_STATE, "split-population mode needs batch == 1");
    CU(cudaSetDevice(o->cfg.device));
    cudaStream_t st = stream ? (cudaStream_t)stream : o->stream;
    // keep the gathered fitness: k_update needs it for prev_fit / best tracking
    CU(cudaMemcpyAsync(o->d.fit, f_all_dev, (size_t)o->d.lambda * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return launch_rank(o, o->d.fit, RANK_PACK, payload_dev, st);
}

int lmcma_b200_mg_update(lmcma_b200_opt* o, const float* payload_all_dev, int32_t world, void* stream) {
    ARG(o && payload_all_dev && world >= 1, "bad argument");
    if (o->cfg.rng != LMCMA_B200_RNG_PHILOX) return fail(LMCMA_B200_ERR_STATE, "split-population mode needs the PHILOX rng");   // before anything is launched
    CU(cudaSetDevice(o->cfg.device));
    cudaStream_t st = stream ? (cudaStream_t)stream : o->stream;
    const long long pf = o->d.ns + 4;
    UpdateArgs a;
    memset(&a, 0, sizeof(a));
    a.f_all = o->d.fit; a.slices = payload_all_dev; a.n_slices = world;
    a.slice_stride = (long long)o->d.B * pf; a.inst_stride = pf; a.payload_mode = 1;
    int rc = launch_update(o, a, false, st);
    if (rc) return rc;
    o->x_cache_valid = false;
    o->mirror_suppressed = true;
    rc = launch_sample(o, st, true);
    o->mirror_suppressed = false;
    return rc;
}

}  // extern "C"
