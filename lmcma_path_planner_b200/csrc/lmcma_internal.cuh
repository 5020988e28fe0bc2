// lmcma_internal.cuh — what the translation units of liblmcma_b200.so share: error helpers, the create-time tuning knobs,
// the two handle types and the launch helpers that cross a file boundary.
//   lmcma_capi.cu        library / device queries, the optimiser entry points, launch configuration, CUDA graphs
//   lmcma_capi_map.cu    cost map handles, distance transform, the cost evaluator (k_cost, k_edt, k_brick)
//   lmcma_capi_state.cu  state getters / setters, the host-side pieces of the reference API
#pragma once
#include <cuda_runtime.h>
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>
#include <string>
#include <vector>
#include "../../include/lmcma_b200.h"
#include "lmcma_host.hpp"
#include "lmcma_common.cuh"

namespace lmcma_capi {
using namespace lmcma;

extern thread_local std::string g_err;     // lmcma_capi.cu
extern std::atomic<long long> g_launches;

inline int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                    \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return fail(LMCMA_B200_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)
#define ARG(cond, msg)                                                   \
    do {                                                                 \
        if (!(cond)) return fail(LMCMA_B200_ERR_ARG, "%s (%s)", msg, #cond); \
    } while (0)

template <class T>
inline int dmalloc(T** p, size_t count) {
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), std::max<size_t>(count, 1) * sizeof(T));
    if (e != cudaSuccess) return fail(LMCMA_B200_ERR_NOMEM, "cudaMalloc(%zu bytes): %s", count * sizeof(T), cudaGetErrorString(e));
    e = cudaMemset(*p, 0, std::max<size_t>(count, 1) * sizeof(T));
    // the memset is asynchronous on the legacy stream, which the library's non-blocking streams do not wait for: without
    // this wait it can land AFTER a copy / kernel that one of those streams issues into the new buffer
    if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamLegacy);
    if (e != cudaSuccess) return fail(LMCMA_B200_ERR_CUDA, "cudaMemset: %s", cudaGetErrorString(e));
    return 0;
}
#define DM(ptr, count)                              \
    do {                                            \
        int rc__ = dmalloc(&(ptr), (count));        \
        if (rc__) return rc__;                      \
    } while (0)


inline int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

// Every LMCMA_B200_* environment knob, read ONCE when a handle (map or optimiser) is created and kept on the handle:
// nothing on a launch path calls getenv.  All of them are experiment / debugging switches; the defaults are the product.
struct Tuning {
    int cost_minb = 0 /* 0 = by launch size */, cost_tpt = 0, cost_cb = 0, zerocopy = 1;
    int sample_rows = -1;   // k_sample_rows: -1 = where it pays (many rows), 0 = never, 1 = also for one large population
    int sample_spec = 1, sample_threads = 0, sample_smem_kb = 160, sample_smem_kb_set = 0, sample_narrow = 0, sample_rbw = 0, sample_r = 0;
    int update_blocked = 0, update_gram = 0, update_streaming = 0, update_sweep_warps = 0;
    int progressive = 1, overlap = 1, tell_overlap = 1, rank_sorted = 1;
    int rank_late = 0;      // fused generation: 1 = k_sample is released when k_rank's CTAs are done, 0 = when k_rank starts (it then works
                            // through the finished pairs beside the ranking: 66.5 -> 63.2 us per C2 generation once the newest row's
                            // chain stopped being the longer branch)
    int tell_spec = 1;      // tell_all: speculative update of the next generation at the end of the graph (LMCMA_B200_TELL_SPEC=0: off)
    int graph_unroll = 8;   // fused generations per CUDA graph for runs of at least that many (LMCMA_B200_GRAPH_UNROLL, 1 = off)
    int update_warps = 0;   // k_update CTA size: 0 = by batch size, 8 / 16 forced (LMCMA_B200_UPDATE_WARPS)
    int update_dry = 1;     // overlapped generation: pre-execute the post-rank code while k_rank is busy (k_update.cuh)
    int graph_dbg = 0, dbg = 0, update_dbg = 0, cost_dbg = 0;
    static Tuning from_env() {
        Tuning t;
        t.cost_minb = env_int("LMCMA_B200_COST_MINB", t.cost_minb);
        t.cost_tpt = env_int("LMCMA_B200_COST_TPT", 0);
        t.cost_cb = env_int("LMCMA_B200_COST_CB", 0);
        t.zerocopy = env_int("LMCMA_B200_ZEROCOPY", 1);
        t.sample_spec = env_int("LMCMA_B200_SAMPLE_SPEC", 1);
        t.sample_rows = env_int("LMCMA_B200_SAMPLE_ROWS", -1);
        t.sample_threads = env_int("LMCMA_B200_SAMPLE_THREADS", 0);
        t.sample_smem_kb_set = env_int("LMCMA_B200_SAMPLE_SMEM_KB", 0);
        t.sample_smem_kb = t.sample_smem_kb_set ? t.sample_smem_kb_set : 160;
        t.sample_narrow = env_int("LMCMA_B200_SAMPLE_NARROW", 0);
        t.sample_rbw = env_int("LMCMA_B200_SAMPLE_RBW", 0);
        t.sample_r = env_int("LMCMA_B200_SAMPLE_R", 0);
        t.update_blocked = env_int("LMCMA_B200_UPDATE_BLOCKED", 0);
        t.update_gram = env_int("LMCMA_B200_UPDATE_GRAM", 0);
        t.update_streaming = env_int("LMCMA_B200_UPDATE_STREAMING", 0);
        t.update_sweep_warps = env_int("LMCMA_B200_UPDATE_SWEEP_WARPS", 0);
        t.progressive = env_int("LMCMA_B200_PROGRESSIVE", 1);
        t.overlap = env_int("LMCMA_B200_OVERLAP", 1);
        t.tell_overlap = env_int("LMCMA_B200_TELL_OVERLAP", 1);
        t.rank_late = env_int("LMCMA_B200_RANK_LATE", 0);
        t.update_dry = env_int("LMCMA_B200_UPDATE_DRY", 1);
        t.tell_spec = env_int("LMCMA_B200_TELL_SPEC", 1);
        t.update_warps = env_int("LMCMA_B200_UPDATE_WARPS", 0);
        t.graph_unroll = env_int("LMCMA_B200_GRAPH_UNROLL", 8);
        if (t.graph_unroll < 1 || t.graph_unroll > 64) t.graph_unroll = 1;
        t.rank_sorted = env_int("LMCMA_B200_RANK_SORTED", 1);
        t.graph_dbg = env_int("LMCMA_B200_GRAPH_DBG", 0);
        t.dbg = getenv("LMCMA_B200_DBG") ? 1 : 0;
        t.update_dbg = getenv("LMCMA_B200_UPDATE_DBG") ? 1 : 0;
        t.cost_dbg = getenv("LMCMA_B200_COST_DBG") ? 1 : 0;
        return t;
    }
};

struct DeviceProps {
    int sm_count = 0;
    size_t l2 = 0, smem_optin = 0, persist_max = 0;
    int cc = 0;
    bool ok = false;
    int cosched = -1;     // probe_coschedule: -1 not probed yet, 0 branches of a forked graph are serialised here, 1 they run concurrently
};
extern DeviceProps g_props[64];           // lmcma_capi.cu
inline int query_props(int device, DeviceProps** out) {
    ARG(device >= 0 && device < 64, "device ordinal out of range");
    DeviceProps& p = g_props[device];
    if (!p.ok) {
        cudaDeviceProp dp;
        CU(cudaGetDeviceProperties(&dp, device));
        p.sm_count = dp.multiProcessorCount;
        p.l2 = dp.l2CacheSize;
        p.smem_optin = dp.sharedMemPerBlockOptin;
        p.persist_max = dp.persistingL2CacheMaxSize;
        p.cc = dp.major * 10 + dp.minor;
        if (dp.major < 10) return fail(LMCMA_B200_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, dp.major, dp.minor);
        p.ok = true;
    }
    *out = &p;
    return 0;
}

}  // namespace lmcma_capi

// =================================================================================================
// handles (global namespace: the opaque types of include/lmcma_b200.h)
// =================================================================================================
struct lmcma_b200_map {
    int device = 0;
    lmcma_capi::Tuning tune;
    lmcma::MapDev dev{};
    int storage = 0;
    float c_min = 0.5f, scale = 1.f;
    size_t cells = 0, stored = 0;   // logical cells / stored elements (bricked, padded)
    float* d_g32 = nullptr;
    unsigned char* d_q8 = nullptr;
    float* d_lut = nullptr;
    bool persist = false;
    cudaStream_t stream = nullptr;   // private stream for the stand-alone evaluate calls
    // staging for the host-buffer evaluate path
    float* d_X = nullptr; size_t d_X_cap = 0;
    float* d_f = nullptr; int* d_nc = nullptr; int* d_ns = nullptr; size_t d_out_cap = 0;
    std::mutex host_path;            // lmcma_b200_cost_evaluate / cost_trace share the staging buffers and the private stream
};

struct lmcma_b200_opt {
    lmcma_b200_config cfg{};
    lmcma_capi::Tuning tune;
    lmcma::OptDev d{};
    lmcma_capi::DeviceProps* props = nullptr;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    // host mirrors
    std::vector<double> weights;       // mu
    std::vector<float> lo_f, hi_f;
    float* d_lo = nullptr; float* d_hi = nullptr; float* d_w = nullptr;
    // rng
    lmcma::HansenStream hansen{1};
    std::vector<float> z_host;         // staging for HANSEN
    bool needs_sample = false, pending_z = false;
    // reference one-at-a-time protocol
    int sample_idx = 0;
    std::vector<float> x_cache; bool x_cache_valid = false;
    std::vector<float> f_host;
    // attached cost
    lmcma_b200_map* map = nullptr;
    lmcma_b200_objective obj{};
    float* d_ends = nullptr;
    // graph
    cudaGraphExec_t graph_exec = nullptr;
    cudaGraphExec_t graph_exec_multi = nullptr;   // Tuning::graph_unroll generations in one graph (lmcma_b200_run of many generations)
    cudaStream_t graph_built_for = nullptr;
    cudaGraphExec_t tell_graph = nullptr;     // tell_all of one query: H2D fitness -> k_rank -> k_sample, k_update on a side branch
    cudaStream_t tell_graph_for = nullptr;
    bool tell_graph_failed = false;
    // speculative update (k_update.cuh, UpdateArgs::phase): every tell_all graph ends with the fitness-independent part of the
    // NEXT generation's update (into d_spec, beside the sampler's tail and the candidates' trip across PCIe); the next tell_all
    // resumes from it (tell_graph_resume) unless anything else touched the optimiser's state in between
    cudaGraphExec_t tell_graph_resume = nullptr;
    float* d_spec = nullptr; size_t spec_stride = 0;
    bool spec_valid = false;
    bool tell_graph_spec = false;             // the tell graphs were captured with the speculative pass (host mirror on)
    float* f_pinned = nullptr;                // the graph's copy source (the caller's fitness array is copied here first)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool have_run_timing = false;
    // sample launch config
    int smp_threads = 128, smp_kc = 1, smp_nv = 1, smp_rb = 1, smp_stages = 2;
    bool smp_wide = false; int smp_R = 1, smp_CW = 1, smp_qpw = 32, smp_RBW = 1;
    bool smp_rows = false; int smp_rows_qw = 8, smp_rows_rl = 1, smp_rows_kc = 8, smp_rows_stages = 2, smp_rows_region = 0;   // k_sample_rows
    size_t smp_rows_smem = 0;
    float* d_Lf = nullptr;             // lower Cholesky factor of the smoothness prior (n x ns FP32) or null
    bool mirror_dirty = true;          // the sequence-ordered pair mirror must be rebuilt (k_pack_pairs) before sampling
    size_t smp_smem = 0;
    lmcma::CostShape cost_shape;
    bool progressive = false;   // k_update -> k_sample hand-over inside the fused generation (k_update.cuh)
    bool overlap = false;       // fused generation with k_update on a side branch, concurrent with k_cost / k_rank (k_update.cuh)
    cudaStream_t side_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    float* x_mirror = nullptr;        // page-locked, mapped host mirror of X (lmcma_b200_ask_all_view); OptDev::Xh is its device alias
    bool mirror_on = false, mirror_suppressed = false, xh_fresh = false;
    int* err_host = nullptr;          // page-locked, mapped: OptDev::err (a kernel of the overlapped generation gave up on its partner)
    long long* graph_dbg = nullptr;   // LMCMA_B200_GRAPH_DBG: k_update's timeline inside the fused generation, printed by lmcma_b200_sync
    int upd_nvb = 4, upd_rmax = 0, upd_sweep_warps = 16, upd_warps = 16;
    bool upd_gram = false; size_t coef_smem = 0;   // Gram-matrix recompute (k_gram.cuh) for rows that fit neither registers nor smem
    bool upd_rows_in_smem = true;
    size_t upd_smem = 0, rank_smem = 0;
    int rank_threads = 1024;
    size_t cost_smem = 0;
};


namespace lmcma_capi {
// ---- defined in lmcma_capi_map.cu ----
int launch_cost(const MapDev& mp, const CostArgs& a, int rows, int B, CostShape shape, bool trace, cudaStream_t st);
CostShape pick_cost_shape(int W, const float* start, const float* goal, int dims, const Tuning& tune);
int apply_l2_window(lmcma_b200_map* m, cudaStream_t st);
int check_obj(const lmcma_b200_map* m, const lmcma_b200_objective* obj);
void register_mirror(const void* host, size_t bytes, const float* dev, long long ld, int device);
void unregister_mirror(const void* host);
// ---- defined in lmcma_capi.cu ----
int check_lost(lmcma_b200_opt* o);

// dense <-> pitched copies
inline int d2h_rows(void* dst, const void* src, size_t rows, size_t width_bytes, size_t src_pitch_bytes, cudaStream_t st) {
    if (width_bytes == src_pitch_bytes) CU(cudaMemcpyAsync(dst, src, rows * width_bytes, cudaMemcpyDeviceToHost, st));   // dense: one 1-D copy
    else CU(cudaMemcpy2DAsync(dst, width_bytes, src, src_pitch_bytes, width_bytes, rows, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return 0;
}
inline int h2d_rows(void* dst, const void* src, size_t rows, size_t width_bytes, size_t dst_pitch_bytes, cudaStream_t st) {
    if (width_bytes == dst_pitch_bytes) CU(cudaMemcpyAsync(dst, src, rows * width_bytes, cudaMemcpyHostToDevice, st));
    else CU(cudaMemcpy2DAsync(dst, dst_pitch_bytes, src, width_bytes, width_bytes, rows, cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));
    return 0;
}

}  // namespace lmcma_capi
