/* TEST INFRASTRUCTURE ONLY (oracle/).  CPU restatement of the trajectory cost the LM-CMA planner
 * minimises.  The reference has no such function; it is ASSEMBLED from the
 * reference's cost-model pieces, each cited below, with the quadrature choices the reference
 * leaves to un-vendored OMPL fixed here (marked DECLARED).
 * PINNED to reference-executed code: the per-sample pieces (matrix cell incl. nearbyint ties, valid <=> E > 0,
 * state cost 1 / clearance above the floor) and the objective weights — sample_state() below is checked against the
 * reference's own ValidityChecker / ClearanceObjective / shortrisky / longsafe compiled from planner.cpp:587-690
 * (oracle/_ref/libref_cost.so, OMPL stand-in) and against tests/golden/cost_pieces_reference.json generated from them
 * (tests/test_cost_reference_pieces.py).
 * PARITY UNPINNED for the DECLARED choices only (sub-step rule, trapezoid, c_min floor, out-of-map = collision): the
 * path integral is OMPL's (un-vendored, unversioned), no reference test or golden vector exists at that boundary; for
 * those this file is the single source of truth the CUDA cost kernel is checked against.
 *
 *   parameters  x[d*W + w], dimension-major (lmcma.cpp:786-791), D = 2 or 3, W interior waypoints
 *   poly-line   P_0 = start, P_1..P_W, P_{W+1} = goal      (start/goal fixed: planner.cpp:701-711)
 *   map         E[row=y][col=x] (2-D) / E[z][y][x] (3-D): distance to the nearest obstacle in cells,
 *               0 on obstacles                             (planner.cpp:597-602, 628)
 *   cell        col = rint(x), row = rint(y) (nearbyint, half-to-even, planner.cpp:595-596),
 *               computed in FP32 with un-contracted mul/add so that the GPU reproduces the index
 *               bit for bit; DECLARED: a sample outside the map is a collision
 *   segment     len = |B-A|_2 (planner.cpp:638, OMPL path length); DECLARED: K = clamp(ceil(|B-A|_inf),
 *               1, 65536) sub-steps, samples Q_k = A + (k * (1/K)) * (B-A), k = 0..K
 *   state cost  g = 1 / clearance (planner.cpp:667); DECLARED: clearance floored at c_min, and a
 *               colliding sample (E <= 0, the negation of isValid planner.cpp:602) contributes 1/c_min
 *   clearance   integral = (len/K) * sum_{k<K} (g_k + g_{k+1})/2  (trapezoid = OMPL's
 *               StateCostIntegralObjective with motion-cost interpolation on, planner.cpp:651)
 *   collisions  every distinct poly-line sample counted once: k in [0,K) of each segment + the goal
 *   fitness     f = w_len * sum len + w_clr * sum integral + w_col * collisions
 *               (w_len, w_clr) = (100, 1) "shortrisky" / (1, 1000) "longsafe" (planner.cpp:677-690)
 *
 * Sums are accumulated in FP64 here; the GPU accumulates in FP32 (tolerance 1e-5 relative).
 * Compile: gcc -O2 -ffp-contract=off -pthread -shared -fPIC  (see oracle/Makefile)
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_KMAX 65536
#define ORC_BIG 1e20

typedef struct {
    int dims;            /* 2 or 3 */
    int shape[3];        /* nx, ny, nz (nz = 1 in 2-D) */
    const float* dist;   /* E, row-major [z][y][x] */
    float c_min;
    float start[3], goal[3];
    int waypoints;       /* W */
    float w_len, w_clr, w_col;
    int threads;
} orc_problem;

static inline int sub_steps(float linf) {
    /* NaN compares false everywhere -> 1 */
    if (linf >= 1.0f) return linf <= (float)ORC_KMAX ? (int)ceilf(linf) : ORC_KMAX;
    return 1;
}

/* linear cell index, or -1 when the sample is outside the map (or NaN) */
static inline int64_t cell_of(const orc_problem* p, const float q[3]) {
    int64_t idx = 0, stride = 1;
    for (int c = 0; c < p->dims; ++c) {
        float r = rintf(q[c]);
        if (!(r >= 0.0f && r <= (float)(p->shape[c] - 1))) return -1;
        idx += (int64_t)r * stride;
        stride *= p->shape[c];
    }
    return idx;
}

/* One poly-line sample q: its cell (planner.cpp:595-600), whether it collides (the negation of
 * ValidityChecker::isValid, planner.cpp:602: valid <=> E > 0) and its state cost 1 / clearance
 * (ClearanceObjective::stateCost, planner.cpp:667) with the DECLARED floor / out-of-map rule.  This is the function
 * eval_one runs per sample; orc_cost_sample exposes it so that tests/test_cost_reference_pieces.py can hold it against
 * the reference's own ValidityChecker / ClearanceObjective (oracle/_ref/libref_cost.so). */
static inline void sample_state(const orc_problem* p, const float q[3], int64_t* cell_out, int* hit_out, double* g_out) {
    const double g_coll = 1.0 / (double)p->c_min;
    const int64_t cell = cell_of(p, q);
    int hit; double g;
    if (cell < 0) { hit = 1; g = g_coll; }
    else {
        const float e = p->dist[cell];
        hit = !(e > 0.0f);
        g = hit ? g_coll : 1.0 / (double)(e > p->c_min ? e : p->c_min);
    }
    *cell_out = cell; *hit_out = hit; *g_out = g;
}
void orc_cost_sample(const orc_problem* p, const float* q, int count, int64_t* cell_out, int* hit_out, double* g_out) {
    for (int i = 0; i < count; ++i) {
        float qq[3] = {q[3 * i], q[3 * i + 1], q[3 * i + 2]};
        sample_state(p, qq, &cell_out[i], &hit_out[i], &g_out[i]);
    }
}

static inline void waypoint(const orc_problem* p, const float* x, int i, float out[3]) {
    const int W = p->waypoints;
    out[2] = 0.0f;
    if (i == 0) { for (int c = 0; c < p->dims; ++c) out[c] = p->start[c]; }
    else if (i == W + 1) { for (int c = 0; c < p->dims; ++c) out[c] = p->goal[c]; }
    else { for (int c = 0; c < p->dims; ++c) out[c] = x[c * W + (i - 1)]; }
}

/* one trajectory; cells_out (nullable) receives the visited linear cell indices (all K+1 samples
 * of every segment, -1 = outside), at most max_cells of them */
static void eval_one(const orc_problem* p, const float* x, double* f, int* ncoll, int* nsamp,
                     double* len_out, double* clr_out, int64_t* cells_out, int64_t max_cells) {
    const int W = p->waypoints;
    double len_sum = 0.0, clr_sum = 0.0;
    int coll = 0; int64_t samples = 0;
    float A[3], B[3], d[3], q[3];
    for (int s = 0; s <= W; ++s) {
        waypoint(p, x, s, A);
        waypoint(p, x, s + 1, B);
        float linf = 0.0f; double l2 = 0.0;
        int bad = 0;
        for (int c = 0; c < p->dims; ++c) {
            d[c] = B[c] - A[c];
            float a = fabsf(d[c]);
            if (a != a) bad = 1;
            if (a > linf) linf = a;
            l2 += (double)d[c] * (double)d[c];
        }
        if (bad) linf = NAN;
        const int K = sub_steps(linf);
        const float invK = 1.0f / (float)K;
        const double len = sqrt(l2);
        double acc = 0.0;
        for (int k = 0; k <= K; ++k) {
            const float t = (float)k * invK;
            for (int c = 0; c < p->dims; ++c) {
                volatile float prod = t * d[c];      /* force the un-contracted FP32 product */
                q[c] = A[c] + prod;
            }
            int64_t cell; int hit; double g;
            sample_state(p, q, &cell, &hit, &g);
            acc += (k == 0 || k == K) ? 0.5 * g : g;
            if (k < K || s == W) coll += hit;
            if (cells_out && samples < max_cells) cells_out[samples] = cell;
            ++samples;
        }
        len_sum += len;
        clr_sum += acc * (len / (double)K);
    }
    *f = (double)p->w_len * len_sum + (double)p->w_clr * clr_sum + (double)p->w_col * (double)coll;
    if (ncoll) *ncoll = coll;
    if (nsamp) *nsamp = (int)samples;
    if (len_out) *len_out = len_sum;
    if (clr_out) *clr_out = clr_sum;
}

/* ---- a tiny pthread parallel-for (this image's default gcc wrapper has no libgomp) ---- */
typedef struct {
    const orc_problem* p; const float* Xf; const double* Xd; int count, n;
    double* f; int* ncoll; int* nsamp; double* len_out; double* clr_out;
    int tid, nthreads;
} orc_job;

static void* job_main(void* arg) {
    orc_job* j = (orc_job*)arg;
    float* tmp = j->Xd ? (float*)malloc(sizeof(float) * (size_t)j->n) : 0;
    /* interleaved blocks of 4 candidates per thread */
    for (int base = j->tid * 4; base < j->count; base += j->nthreads * 4)
        for (int i = base; i < base + 4 && i < j->count; ++i) {
            const float* x;
            if (j->Xd) {
                for (int k = 0; k < j->n; ++k) tmp[k] = (float)j->Xd[(size_t)i * j->n + k];
                x = tmp;
            } else x = j->Xf + (size_t)i * j->n;
            eval_one(j->p, x, &j->f[i], j->ncoll ? &j->ncoll[i] : 0, j->nsamp ? &j->nsamp[i] : 0,
                     j->len_out ? &j->len_out[i] : 0, j->clr_out ? &j->clr_out[i] : 0, 0, 0);
        }
    free(tmp);
    return 0;
}

static void run_jobs(orc_job proto) {
    int nt = proto.p->threads > 0 ? proto.p->threads : 1;
    if (nt > 256) nt = 256;
    if (nt == 1) { proto.tid = 0; proto.nthreads = 1; job_main(&proto); return; }
    pthread_t th[256]; orc_job jobs[256];
    for (int t = 0; t < nt; ++t) { jobs[t] = proto; jobs[t].tid = t; jobs[t].nthreads = nt; pthread_create(&th[t], 0, job_main, &jobs[t]); }
    for (int t = 0; t < nt; ++t) pthread_join(th[t], 0);
}

/* X: count x n FP32 candidates (n = dims * W).  Any output pointer except f may be NULL. */
void orc_cost_batch(const orc_problem* p, const float* X, int count, double* f, int* ncoll, int* nsamp,
                    double* len_out, double* clr_out) {
    orc_job j; memset(&j, 0, sizeof(j));
    j.p = p; j.Xf = X; j.count = count; j.n = p->dims * p->waypoints;
    j.f = f; j.ncoll = ncoll; j.nsamp = nsamp; j.len_out = len_out; j.clr_out = clr_out;
    run_jobs(j);
}

/* FP64-candidate entry with the signature oracle/ref_harness.cpp's ref_lmcma_generation expects:
 * the candidates are rounded to FP32 first (the device stores candidates in FP32). */
void orc_cost_batch_f64(const double* X, int count, int n, double* f, void* ctx) {
    orc_job j; memset(&j, 0, sizeof(j));
    j.p = (const orc_problem*)ctx; j.Xd = X; j.count = count; j.n = n; j.f = f;
    run_jobs(j);
}

/* visited cells of ONE trajectory; returns the number of samples (may exceed max_cells) */
int64_t orc_cost_trace(const orc_problem* p, const float* x, int64_t* cells_out, int64_t max_cells) {
    double f; int nc, ns;
    eval_one(p, x, &f, &nc, &ns, 0, 0, cells_out, max_cells);
    return ns;
}

/* ---- distance fields ----------------------------------------------------------------------- */

/* Exact Euclidean distance to the nearest obstacle cell (occ != 0), 0 on obstacles, by the
 * separable lower-envelope method; clamp > 0 caps the result.  Replaces the un-vendored
 * dynamicEDT3D the reference calls (planner.cpp:81-87, 305-307).  PARITY UNPINNED (cross-checked
 * against scipy.ndimage.distance_transform_edt in tests). */
static void edt_1d(const double* fsrc, int n, double* out, int* v, double* z) {
    /* lower envelope of the parabolas q -> (q-r)^2 + fsrc[r]; non-seed cells carry ORC_BIG */
    int k = 0; v[0] = 0; z[0] = -ORC_BIG; z[1] = ORC_BIG;
    for (int q = 1; q < n; ++q) {
        double s;
        for (;;) {
            const int r = v[k];
            s = ((fsrc[q] + (double)q * q) - (fsrc[r] + (double)r * r)) / (2.0 * q - 2.0 * r);
            if (s <= z[k]) { --k; continue; }
            break;
        }
        ++k; v[k] = q; z[k] = s; z[k + 1] = ORC_BIG;
    }
    k = 0;
    for (int q = 0; q < n; ++q) {
        while (z[k + 1] < (double)q) ++k;
        const int r = v[k];
        out[q] = (double)(q - r) * (q - r) + fsrc[r];
    }
}

void orc_edt_exact(const uint8_t* occ, int dims, const int* shape, float clamp, float* dist) {
    const int nx = shape[0], ny = shape[1], nz = dims == 3 ? shape[2] : 1;
    const size_t total = (size_t)nx * ny * nz;
    double* sq = (double*)malloc(sizeof(double) * total);
    for (size_t i = 0; i < total; ++i) sq[i] = occ[i] ? 0.0 : ORC_BIG;
    int maxn = nx > ny ? nx : ny; if (nz > maxn) maxn = nz;
    double* line = (double*)malloc(sizeof(double) * maxn);
    double* res = (double*)malloc(sizeof(double) * maxn);
    int* v = (int*)malloc(sizeof(int) * maxn);
    double* z = (double*)malloc(sizeof(double) * (maxn + 1));
    /* x */
    for (size_t r = 0; r < (size_t)ny * nz; ++r) {
        edt_1d(sq + r * nx, nx, res, v, z);
        memcpy(sq + r * nx, res, sizeof(double) * nx);
    }
    /* y */
    for (int zz = 0; zz < nz; ++zz)
        for (int x = 0; x < nx; ++x) {
            for (int y = 0; y < ny; ++y) line[y] = sq[((size_t)zz * ny + y) * nx + x];
            edt_1d(line, ny, res, v, z);
            for (int y = 0; y < ny; ++y) sq[((size_t)zz * ny + y) * nx + x] = res[y];
        }
    /* z */
    if (nz > 1)
        for (int y = 0; y < ny; ++y)
            for (int x = 0; x < nx; ++x) {
                for (int zz = 0; zz < nz; ++zz) line[zz] = sq[((size_t)zz * ny + y) * nx + x];
                edt_1d(line, nz, res, v, z);
                for (int zz = 0; zz < nz; ++zz) sq[((size_t)zz * ny + y) * nx + x] = res[zz];
            }
    for (size_t i = 0; i < total; ++i) {
        double dd = sqrt(sq[i]);
        if (clamp > 0.0f && dd > (double)clamp) dd = (double)clamp;
        dist[i] = (float)dd;
    }
    free(sq); free(line); free(res); free(v); free(z);
}

/* 8SSEDT: two-pass 8-neighbour sequential propagation of (dx,dy) offsets to the nearest seed
 * (planner.cpp:403-490), generalised from the reference's fixed 100x100 grid to h x w.
 * seed != 0 marks the "inside" cells (offset 0); everything else starts at (9999, 9999)
 * (planner.cpp:421-422).  Output: squared offset length per cell (DistSq, planner.cpp:413). */
typedef struct { int dx, dy; } orc_off;
static inline int off_sq(orc_off o) { return o.dx * o.dx + o.dy * o.dy; }
static inline orc_off grid_get(const orc_off* g, int w, int h, int x, int y) {
    if (x >= 0 && y >= 0 && x < w && y < h) return g[(size_t)y * w + x];
    orc_off far = {9999, 9999};
    return far;
}
static inline void relax(const orc_off* g, int w, int h, orc_off* p, int x, int y, int ox, int oy) {
    orc_off o = grid_get(g, w, h, x + ox, y + oy);
    o.dx += ox; o.dy += oy;
    if (off_sq(o) < off_sq(*p)) *p = o;
}
void orc_8ssedt_sq(const uint8_t* seed, int w, int h, int* sq_out) {
    orc_off* g = (orc_off*)malloc(sizeof(orc_off) * (size_t)w * h);
    for (size_t i = 0; i < (size_t)w * h; ++i) {
        if (seed[i]) { g[i].dx = 0; g[i].dy = 0; } else { g[i].dx = 9999; g[i].dy = 9999; }
    }
    for (int y = 0; y < h; ++y) {                       /* forward pass, planner.cpp:445-463 */
        for (int x = 0; x < w; ++x) {
            orc_off p = grid_get(g, w, h, x, y);
            relax(g, w, h, &p, x, y, -1, 0); relax(g, w, h, &p, x, y, 0, -1);
            relax(g, w, h, &p, x, y, -1, -1); relax(g, w, h, &p, x, y, 1, -1);
            g[(size_t)y * w + x] = p;
        }
        for (int x = w - 1; x >= 0; --x) {
            orc_off p = grid_get(g, w, h, x, y);
            relax(g, w, h, &p, x, y, 1, 0);
            g[(size_t)y * w + x] = p;
        }
    }
    for (int y = h - 1; y >= 0; --y) {                  /* backward pass, planner.cpp:466-486 */
        for (int x = w - 1; x >= 0; --x) {
            orc_off p = grid_get(g, w, h, x, y);
            relax(g, w, h, &p, x, y, 1, 0); relax(g, w, h, &p, x, y, 0, 1);
            relax(g, w, h, &p, x, y, -1, 1); relax(g, w, h, &p, x, y, 1, 1);
            g[(size_t)y * w + x] = p;
        }
        for (int x = 0; x < w; ++x) {
            orc_off p = grid_get(g, w, h, x, y);
            relax(g, w, h, &p, x, y, -1, 0);
            g[(size_t)y * w + x] = p;
        }
    }
    for (size_t i = 0; i < (size_t)w * h; ++i) sq_out[i] = off_sq(g[i]);
    free(g);
}
/* signed distance exactly as the reference renders it (planner.cpp:536-540):
 * int(sqrt(d_to_obstacle^2)) - int(sqrt(d_to_free^2)), obstacle = occ != 0 (g < 128, :515) */
void orc_8ssedt_signed(const uint8_t* occ, int w, int h, int* signed_out) {
    const size_t total = (size_t)w * h;
    uint8_t* inv = (uint8_t*)malloc(total);
    int* a = (int*)malloc(sizeof(int) * total);
    int* b = (int*)malloc(sizeof(int) * total);
    for (size_t i = 0; i < total; ++i) inv[i] = !occ[i];
    orc_8ssedt_sq(occ, w, h, a);
    orc_8ssedt_sq(inv, w, h, b);
    for (size_t i = 0; i < total; ++i)
        signed_out[i] = (int)sqrt((double)a[i]) - (int)sqrt((double)b[i]);
    free(inv); free(a); free(b);
}
