// TEST INFRASTRUCTURE ONLY (oracle/): extern "C" peep-holes into the UNMODIFIED reference
// optimiser, compiled from where it lies (/root/reference/lmcma_path_planner/src/lmcma.{hpp,cpp})
// by oracle/Makefile into oracle/_ref/libref_lmcma.so.  Built with -fno-access-control so the
// private state of `LMCMA` (lmcma.hpp:91-118) is readable.  Nothing here re-implements the
// algorithm: every function forwards to a reference symbol or copies a reference field out.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library.
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <vector>
#include "lmcma.hpp"   // -I/root/reference/lmcma_path_planner/src

namespace {
struct RefOpt {
    // the reference borrows these pointers for its whole life (lmcma.cpp:109-110, 222-229)
    std::vector<double> x0, lo, hi, cov;
    LMCMA* opt;
    int n;
};
}  // namespace

extern "C" {

void* ref_lmcma_create(const double* x0, int n, int lambda, const double* lo, const double* hi,
                       double sigma, const double* cov, int seed) {
    RefOpt* r = new RefOpt();
    r->n = n;
    if (x0) r->x0.assign(x0, x0 + n);
    if (lo) r->lo.assign(lo, lo + n);
    if (hi) r->hi.assign(hi, hi + n);
    if (cov) r->cov.assign(cov, cov + static_cast<size_t>(n) * n);
    r->opt = new LMCMA(x0 ? r->x0.data() : 0, lambda, lo ? r->lo.data() : 0, hi ? r->hi.data() : 0,
                       sigma, cov ? r->cov.data() : 0, seed, false);
    r->opt->init(n);
    return r;
}
void ref_lmcma_destroy(void* h) {
    RefOpt* r = static_cast<RefOpt*>(h);
    delete r->opt;
    delete r;
}
void ref_lmcma_ask(void* h, double* params) {
    RefOpt* r = static_cast<RefOpt*>(h);
    r->opt->getNextParameterVector(params, r->n);
}
void ref_lmcma_tell(void* h, double f) {
    static_cast<RefOpt*>(h)->opt->setEvaluationFeedback(&f, 1);
}
int ref_lmcma_done(void* h) { return static_cast<RefOpt*>(h)->opt->isBehaviorLearningDone() ? 1 : 0; }

// ints: [N, lambda, mu, itr, sampleIdx, counteval, nvectors, maxsteps, iterator_sz]
void ref_lmcma_get_ints(void* h, int* out) {
    LMCMA* o = static_cast<RefOpt*>(h)->opt;
    out[0] = o->N; out[1] = o->lambda; out[2] = o->mu; out[3] = o->itr; out[4] = o->sampleIdx;
    out[5] = o->counteval; out[6] = o->nvectors; out[7] = o->maxsteps; out[8] = o->iterator_sz;
}
// doubles: [sigma, s, c1, cc, cs, val_target, K, M, mueff, BestF]
void ref_lmcma_get_doubles(void* h, double* out) {
    LMCMA* o = static_cast<RefOpt*>(h)->opt;
    out[0] = o->sigma; out[1] = o->s; out[2] = o->c1; out[3] = o->cc; out[4] = o->cs;
    out[5] = o->val_target; out[6] = o->K; out[7] = o->M; out[8] = o->mueff; out[9] = o->BestF;
}
// which: 0 xmean[N] 1 xold[N] 2 pc[N] 3 v_arr[nvectors*N] 4 pc_arr[nvectors*N] 5 Nj[nvectors]
//        6 Lj[nvectors] 7 arx[lambda*N] 8 arfitness[lambda] 9 prev_arfitness[lambda] 10 weights[mu]
int ref_lmcma_get_array(void* h, int which, double* out) {
    LMCMA* o = static_cast<RefOpt*>(h)->opt;
    const double* src = 0; size_t cnt = 0;
    const size_t N = o->N, nv = o->nvectors, lam = o->lambda;
    switch (which) {
        case 0: src = o->xmean; cnt = N; break;
        case 1: src = o->xold; cnt = N; break;
        case 2: src = o->pc; cnt = N; break;
        case 3: src = o->v_arr; cnt = nv * N; break;
        case 4: src = o->pc_arr; cnt = nv * N; break;
        case 5: src = o->Nj_arr; cnt = nv; break;
        case 6: src = o->Lj_arr; cnt = nv; break;
        case 7: src = o->arx; cnt = lam * N; break;
        case 8: src = o->arfitness; cnt = lam; break;
        case 9: src = o->prev_arfitness; cnt = lam; break;
        case 10: src = o->weights; cnt = o->mu; break;
        default: return -1;
    }
    std::memcpy(out, src, cnt * sizeof(double));
    return static_cast<int>(cnt);
}
// which: 0 t[nvectors] 1 vec[nvectors] 2 iterator[nvectors] 3 arindex[lambda]
int ref_lmcma_get_int_array(void* h, int which, int* out) {
    LMCMA* o = static_cast<RefOpt*>(h)->opt;
    const int* src = 0; size_t cnt = 0;
    switch (which) {
        case 0: src = o->t; cnt = o->nvectors; break;
        case 1: src = o->vec; cnt = o->nvectors; break;
        case 2: src = o->iterator; cnt = o->nvectors; break;
        case 3: src = o->arindex; cnt = o->lambda; break;
        default: return -1;
    }
    std::memcpy(out, src, cnt * sizeof(int));
    return static_cast<int>(cnt);
}

// One reference generation driven through the reference's own one-candidate-at-a-time protocol
// (example_lmcma.cpp:49-55): ask, hand the whole population to `cost` (caller-supplied batch
// evaluator, may be threaded), tell in order.  X_out (lambda*N doubles) receives the candidates.
typedef void (*ref_batch_cost_fn)(const double* X, int count, int n, double* f, void* ctx);
void ref_lmcma_generation(void* h, ref_batch_cost_fn cost, void* ctx, double* X_out, double* f_out) {
    RefOpt* r = static_cast<RefOpt*>(h);
    LMCMA* o = r->opt;
    const int lam = o->lambda, n = r->n;
    // the population of a generation is fixed once sample() ran (lmcma.cpp:199-204), so reading
    // all rows before the first tell is the same as interleaving ask/tell.
    std::memcpy(X_out, o->arx, sizeof(double) * static_cast<size_t>(lam) * n);
    cost(X_out, lam, n, f_out, ctx);
    std::vector<double> row(n);
    for (int i = 0; i < lam; ++i) {
        o->getNextParameterVector(row.data(), n);
        double f = f_out[i];
        o->setEvaluationFeedback(&f, 1);
    }
}

// ---- reference RNG (lmcma.cpp:9-82) ----
void ref_rng_uniform(long seed, int count, double* out) {
    random_t r; random_init(&r, seed);
    for (int i = 0; i < count; ++i) out[i] = random_Uniform(&r);
    random_exit(&r);
}
void ref_rng_gauss(long seed, long skip, long count, double* out) {
    random_t r; random_init(&r, seed);
    for (long i = 0; i < skip; ++i) (void)random_Gauss(&r);
    for (long i = 0; i < count; ++i) out[i] = random_Gauss(&r);
    random_exit(&r);
}
// ---- reference ranking (lmcma.cpp:84-104) ----
void ref_myqsort(int sz, double* fitness_inout, int* index_out) {
    std::vector<sortedvals> tmp(sz);
    myqsort(sz, fitness_inout, index_out, tmp.data());
}
// ---- reference smoothness prior (lmcma.cpp:769-864), shim-pinned ----
void ref_covariance(int dims, int waypoints, double* out) { covariance(dims, waypoints, out); }
void ref_cholesky(double* C, double* L, int n) { cholesky(C, L, n); }
void ref_apply_cov_l(double* L, double* z, int n) { applyCovL(L, z, n); }

// ---- sibling optimisers of the reference (SepCMA lmcma.hpp:152-205, CMAChol lmcma.hpp:211-254) ----
// CPU cross-checks of solution quality only: never GPU targets.  kind: 0 LMCMA, 1 SepCMA, 2 CMAChol,
// driven through the shared CMABase protocol (lmcma.cpp:172-205).
struct RefSibling {
    std::vector<double> x0, lo, hi;
    CMABase* opt;
    int n;
};
void* ref_sibling_create(int kind, const double* x0, int n, int lambda, const double* lo, const double* hi,
                         double sigma, int seed) {
    if (kind < 0 || kind > 2) return 0;
    RefSibling* r = new RefSibling();
    r->n = n;
    if (x0) r->x0.assign(x0, x0 + n);
    if (lo) r->lo.assign(lo, lo + n);
    if (hi) r->hi.assign(hi, hi + n);
    double* px = x0 ? r->x0.data() : 0;
    double* pl = lo ? r->lo.data() : 0;
    double* ph = hi ? r->hi.data() : 0;
    if (kind == 0) r->opt = new LMCMA(px, lambda, pl, ph, sigma, 0, seed, false);
    else if (kind == 1) r->opt = new SepCMA(px, lambda, pl, ph, sigma, 0, seed, false);
    else r->opt = new CMAChol(px, lambda, pl, ph, sigma, 0, seed, false);
    r->opt->init(n);
    return r;
}
void ref_sibling_destroy(void* h) {
    RefSibling* r = static_cast<RefSibling*>(h);
    delete r->opt;
    delete r;
}
int ref_sibling_lambda(void* h) { return static_cast<RefSibling*>(h)->opt->lambda; }
// `generations` full generations of ask / tell with the caller's batch cost; best_x receives the best candidate
// seen (strict <, first evaluation included - the BestF rule of lmcma.cpp:190-193).  Returns BestF.
double ref_sibling_run(void* h, ref_batch_cost_fn cost, void* ctx, int generations, double* best_x) {
    RefSibling* r = static_cast<RefSibling*>(h);
    CMABase* o = r->opt;
    const int lam = o->lambda, n = r->n;
    std::vector<double> row(n);
    double best = 0.0; bool have = false;
    for (int g = 0; g < generations && !o->isBehaviorLearningDone(); ++g)
        for (int i = 0; i < lam; ++i) {
            o->getNextParameterVector(row.data(), n);
            double f = 0.0;
            cost(row.data(), 1, n, &f, ctx);
            if (!have || f < best) { best = f; have = true; if (best_x) std::memcpy(best_x, row.data(), sizeof(double) * n); }
            o->setEvaluationFeedback(&f, 1);
        }
    return o->BestF;
}

}  // extern "C"
