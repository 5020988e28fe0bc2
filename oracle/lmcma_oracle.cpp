// TEST INFRASTRUCTURE ONLY (oracle/).  FP64 CPU restatement of the reference's LM-CMA hot path,
// parameterised in the number of stored direction pairs m (the reference hard-wires m = lambda,
// lmcma.cpp:266-267).  With m == lambda every function below follows the cited reference lines
// operation for operation (same FP64 expression order) and tests/test_oracle_vs_reference.py
// asserts bit-identical state against the compiled reference (oracle/_ref/libref_lmcma.so).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this; the product (lmcma_path_planner_b200/) never does.
//
// Compile: g++ -std=c++11 -O2 -ffp-contract=off -shared -fPIC  (see oracle/Makefile)
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

namespace {

// ---- RNG: Hansen's c-cmaes generator as used by the reference (lmcma.cpp:14-82) -------------
// Park-Miller minimal standard (a=16807, m=2^31-1, Schrage q=127773 r=2836) feeding a 32-entry
// Bays-Durham shuffle table; uniform = shuffled / 2.147483647e9; polar Box-Muller returning
// fac*x2 first and caching fac*x1.
struct HansenRng {
    long seed_state;
    long last;
    long table[32];
    bool have_spare;
    double spare;

    static long lcg(long s) {           // lmcma.cpp:25-27 / 52-56
        long q = s / 127773;
        s = 16807 * (s - q * 127773) - 2836 * q;
        if (s < 0) s += 2147483647;
        return s;
    }
    void start(unsigned long seed) {    // lmcma.cpp:14-33
        have_spare = false;
        if (seed < 1) seed = 1;
        seed_state = static_cast<long>(seed);
        for (int i = 39; i >= 0; --i) {
            seed_state = lcg(seed_state);
            if (i < 32) table[i] = seed_state;
        }
        last = table[0];
    }
    double uniform() {                  // lmcma.cpp:49-61
        seed_state = lcg(seed_state);
        long slot = last / 67108865;
        last = table[slot];
        table[slot] = seed_state;
        return static_cast<double>(last) / 2.147483647e9;
    }
    double gauss() {                    // lmcma.cpp:63-82
        if (have_spare) { have_spare = false; return spare; }
        double a, b, r2;
        do {
            a = 2.0 * uniform() - 1.0;
            b = 2.0 * uniform() - 1.0;
            r2 = a * a + b * b;
        } while (r2 >= 1 || r2 <= 0);
        double fac = std::sqrt(-2.0 * std::log(r2) / r2);
        have_spare = true;
        spare = fac * a;
        return fac * b;
    }
};

// ---- ranking (lmcma.cpp:84-104): ascending, ties keep the lower id first (glibc 2.39 qsort is a
// stable merge sort for these sizes); the sorted values are written back over the input. ---------
struct KeyId { double key; int id; };
void rank_ascending(int count, double* values_inout, int* ids_out) {
    std::vector<KeyId> a(count);
    for (int i = 0; i < count; ++i) { a[i].key = values_inout[i]; a[i].id = i; }
    std::stable_sort(a.begin(), a.end(), [](const KeyId& x, const KeyId& y) { return x.key < y.key; });
    for (int i = 0; i < count; ++i) { values_inout[i] = a[i].key; ids_out[i] = a[i].id; }
}

struct Opt {
    int n, lambda, mu, m;                 // m = nvectors
    int maxsteps, itr, sample_idx, counteval, live;  // live = iterator_sz
    double sigma, s, c1, cc, cs, target, K, M, mueff, best_f;
    bool has_lo, has_hi;
    std::vector<double> lo, hi, X, fit, prev_fit, xmean, xold, w, pc, V, P, Nj, Lj;
    std::vector<int> order /* t */, stamp /* vec */, live_slots /* iterator */, arindex;
    HansenRng rng;

    // lmcma.cpp:431-447 — reconstruct A*z from the stored pairs; every dot is against the ORIGINAL z
    void apply_A(const double* z, double* out) const {
        for (int k = 0; k < n; ++k) out[k] = z[k];
        for (int k = 0; k < live; ++k) {
            const int j = live_slots[k];
            const double* p = &P[static_cast<size_t>(j) * n];
            const double* v = &V[static_cast<size_t>(j) * n];
            double d = 0;
            for (int q = 0; q < n; ++q) d = d + v[q] * z[q];
            d = Nj[j] * d;
            for (int q = 0; q < n; ++q) out[q] = M * out[q] + d * p[q];
        }
    }
    // lmcma.cpp:449-463 — apply the first `upto` inverse factors in place (dots use the running vector)
    void apply_Ainv_prefix(double* a, int upto) const {
        for (int jj = 0; jj < upto; ++jj) {
            const int j = live_slots[jj];
            const double* v = &V[static_cast<size_t>(j) * n];
            double d = 0;
            for (int q = 0; q < n; ++q) d += v[q] * a[q];
            d = Lj[j] * d;
            for (int q = 0; q < n; ++q) a[q] = K * a[q] - d * v[q];
        }
    }
    // lmcma.cpp:301-311 (+212-218, 220-230); Z == nullptr -> draw from the Hansen stream
    void sample(const double* Z) {
        std::vector<double> z(n), az(n);
        for (int i = 0; i < lambda; ++i) {
            if (Z) std::memcpy(z.data(), Z + static_cast<size_t>(i) * n, sizeof(double) * n);
            else for (int k = 0; k < n; ++k) z[k] = rng.gauss();
            apply_A(z.data(), az.data());
            for (int k = 0; k < n; ++k) X[static_cast<size_t>(i) * n + k] = xmean[k] + sigma * az[k];
        }
        if (has_lo)
            for (int i = 0; i < lambda; ++i)
                for (int k = 0; k < n; ++k) {
                    double& x = X[static_cast<size_t>(i) * n + k];
                    x = std::max(x, lo[k]);
                }
        if (has_hi)
            for (int i = 0; i < lambda; ++i)
                for (int k = 0; k < n; ++k) {
                    double& x = X[static_cast<size_t>(i) * n + k];
                    x = std::min(x, hi[k]);
                }
    }
    // lmcma.cpp:313-424
    void update() {
        rank_ascending(lambda, fit.data(), arindex.data());              // :315
        for (int k = 0; k < n; ++k) { xold[k] = xmean[k]; xmean[k] = 0; }  // :316-320
        for (int i = 0; i < mu; ++i) {                                    // :321-326
            const double* row = &X[static_cast<size_t>(arindex[i]) * n];
            for (int k = 0; k < n; ++k) xmean[k] += w[i] * row[k];
        }
        for (int k = 0; k < n; ++k)                                       // :327-329
            pc[k] = (1 - cc) * pc[k] + std::sqrt(cc * (2 - cc) * mueff) * (xmean[k] - xold[k]) / sigma;

        // slot bookkeeping (:331-366): `order` lists slots oldest->newest, `stamp` = generation a slot
        // was written.  Once full, retire the later member of the closest-in-time adjacent pair
        // (first such pair wins), or the oldest slot when every gap is >= maxsteps.
        int first_stale = 1;
        if (itr < m) {
            order[itr] = itr;
        } else {
            int gap_min = stamp[order[1]] - stamp[order[0]];
            for (int j = 1; j < m - 1; ++j) {
                int gap = stamp[order[j + 1]] - stamp[order[j]];
                if (gap < gap_min) { gap_min = gap; first_stale = j + 1; }
            }
            if (gap_min >= maxsteps) first_stale = 0;
            if (first_stale != m - 1) {
                int recycled = order[first_stale];
                for (int j = first_stale; j < m - 1; ++j) order[j] = order[j + 1];
                order[m - 1] = recycled;
            }
        }
        live = std::min(itr + 1, m);                                      // :357-359
        for (int i = 0; i < live; ++i) live_slots[i] = order[i];          // :360-361
        const int slot_new = order[live - 1];                             // :362
        stamp[slot_new] = itr;                                            // :364
        for (int k = 0; k < n; ++k) P[static_cast<size_t>(slot_new) * n + k] = pc[k];  // :365-366

        // recompute the inverse-direction vectors from the first stale position (:373-390)
        if (first_stale == 1) first_stale = 0;
        std::vector<double> a(n);
        for (int i = first_stale; i < live; ++i) {
            const int slot = order[i];
            for (int k = 0; k < n; ++k) a[k] = P[static_cast<size_t>(slot) * n + k];
            apply_Ainv_prefix(a.data(), i);
            double* v = &V[static_cast<size_t>(slot) * n];
            for (int k = 0; k < n; ++k) v[k] = a[k];
            double nv = 0;
            for (int k = 0; k < n; ++k) nv += v[k] * v[k];
            Nj[slot] = (std::sqrt(1 - c1) / nv) * (std::sqrt(1 + (c1 / (1 - c1)) * nv) - 1);
            Lj[slot] = (1 / (std::sqrt(1 - c1) * nv)) * (1 - (1 / std::sqrt(1 + (c1 / (1 - c1)) * nv)));
        }

        // population-success step size (:393-419): rank the union of this and the previous (both
        // already sorted) generation; ties favour the current one (lower index, stable sort).
        if (itr > 0) {
            std::vector<double> both(2 * lambda);
            std::vector<int> ids(2 * lambda), pos(2 * lambda);
            for (int i = 0; i < lambda; ++i) { both[i] = fit[i]; both[lambda + i] = prev_fit[i]; }
            rank_ascending(2 * lambda, both.data(), ids.data());
            for (int i = 0; i < 2 * lambda; ++i) pos[ids[i]] = i;
            double sum_cur = 0, sum_prev = 0;   // reference names them meanprev / meancur (:401-411)
            for (int i = 0; i < lambda; ++i) { sum_cur = sum_cur + pos[i]; sum_prev = sum_prev + pos[lambda + i]; }
            sum_cur = sum_cur / lambda;
            sum_prev = sum_prev / lambda;
            double success = (sum_prev - sum_cur) / lambda;
            double z1 = success - target;
            s = (1 - cs) * s + cs * z1;
            double d_s = 1;
            sigma = sigma * std::exp(s / d_s);
        }
        for (int i = 0; i < lambda; ++i) prev_fit[i] = fit[i];            // :420-421
        itr++;
    }
    // lmcma.cpp:184-205
    void tell_one(double f, const double* Z_next) {
        fit[sample_idx] = f;
        counteval++;
        if (f < best_f || counteval == 1) best_f = f;
        if (++sample_idx % lambda == 0) {
            update();
            sample(Z_next);
            sample_idx = 0;
        }
    }
};

}  // namespace

extern "C" {

// x0 == NULL -> uniform start from the RNG (lmcma.cpp:161-163).  lambda < 1 -> 4 + int(3 ln n)
// (lmcma.cpp:134-135).  m < 1 -> m = lambda (the reference's rule, lmcma.cpp:266).
// Z0 != NULL -> first population built from these lambda*n deviates instead of the RNG stream.
void* orc_lmcma_create(int n, int lambda, int m, const double* x0, const double* lo, const double* hi,
                       double sigma, long seed, const double* Z0) {
    Opt* o = new Opt();
    o->rng.start(static_cast<unsigned long>(seed));
    o->n = n;
    if (lambda < 1) lambda = 4 + int(3 * std::log(n));
    o->lambda = lambda;
    o->mu = lambda / 2;                                                  // lmcma.cpp:136
    o->itr = 0; o->sample_idx = 0; o->counteval = 0;
    o->best_f = std::numeric_limits<double>::max();
    o->sigma = sigma; o->cs = 0.3; o->target = 0.25;                      // lmcma.cpp:238
    o->X.assign(static_cast<size_t>(n) * lambda, 0.0);
    o->fit.assign(lambda, 0.0); o->prev_fit.assign(lambda, 0.0);
    o->xmean.assign(n, 0.0); o->xold.assign(n, 0.0); o->w.assign(o->mu, 0.0);
    double sw = 0;                                                       // lmcma.cpp:144-156
    for (int i = 0; i < o->mu; ++i) { o->w[i] = std::log(double(o->mu) + 0.5) - std::log(double(1 + i)); sw += o->w[i]; }
    o->mueff = 0.0;
    for (int i = 0; i < o->mu; ++i) { o->w[i] /= sw; o->mueff += o->w[i] * o->w[i]; }
    o->mueff = 1.0 / o->mueff;
    if (x0) for (int k = 0; k < n; ++k) o->xmean[k] = x0[k];             // lmcma.cpp:158-163
    else for (int k = 0; k < n; ++k) o->xmean[k] = o->rng.uniform();
    o->has_lo = lo != 0; o->has_hi = hi != 0;
    if (lo) o->lo.assign(lo, lo + n);
    if (hi) o->hi.assign(hi, hi + n);
    if (m < 1) m = lambda;                                               // lmcma.cpp:266-272
    o->m = m; o->maxsteps = m;
    o->c1 = 1.0 / (10 * std::log(n + 1));
    o->cc = 1.0 / m;
    o->K = 1 / std::sqrt(1 - o->c1);
    o->M = std::sqrt(1 - o->c1);
    o->V.assign(static_cast<size_t>(n) * m, 0.0); o->P.assign(static_cast<size_t>(n) * m, 0.0);
    o->pc.assign(n, 0.0); o->Nj.assign(m, 0.0); o->Lj.assign(m, 0.0);
    o->order.assign(m, 0); o->stamp.assign(m, 0); o->live_slots.assign(m, 0); o->arindex.assign(lambda, 0);
    o->s = 0.0; o->live = 0;
    o->sample(Z0);                                                       // lmcma.cpp:298
    return o;
}
void orc_lmcma_destroy(void* h) { delete static_cast<Opt*>(h); }
void orc_lmcma_ask(void* h, double* params) {                            // lmcma.cpp:172-182
    Opt* o = static_cast<Opt*>(h);
    std::memcpy(params, &o->X[static_cast<size_t>(o->sample_idx) * o->n], sizeof(double) * o->n);
}
void orc_lmcma_tell(void* h, double f) { static_cast<Opt*>(h)->tell_one(f, 0); }
// whole-generation tell in candidate order; Z_next (nullable) feeds the sample() that follows.
void orc_lmcma_tell_all(void* h, const double* f, const double* Z_next) {
    Opt* o = static_cast<Opt*>(h);
    for (int i = 0; i < o->lambda; ++i) o->tell_one(f[i], i == o->lambda - 1 ? Z_next : 0);
}
int orc_lmcma_done(void* h) { return static_cast<Opt*>(h)->sigma < 1e-20 ? 1 : 0; }   // lmcma.cpp:426-429

// same slot numbering as oracle/ref_harness.cpp
void orc_lmcma_get_ints(void* h, int* out) {
    Opt* o = static_cast<Opt*>(h);
    out[0] = o->n; out[1] = o->lambda; out[2] = o->mu; out[3] = o->itr; out[4] = o->sample_idx;
    out[5] = o->counteval; out[6] = o->m; out[7] = o->maxsteps; out[8] = o->live;
}
void orc_lmcma_get_doubles(void* h, double* out) {
    Opt* o = static_cast<Opt*>(h);
    out[0] = o->sigma; out[1] = o->s; out[2] = o->c1; out[3] = o->cc; out[4] = o->cs;
    out[5] = o->target; out[6] = o->K; out[7] = o->M; out[8] = o->mueff; out[9] = o->best_f;
}
int orc_lmcma_get_array(void* h, int which, double* out) {
    Opt* o = static_cast<Opt*>(h);
    const std::vector<double>* src = 0;
    switch (which) {
        case 0: src = &o->xmean; break;  case 1: src = &o->xold; break;  case 2: src = &o->pc; break;
        case 3: src = &o->V; break;      case 4: src = &o->P; break;     case 5: src = &o->Nj; break;
        case 6: src = &o->Lj; break;     case 7: src = &o->X; break;     case 8: src = &o->fit; break;
        case 9: src = &o->prev_fit; break; case 10: src = &o->w; break;
        default: return -1;
    }
    std::memcpy(out, src->data(), src->size() * sizeof(double));
    return static_cast<int>(src->size());
}
int orc_lmcma_get_int_array(void* h, int which, int* out) {
    Opt* o = static_cast<Opt*>(h);
    const std::vector<int>* src = 0;
    switch (which) {
        case 0: src = &o->order; break; case 1: src = &o->stamp; break;
        case 2: src = &o->live_slots; break; case 3: src = &o->arindex; break;
        default: return -1;
    }
    std::memcpy(out, src->data(), src->size() * sizeof(int));
    return static_cast<int>(src->size());
}
// Warm start (tests only): overwrite the distribution state, then sample() with Z (nullable -> Hansen stream).  Lets a
// large-population oracle (C4: lambda = 8192) start from a state in which all m slots are live and being recycled — a
// state produced by a cheap small-lambda run of this same restatement — instead of paying ~m full-size generations.
// prev_fit: lambda values, ascending (what update() leaves behind, lmcma.cpp:99-103, 420-421).
void orc_lmcma_set_state(void* h, const double* xmean, const double* pc, const double* V, const double* P, const double* Nj,
                         const double* Lj, const double* prev_fit, const int* order, const int* stamp, int itr, int live,
                         double sigma, double s, const double* Z) {
    Opt* o = static_cast<Opt*>(h);
    const size_t n = o->n, m = o->m;
    std::memcpy(o->xmean.data(), xmean, n * sizeof(double));
    std::memcpy(o->pc.data(), pc, n * sizeof(double));
    std::memcpy(o->V.data(), V, n * m * sizeof(double));
    std::memcpy(o->P.data(), P, n * m * sizeof(double));
    std::memcpy(o->Nj.data(), Nj, m * sizeof(double));
    std::memcpy(o->Lj.data(), Lj, m * sizeof(double));
    std::memcpy(o->prev_fit.data(), prev_fit, o->lambda * sizeof(double));
    std::memcpy(o->order.data(), order, m * sizeof(int));
    std::memcpy(o->stamp.data(), stamp, m * sizeof(int));
    o->itr = itr; o->live = live; o->sigma = sigma; o->s = s;
    for (int i = 0; i < live; ++i) o->live_slots[i] = o->order[i];
    o->sample_idx = 0;
    o->sample(Z);
}

void orc_rng_uniform(long seed, int count, double* out) {
    HansenRng r; r.start(static_cast<unsigned long>(seed));
    for (int i = 0; i < count; ++i) out[i] = r.uniform();
}
void orc_rng_gauss(long seed, long skip, long count, double* out) {
    HansenRng r; r.start(static_cast<unsigned long>(seed));
    for (long i = 0; i < skip; ++i) (void)r.gauss();
    for (long i = 0; i < count; ++i) out[i] = r.gauss();
}
void orc_rank(int count, double* values_inout, int* ids_out) { rank_ascending(count, values_inout, ids_out); }

}  // extern "C"
