// lmcma_host.hpp — host-side pieces of the optimiser that are inherently serial and tiny:
// the reference's random stream (so a seeded drop-in run draws the reference's deviates) and the
// smoothness-prior covariance builder.  Pure C++, no CUDA.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

namespace lmcma {

// Serial generator with the reference's stream (random_Start / random_Uniform / random_Gauss,
// lmcma.cpp:14-82): Park-Miller "minimal standard" LCG advanced with Schrage's decomposition,
// decorrelated through a 32-entry shuffle table, and a polar Box-Muller that hands out the second
// deviate of each pair first.
class HansenStream {
public:
    explicit HansenStream(int64_t seed = 1) { reseed(seed); }
    void reseed(int64_t seed) {
        spare_valid_ = false;
        state_ = seed < 1 ? 1 : seed;
        for (int i = 39; i >= 0; --i) {
            advance();
            if (i < 32) table_[i] = state_;
        }
        out_ = table_[0];
    }
    double uniform() {
        advance();
        const int64_t slot = out_ / 67108865;   // 2^31 / 32, rounded up
        out_ = table_[slot];
        table_[slot] = state_;
        return static_cast<double>(out_) / 2.147483647e9;
    }
    double gauss() {
        if (spare_valid_) { spare_valid_ = false; return spare_; }
        double u, v, r2;
        do {
            u = 2.0 * uniform() - 1.0;
            v = 2.0 * uniform() - 1.0;
            r2 = u * u + v * v;
        } while (r2 >= 1 || r2 <= 0);
        const double f = std::sqrt(-2.0 * std::log(r2) / r2);
        spare_ = f * u;
        spare_valid_ = true;
        return f * v;
    }

private:
    void advance() {
        const int64_t q = state_ / 127773;
        state_ = 16807 * (state_ - q * 127773) - 2836 * q;
        if (state_ < 0) state_ += 2147483647;
    }
    int64_t state_, out_, table_[32];
    bool spare_valid_;
    double spare_;
};

// covariance(num_dims, num_waypoints, out) of the reference (lmcma.cpp:769-810) on the heap:
// block-diagonal acceleration finite-difference operator A (stencil lmcma.cpp:763), C = inv(A*A),
// scaled by max(diag C) * num_waypoints.  Works per dimension block (all blocks are identical), so
// the cost is O(W^3) instead of the reference's O((D*W)^3) with 2*(D*W)^2 doubles on the stack.
// Returns false if a block is singular.
bool smoothness_covariance(int dims, int waypoints, double* out);

// Lower Cholesky factor of a symmetric positive definite n x n matrix (cholesky(), lmcma.cpp:844-855, which takes
// Eigen's LLT): L row-major, zero above the diagonal.  Returns false if C is not positive definite.
bool cholesky_lower(const double* C, int n, double* L);

// differentiationMatrix() of the reference (lmcma.cpp:812-834): centred 7-tap finite-difference operator of the given
// order (0 position, 1 velocity, 2 acceleration, 3 jerk; rules lmcma.cpp:759-764) for `steps` time steps, truncated at
// the ends, written into the top-left steps x steps block of a row-major matrix with row stride `row_len`.
bool differentiation_matrix(int steps, int order, double dt, double* out, int row_len);

// invert() of the reference (lmcma.cpp:836-842, Eigen's dense inverse there): Gauss-Jordan with partial pivoting, FP64.
// Returns false for a singular matrix.  A and Ainv may not alias.
bool invert_dense(const double* A, int n, double* Ainv);

// myqsort() of the reference (lmcma.cpp:93-104): ascending, ties keep the lower id (glibc's qsort is a stable merge sort
// for these sizes), the sorted values are written back over the input.
void stable_rank(int count, double* values_inout, int* ids_out);

}  // namespace lmcma
