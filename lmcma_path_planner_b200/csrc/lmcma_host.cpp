// lmcma_host.cpp — see lmcma_host.hpp
#include "lmcma_host.hpp"

#include <algorithm>
#include <cstring>

namespace lmcma {

bool smoothness_covariance(int dims, int waypoints, double* out) {
    const int W = waypoints, n = dims * waypoints;
    // centred 7-tap acceleration rule of the reference (DIFF_RULES[2], lmcma.cpp:763), dt = 1
    static const double taps[7] = {0.0, -1.0 / 12.0, 16.0 / 12.0, -30.0 / 12.0, 16.0 / 12.0, -1.0 / 12.0, 0.0};
    std::vector<double> A(static_cast<size_t>(W) * W, 0.0), AA(static_cast<size_t>(W) * W, 0.0);
    for (int i = 0; i < W; ++i)
        for (int j = -3; j <= 3; ++j) {
            const int c = i + j;
            if (c < 0 || c >= W) continue;          // truncated at the trajectory ends (lmcma.cpp:826-829)
            A[static_cast<size_t>(i) * W + c] += taps[j + 3];
        }
    for (int i = 0; i < W; ++i)                      // A*A (lmcma.cpp:793-798), banded so skip zeros
        for (int k = std::max(0, i - 3); k <= std::min(W - 1, i + 3); ++k) {
            const double a = A[static_cast<size_t>(i) * W + k];
            if (a == 0.0) continue;
            for (int j = std::max(0, k - 3); j <= std::min(W - 1, k + 3); ++j)
                AA[static_cast<size_t>(i) * W + j] += a * A[static_cast<size_t>(k) * W + j];
        }
    // inverse of one block by Gauss-Jordan with partial pivoting
    std::vector<double> inv(static_cast<size_t>(W) * W, 0.0);
    for (int i = 0; i < W; ++i) inv[static_cast<size_t>(i) * W + i] = 1.0;
    for (int col = 0; col < W; ++col) {
        int piv = col;
        for (int r = col + 1; r < W; ++r)
            if (std::fabs(AA[static_cast<size_t>(r) * W + col]) > std::fabs(AA[static_cast<size_t>(piv) * W + col])) piv = r;
        if (AA[static_cast<size_t>(piv) * W + col] == 0.0) return false;
        if (piv != col)
            for (int j = 0; j < W; ++j) {
                std::swap(AA[static_cast<size_t>(col) * W + j], AA[static_cast<size_t>(piv) * W + j]);
                std::swap(inv[static_cast<size_t>(col) * W + j], inv[static_cast<size_t>(piv) * W + j]);
            }
        const double p = AA[static_cast<size_t>(col) * W + col];
        for (int j = 0; j < W; ++j) { AA[static_cast<size_t>(col) * W + j] /= p; inv[static_cast<size_t>(col) * W + j] /= p; }
        for (int r = 0; r < W; ++r) {
            if (r == col) continue;
            const double f = AA[static_cast<size_t>(r) * W + col];
            if (f == 0.0) continue;
            for (int j = 0; j < W; ++j) {
                AA[static_cast<size_t>(r) * W + j] -= f * AA[static_cast<size_t>(col) * W + j];
                inv[static_cast<size_t>(r) * W + j] -= f * inv[static_cast<size_t>(col) * W + j];
            }
        }
    }
    double vmax = 0.0;                                // lmcma.cpp:802-806
    for (int i = 0; i < W; ++i) vmax = std::max(vmax, inv[static_cast<size_t>(i) * W + i]);
    const double scaling = vmax * W;
    std::memset(out, 0, sizeof(double) * static_cast<size_t>(n) * n);
    for (int d = 0; d < dims; ++d)                    // dimension-major blocks (lmcma.cpp:786-791)
        for (int i = 0; i < W; ++i)
            for (int j = 0; j < W; ++j)
                out[static_cast<size_t>(d * W + i) * n + (d * W + j)] = inv[static_cast<size_t>(i) * W + j] / scaling;
    return true;
}

bool cholesky_lower(const double* C, int n, double* L) {
    std::memset(L, 0, sizeof(double) * static_cast<size_t>(n) * n);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j) {                      // Cholesky-Banachiewicz, row by row
            double s = C[static_cast<size_t>(i) * n + j];
            for (int k = 0; k < j; ++k) s -= L[static_cast<size_t>(i) * n + k] * L[static_cast<size_t>(j) * n + k];
            if (i == j) {
                if (!(s > 0.0)) return false;
                L[static_cast<size_t>(i) * n + i] = std::sqrt(s);
            } else {
                L[static_cast<size_t>(i) * n + j] = s / L[static_cast<size_t>(j) * n + j];
            }
        }
    return true;
}

bool differentiation_matrix(int steps, int order, double dt, double* out, int row_len) {
    if (order < 0 || order > 3 || steps < 1) return false;
    if (row_len < 0) row_len = steps;
    // centred finite-difference rules, 7 taps (lmcma.cpp:759-764)
    static const double rules[4][7] = {
        {0, 0, 0, 1, 0, 0, 0},
        {0, 0, -1, 1, 0, 0, 0},
        {0, -1 / 12.0, 16 / 12.0, -30 / 12.0, 16 / 12.0, -1 / 12.0, 0},
        {0, 1 / 12.0, -17 / 12.0, 46 / 12.0, -46 / 12.0, 17 / 12.0, -1 / 12.0}};
    const double mult = 1.0 / std::pow(dt, order);
    for (int i = 0; i < steps; ++i) {
        double* row = out + static_cast<size_t>(i) * row_len;
        std::fill(row, row + steps, 0.0);
        for (int j = -3; j <= 3; ++j) {
            const int c = i + j;
            if (c >= 0 && c < steps) row[c] += mult * rules[order][j + 3];
        }
    }
    return true;
}

bool invert_dense(const double* A, int n, double* Ainv) {
    std::vector<double> a(A, A + static_cast<size_t>(n) * n);
    std::fill(Ainv, Ainv + static_cast<size_t>(n) * n, 0.0);
    for (int i = 0; i < n; ++i) Ainv[static_cast<size_t>(i) * n + i] = 1.0;
    for (int col = 0; col < n; ++col) {
        int piv = col;
        for (int r = col + 1; r < n; ++r)
            if (std::fabs(a[static_cast<size_t>(r) * n + col]) > std::fabs(a[static_cast<size_t>(piv) * n + col])) piv = r;
        const double p = a[static_cast<size_t>(piv) * n + col];
        if (p == 0.0) return false;
        if (piv != col)
            for (int j = 0; j < n; ++j) {
                std::swap(a[static_cast<size_t>(col) * n + j], a[static_cast<size_t>(piv) * n + j]);
                std::swap(Ainv[static_cast<size_t>(col) * n + j], Ainv[static_cast<size_t>(piv) * n + j]);
            }
        for (int j = 0; j < n; ++j) { a[static_cast<size_t>(col) * n + j] /= p; Ainv[static_cast<size_t>(col) * n + j] /= p; }
        for (int r = 0; r < n; ++r) {
            const double f = a[static_cast<size_t>(r) * n + col];
            if (r == col || f == 0.0) continue;
            for (int j = 0; j < n; ++j) {
                a[static_cast<size_t>(r) * n + j] -= f * a[static_cast<size_t>(col) * n + j];
                Ainv[static_cast<size_t>(r) * n + j] -= f * Ainv[static_cast<size_t>(col) * n + j];
            }
        }
    }
    return true;
}

void stable_rank(int count, double* values_inout, int* ids_out) {
    std::vector<int> ids(count);
    for (int i = 0; i < count; ++i) ids[i] = i;
    std::stable_sort(ids.begin(), ids.end(), [&](int x, int y) { return values_inout[x] < values_inout[y]; });
    std::vector<double> sorted(count);
    for (int i = 0; i < count; ++i) sorted[i] = values_inout[ids[i]];
    for (int i = 0; i < count; ++i) { values_inout[i] = sorted[i]; ids_out[i] = ids[i]; }
}

}  // namespace lmcma
