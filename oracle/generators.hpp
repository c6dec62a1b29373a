// ORACLE — TEST INFRASTRUCTURE ONLY (see sparsemat_oracle.hpp).
//
// CPU side of the synthetic workloads of BASELINE.json.  The CUDA library has its
// own device generators (sparsemat_b200/csrc/generators.cu); tests require both to produce the same
// arrays bit for bit, so every formula here is integer or correctly-rounded IEEE arithmetic
// (+, *, /, sqrt, floor) only.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>

namespace oracle {
namespace gen {

// splitmix64 finaliser, used as a counter-based (stateless) generator.
inline std::uint64_t mix(std::uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
inline std::uint64_t rng1(std::uint64_t seed, std::uint64_t i) { return mix(mix(seed) + i); }
inline std::uint64_t rng2(std::uint64_t seed, std::uint64_t i, std::uint64_t k) { return mix(rng1(seed, i) + k); }
inline double u01(std::uint64_t bits) { return static_cast<double>(bits >> 11) * 0x1.0p-53; }   // [0,1)
inline double pm1(std::uint64_t bits) { return 2.0 * u01(bits) - 1.0; }                         // [-1,1)

template <class T> void uniform_pm1(std::uint64_t seed, std::size_t n, T* out) {
    for (std::size_t i = 0; i < n; ++i) out[i] = static_cast<T>(pm1(rng1(seed, i)));
}

// Dirichlet Laplacian stencils, row = (iz*ny + iy)*nx + ix, ascending columns inside a row,
// diagonal = 2*dim, off-diagonals = -1, out-of-grid neighbours skipped.  nz == 1 gives the 2-D
// 5-point operator (C1), nz > 1 the 3-D 7-point one (C2/C4/C5).
inline std::uint64_t laplace_row_len(std::uint64_t nx, std::uint64_t ny, std::uint64_t nz,
                                     std::uint64_t ix, std::uint64_t iy, std::uint64_t iz) {
    std::uint64_t n = 1;
    n += (ix > 0) + (ix + 1 < nx);
    n += (iy > 0) + (iy + 1 < ny);
    if (nz > 1) n += (iz > 0) + (iz + 1 < nz);
    return n;
}
inline std::uint64_t laplace_nnz(std::uint64_t nx, std::uint64_t ny, std::uint64_t nz) {
    std::uint64_t n = nx * ny * nz;                       // diagonal
    n += 2 * (nx - 1) * ny * nz + 2 * nx * (ny - 1) * nz;
    if (nz > 1) n += 2 * nx * ny * (nz - 1);
    return n;
}

// Rows [row_lo, row_hi) of the global operator, offsets rebased so that out_offsets[0] == 0
// (row_lo = 0, row_hi = N gives the whole matrix).  Columns stay global.
template <class T, class I>
void laplace_rows(std::uint64_t nx, std::uint64_t ny, std::uint64_t nz,
                  std::uint64_t row_lo, std::uint64_t row_hi, T* values, I* columns, I* offsets) {
    const T diag = static_cast<T>(nz > 1 ? 6.0 : 4.0);
    const T off = static_cast<T>(-1.0);
    const std::uint64_t plane = nx * ny;
    std::uint64_t k = 0;
    for (std::uint64_t r = row_lo; r < row_hi; ++r) {
        offsets[r - row_lo] = static_cast<I>(k);
        const std::uint64_t ix = r % nx, iy = (r / nx) % ny, iz = r / plane;
        if (nz > 1 && iz > 0)      { columns[k] = static_cast<I>(r - plane); values[k++] = off; }
        if (iy > 0)                { columns[k] = static_cast<I>(r - nx);    values[k++] = off; }
        if (ix > 0)                { columns[k] = static_cast<I>(r - 1);     values[k++] = off; }
        columns[k] = static_cast<I>(r); values[k++] = diag;
        if (ix + 1 < nx)           { columns[k] = static_cast<I>(r + 1);     values[k++] = off; }
        if (iy + 1 < ny)           { columns[k] = static_cast<I>(r + nx);    values[k++] = off; }
        if (nz > 1 && iz + 1 < nz) { columns[k] = static_cast<I>(r + plane); values[k++] = off; }
    }
    offsets[row_hi - row_lo] = static_cast<I>(k);
}

// C3: Pareto(alpha = 2, x_min = 8) row lengths, L = clamp(floor(8 / sqrt(u)), 1, max_len).
inline std::uint64_t powerlaw_row_len(std::uint64_t seed, std::uint64_t i, std::uint64_t max_len) {
    const double u = u01(rng1(seed, i));
    double l = std::floor(8.0 / std::sqrt(u));            // u == 0 -> +inf -> clamped
    if (!(l >= 1.0)) l = 1.0;
    if (l > static_cast<double>(max_len)) l = static_cast<double>(max_len);
    return static_cast<std::uint64_t>(l);
}
// offsets[n_rows + 1] must already hold the exclusive scan of the row lengths.
template <class T, class I>
void powerlaw_fill(std::uint64_t seed_col, std::uint64_t seed_val, std::uint64_t n_rows, std::uint64_t n_cols,
                   const I* offsets, T* values, I* columns) {
    for (std::uint64_t i = 0; i < n_rows; ++i) {
        const std::uint64_t b = static_cast<std::uint64_t>(offsets[i]);
        const std::uint64_t e = static_cast<std::uint64_t>(offsets[i + 1]);
        for (std::uint64_t k = b; k < e; ++k) {
            columns[k] = static_cast<I>(rng2(seed_col, i, k - b) % n_cols);
            values[k] = static_cast<T>(pm1(rng2(seed_val, i, k - b)));
        }
    }
}

// SparseMatrix::mvp (sparsematrix.rs:146-158) restricted to a sample of rows of the IMPLICIT power-law matrix: row i is
// regenerated from the seeds (same formulas as powerlaw_fill), summed in storage order with separate mul and add like the
// reference, and |row|.|x| is returned beside it as the scale of the reordered-reduction tolerance.  Lets the full-size
// C3 product (800 M non-zeros, never held on the host) be checked row by row.
template <class T>
void powerlaw_sample_mvp(std::uint64_t seed_len, std::uint64_t seed_col, std::uint64_t seed_val, std::uint64_t n_cols,
                         std::uint64_t max_len, std::uint64_t n_sample, const std::uint64_t* rows, const T* x, T* y,
                         double* abs_sum, std::uint64_t* lens) {
    for (std::uint64_t s = 0; s < n_sample; ++s) {
        const std::uint64_t i = rows[s];
        const std::uint64_t len = powerlaw_row_len(seed_len, i, max_len);
        T sum = T(0);
        double a = 0.0;
        for (std::uint64_t k = 0; k < len; ++k) {
            const std::uint64_t c = rng2(seed_col, i, k) % n_cols;
            const T v = static_cast<T>(pm1(rng2(seed_val, i, k)));
            const T prod = x[c] * v;
            sum += prod;
            a += std::fabs(static_cast<double>(x[c]) * static_cast<double>(v));
        }
        y[s] = sum;
        if (abs_sum) abs_sum[s] = a;
        if (lens) lens[s] = len;
    }
}

}  // namespace gen
}  // namespace oracle
