/*
 * C restatement ("port") of the reference's CPU path for block-tridiagonal Cholesky + solve.
 *
 * TEST INFRASTRUCTURE / CPU BASELINE ONLY (see oracle/__init__.py): the product never links this.
 *
 * It follows what markovflow does on every call, step by step:
 *   1. BlockTriDiagonal._convert_to_band           (markovflow/block_tri_diag.py:206-237)
 *        blocks [T,D,D] (+[T-1,D,D]) -> lower band [2D, T*D], band[r][j] = M[j+r][j]
 *   2. banded_matrices.cholesky_band               (call site block_tri_diag.py:436)
 *        scalar lower-band Cholesky, the dpbtf2 algorithm class (banded-matrices==0.0.6 is not in
 *        the reference tree; its published algorithm is restated)
 *   3. _banded_to_block_tri                        (block_tri_diag.py:549-592)
 *   4. LowerTriangularBlockTriDiagonal.solve       (block_tri_diag.py:339-351): as_band again
 *        (another block->band repack) + banded_matrices.solve_triang_mat
 * Band tensors are row-major [K, N] like the TensorFlow tensors the reference passes around.
 * Chains are independent; OpenMP threads over chains stand in for TF's inter-op batch threading.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* 1. blocks -> lower band [2D][N] (row-major), N = T*D. */
void ref_block_to_band(const double* diag, const double* sub, double* band, int64_t T, int D) {
  const int64_t N = T * D;
  const int rows = sub ? 2 * D : D;
  memset(band, 0, sizeof(double) * rows * N);
  for (int64_t k = 0; k < T; ++k)
    for (int c = 0; c < D; ++c) {
      const int64_t j = k * D + c;
      for (int i = c; i < D; ++i) band[(int64_t)(i - c) * N + j] = diag[(k * D + i) * D + c];
      if (sub && k + 1 < T)
        for (int i = 0; i < D; ++i) band[(int64_t)(D + i - c) * N + j] = sub[(k * D + i) * D + c];
    }
}

/* 2. in-place scalar lower-band Cholesky; returns 0 or 1-based failing column. */
int64_t ref_cholesky_band(double* band, int64_t N, int KB) {
  for (int64_t j = 0; j < N; ++j) {
    double ajj = band[j];
    if (!(ajj > 0.0)) return j + 1;
    ajj = sqrt(ajj);
    band[j] = ajj;
    const int kn = (int)((N - 1 - j < KB) ? (N - 1 - j) : KB);
    const double inv = 1.0 / ajj;
    for (int a = 1; a <= kn; ++a) band[(int64_t)a * N + j] *= inv;
    for (int b = 1; b <= kn; ++b) {
      const double lb = band[(int64_t)b * N + j];
      for (int a = b; a <= kn; ++a) band[(int64_t)(a - b) * N + j + b] -= band[(int64_t)a * N + j] * lb;
    }
  }
  return 0;
}

/* 3. lower band -> blocks (upper triangles of diagonal blocks are zero). */
void ref_band_to_block(const double* band, double* ld, double* ls, int64_t T, int D) {
  const int64_t N = T * D;
  for (int64_t k = 0; k < T; ++k)
    for (int c = 0; c < D; ++c) {
      const int64_t j = k * D + c;
      for (int i = 0; i < D; ++i)
        ld[(k * D + i) * D + c] = (i >= c) ? band[(int64_t)(i - c) * N + j] : 0.0;
      if (ls && k + 1 < T)
        for (int i = 0; i < D; ++i) ls[(k * D + i) * D + c] = band[(int64_t)(D + i - c) * N + j];
    }
}

/* 4. banded substitution  L x = b  (transpose == 0)  or  L^T x = b. */
void ref_solve_triang_mat(const double* lband, double* x, int64_t N, int KB, int transpose) {
  if (!transpose) {
    for (int64_t i = 0; i < N; ++i) {
      double v = x[i];
      const int kn = (int)(i < KB ? i : KB);
      for (int r = 1; r <= kn; ++r) v -= lband[(int64_t)r * N + i - r] * x[i - r];
      x[i] = v / lband[i];
    }
  } else {
    for (int64_t i = N - 1; i >= 0; --i) {
      double v = x[i];
      const int kn = (int)((N - 1 - i < KB) ? (N - 1 - i) : KB);
      for (int r = 1; r <= kn; ++r) v -= lband[(int64_t)r * N + i] * x[i + r];
      x[i] = v / lband[i];
    }
  }
}

/* The whole reference path for a batch: cholesky (repack, factor, repack) then solve (repack,
 * substitute).  Returns the number of chains that failed; info[b] as in the CUDA ABI (block). */
int64_t ref_chol_solve_batch(const double* diag, const double* sub, const double* rhs, double* ld,
                             double* ls, double* x, int32_t* info, int64_t B, int64_t T, int D,
                             int nthreads) {
  const int64_t N = T * D;
  const int rows = sub ? 2 * D : D;
  const int KB = rows - 1;
  int64_t nfail = 0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel reduction(+ : nfail)
  {
    double* band = (double*)malloc(sizeof(double) * rows * N);
#pragma omp for schedule(static)
    for (int64_t b = 0; b < B; ++b) {
      const double* dg = diag + b * T * D * D;
      const double* sb = sub ? sub + b * (T - 1) * D * D : NULL;
      double* ldb = ld + b * T * D * D;
      double* lsb = ls ? ls + b * (T - 1) * D * D : NULL;
      ref_block_to_band(dg, sb, band, T, D);
      const int64_t f = ref_cholesky_band(band, N, KB);
      if (info) info[b] = f ? (int32_t)((f - 1) / D + 1) : 0;
      if (f) ++nfail;
      ref_band_to_block(band, ldb, lsb, T, D);
      if (rhs) {
        ref_block_to_band(ldb, lsb, band, T, D); /* .solve() re-derives the band (as_band) */
        double* xb = x + b * N;
        memcpy(xb, rhs + b * N, sizeof(double) * N);
        ref_solve_triang_mat(band, xb, N, KB, 0);
      }
    }
    free(band);
  }
  return nfail;
}

int ref_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
