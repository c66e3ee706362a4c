/*
 * C restatement ("port") of the reference's CPU path for the state-space-model operators that the
 * other BASELINE configurations time: KalmanFilter.log_likelihood (configs 1, 3),
 * naturals_to_ssm_params and ssm_to_expectations (config 5).
 *
 * TEST INFRASTRUCTURE / CPU BASELINE ONLY (see oracle/__init__.py): the product never links this.
 *
 * Each function walks the SAME sequence of operations the reference issues (TensorFlow batched small
 * linear algebra + the banded ops of banded-matrices==0.0.6 with their block<->band repacks), per chain,
 * with OpenMP over chains standing in for TF's batch threading:
 *
 *   ref_kalman_loglik_batch        markovflow/kalman_filter.py:184-255  (SpInGP form: _k_inv_post :85-101,
 *                                  StateSpaceModel._build_precision state_space_model.py:431-483,
 *                                  marginal_means :231-251, log_det_precision :343-373,
 *                                  _back_project_y_to_state kalman_filter.py:257-271, KalmanFilter._r_inv :341-348)
 *   ref_nat_to_ssm_batch           markovflow/ssm_gaussian_transformations.py:332-511
 *   ref_ssm_to_expectations_batch  markovflow/ssm_gaussian_transformations.py:31-89
 *                                  (marginal_covariances = precision.cholesky.block_diagonal_of_inverse(),
 *                                  state_space_model.py:253-262)
 *
 * Checked against oracle/np_oracle.py in tests/test_capi_and_cport.py.
 */
#define _USE_MATH_DEFINES
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXD 32

/* from banded_ref.c */
void ref_block_to_band(const double* diag, const double* sub, double* band, int64_t T, int D);
int64_t ref_cholesky_band(double* band, int64_t N, int KB);
void ref_band_to_block(const double* band, double* ld, double* ls, int64_t T, int D);
void ref_solve_triang_mat(const double* lband, double* x, int64_t N, int KB, int transpose);

/* ---- small dense helpers (row-major n x n), what tf.linalg.* does per batch element ---------------- */
static int chol_small(const double* a, double* l, int n) { /* lower Cholesky, reads the lower triangle */
  memset(l, 0, sizeof(double) * n * n);
  for (int j = 0; j < n; ++j) {
    double s = a[j * n + j];
    for (int p = 0; p < j; ++p) s -= l[j * n + p] * l[j * n + p];
    if (!(s > 0.0)) return j + 1;
    const double ljj = sqrt(s);
    l[j * n + j] = ljj;
    for (int i = j + 1; i < n; ++i) {
      double v = a[i * n + j];
      for (int p = 0; p < j; ++p) v -= l[i * n + p] * l[j * n + p];
      l[i * n + j] = v / ljj;
    }
  }
  return 0;
}

/* x <- (L L^T)^-1 x for an n x k right-hand side (tf.linalg.cholesky_solve) */
static void chol_solve_small(const double* l, double* x, int n, int k) {
  for (int c = 0; c < k; ++c) {
    for (int i = 0; i < n; ++i) {
      double v = x[i * k + c];
      for (int p = 0; p < i; ++p) v -= l[i * n + p] * x[p * k + c];
      x[i * k + c] = v / l[i * n + i];
    }
    for (int i = n - 1; i >= 0; --i) {
      double v = x[i * k + c];
      for (int p = i + 1; p < n; ++p) v -= l[p * n + i] * x[p * k + c];
      x[i * k + c] = v / l[i * n + i];
    }
  }
}

/* x <- a^-1 x by LU with partial pivoting (tf.linalg.solve); a is destroyed */
static void lu_solve_small(double* a, double* x, int n, int k) {
  for (int j = 0; j < n; ++j) {
    int piv = j;
    for (int i = j + 1; i < n; ++i)
      if (fabs(a[i * n + j]) > fabs(a[piv * n + j])) piv = i;
    if (piv != j) {
      for (int c = 0; c < n; ++c) { double t = a[j * n + c]; a[j * n + c] = a[piv * n + c]; a[piv * n + c] = t; }
      for (int c = 0; c < k; ++c) { double t = x[j * k + c]; x[j * k + c] = x[piv * k + c]; x[piv * k + c] = t; }
    }
    for (int i = j + 1; i < n; ++i) {
      const double f = a[i * n + j] / a[j * n + j];
      for (int c = j; c < n; ++c) a[i * n + c] -= f * a[j * n + c];
      for (int c = 0; c < k; ++c) x[i * k + c] -= f * x[j * k + c];
    }
  }
  for (int i = n - 1; i >= 0; --i)
    for (int c = 0; c < k; ++c) {
      double v = x[i * k + c];
      for (int p = i + 1; p < n; ++p) v -= a[i * n + p] * x[p * k + c];
      x[i * k + c] = v / a[i * n + i];
    }
}

static void eye_small(double* a, int n) {
  memset(a, 0, sizeof(double) * n * n);
  for (int i = 0; i < n; ++i) a[i * n + i] = 1.0;
}

/* ---- StateSpaceModel._build_precision (+ optional H^T R^-1 H), state_space_model.py:431-483 --------- */
static void build_precision(const double* chol_p0, const double* A, const double* cholQ, const double* H,
                            const double* r_inv, double* diag, double* sub, int64_t T, int D, int m) {
  double inv_q_a[MAXD * MAXD], inv_q[MAXD * MAXD], tmp[MAXD * MAXD];
  const int DD = D * D;
  for (int64_t k = 0; k < T; ++k) {
    const double* lk = k == 0 ? chol_p0 : cholQ + (k - 1) * DD;
    eye_small(inv_q, D);
    chol_solve_small(lk, inv_q, D, D); /* concatted_inv_q_s */
    memcpy(diag + k * DD, inv_q, sizeof(double) * DD);
  }
  for (int64_t k = 0; k + 1 < T; ++k) {
    memcpy(inv_q_a, A + k * DD, sizeof(double) * DD);
    chol_solve_small(cholQ + k * DD, inv_q_a, D, D); /* Q_k^-1 A_k */
    for (int i = 0; i < D; ++i)
      for (int j = 0; j < D; ++j) {
        double s = 0.0;
        for (int p = 0; p < D; ++p) s += A[k * DD + p * D + i] * inv_q_a[p * D + j]; /* A^T Q^-1 A */
        diag[k * DD + i * D + j] += s;
        sub[k * DD + i * D + j] = -inv_q_a[i * D + j];
      }
  }
  if (H) /* + H_k^T R^-1 H_k, kalman_filter.py:85-101 */
    for (int64_t k = 0; k < T; ++k) {
      const double* hk = H + k * m * D;
      for (int o = 0; o < m; ++o)
        for (int j = 0; j < D; ++j) {
          double s = 0.0;
          for (int p = 0; p < m; ++p) s += r_inv[o * m + p] * hk[p * D + j];
          tmp[o * D + j] = s;
        }
      for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j) {
          double s = 0.0;
          for (int o = 0; o < m; ++o) s += hk[o * D + i] * tmp[o * D + j];
          diag[k * DD + i * D + j] += s;
        }
    }
}

/* marginal_means = a_inv_block.solve(concatenated_state_offsets): band of A^-1 (identity diagonal, -A_k
 * sub-diagonal) + banded substitution, state_space_model.py:231-251,277-296 */
static void marginal_means(const double* mu0, const double* A, const double* b, double* mean, double* band,
                           double* blk_d, double* blk_s, int64_t T, int D) {
  const int DD = D * D;
  const int64_t N = T * D;
  for (int64_t k = 0; k < T; ++k) eye_small(blk_d + k * DD, D);
  for (int64_t i = 0; i < (T - 1) * DD; ++i) blk_s[i] = -A[i];
  ref_block_to_band(blk_d, blk_s, band, T, D);
  memcpy(mean, mu0, sizeof(double) * D);
  memcpy(mean + D, b, sizeof(double) * (T - 1) * D);
  ref_solve_triang_mat(band, mean, N, 2 * D - 1, 0);
}

/* banded_matrices.inverse_from_cholesky_band: lower band of (L L^T)^-1 (Takahashi recurrence) */
static void inverse_from_cholesky_band(const double* lb, double* sb, int64_t N, int KB) {
  for (int64_t j = N - 1; j >= 0; --j) {
    const int64_t hi = (N - 1 < j + KB) ? N - 1 : j + KB;
    const double ljj = lb[j];
    for (int64_t i = hi; i > j; --i) {
      double acc = 0.0;
      for (int64_t p = j + 1; p <= hi; ++p) {
        const double sip = (i >= p) ? sb[(i - p) * N + p] : sb[(p - i) * N + i];
        acc += sip * lb[(p - j) * N + j];
      }
      sb[(i - j) * N + j] = -acc / ljj;
    }
    double acc = 0.0;
    for (int64_t p = j + 1; p <= hi; ++p) acc += sb[(p - j) * N + j] * lb[(p - j) * N + j];
    sb[j] = 1.0 / (ljj * ljj) - acc / ljj;
    for (int64_t r = hi - j + 1; r <= KB; ++r) sb[r * N + j] = 0.0;
  }
}

/* ---- KalmanFilter.log_likelihood, one chain ---------------------------------------------------------- */
typedef struct {
  double *diag, *sub, *band, *ld, *ls, *mean, *vec, *blk_d, *blk_s;
} Scratch;

static Scratch scratch_alloc(int64_t T, int D) {
  Scratch s;
  const size_t DD = (size_t)D * D, N = (size_t)T * D;
  s.diag = (double*)malloc(sizeof(double) * T * DD);
  s.sub = (double*)malloc(sizeof(double) * T * DD);
  s.band = (double*)malloc(sizeof(double) * 2 * D * N);
  s.ld = (double*)malloc(sizeof(double) * T * DD);
  s.ls = (double*)malloc(sizeof(double) * T * DD);
  s.mean = (double*)malloc(sizeof(double) * N);
  s.vec = (double*)malloc(sizeof(double) * N);
  s.blk_d = (double*)malloc(sizeof(double) * T * DD);
  s.blk_s = (double*)malloc(sizeof(double) * T * DD);
  return s;
}
static void scratch_free(Scratch* s) {
  free(s->diag); free(s->sub); free(s->band); free(s->ld); free(s->ls); free(s->mean); free(s->vec);
  free(s->blk_d); free(s->blk_s);
}

static double kalman_loglik_chain(const double* mu0, const double* chol_p0, const double* A, const double* b,
                                  const double* cholQ, const double* H, const double* y, const double* cholR,
                                  int64_t T, int D, int m, Scratch* s, int* fail) {
  const int DD = D * D;
  const int64_t N = T * D;
  double r_inv[MAXD * MAXD], disp[MAXD], rd[MAXD];
  eye_small(r_inv, m);
  chol_solve_small(cholR, r_inv, m, m); /* KalmanFilter._r_inv :341-348 */
  build_precision(chol_p0, A, cholQ, H, r_inv, s->diag, s->sub, T, D, m);
  /* l_post = _k_inv_post.cholesky: block->band, banded Cholesky, band->block */
  ref_block_to_band(s->diag, s->sub, s->band, T, D);
  if (ref_cholesky_band(s->band, N, 2 * D - 1)) *fail = 1;
  ref_band_to_block(s->band, s->ld, s->ls, T, D);
  marginal_means(mu0, A, b, s->mean, s->band, s->blk_d, s->blk_s, T, D);
  double term1 = 0.0;
  for (int64_t k = 0; k < T; ++k) {
    const double* hk = H + k * m * D;
    for (int o = 0; o < m; ++o) { /* disp = obs - H mu */
      double f = 0.0;
      for (int j = 0; j < D; ++j) f += hk[o * D + j] * s->mean[k * D + j];
      disp[o] = y[k * m + o] - f;
    }
    for (int o = 0; o < m; ++o) {
      double v = 0.0;
      for (int p = 0; p < m; ++p) v += r_inv[o * m + p] * disp[p];
      rd[o] = v;
      term1 += v * disp[o];
    }
    for (int j = 0; j < D; ++j) { /* obs_proj = H^T R^-1 disp */
      double v = 0.0;
      for (int o = 0; o < m; ++o) v += hk[o * D + j] * rd[o];
      s->vec[k * D + j] = v;
    }
  }
  /* l_post.solve(obs_proj): as_band again + banded substitution */
  ref_block_to_band(s->ld, s->ls, s->band, T, D);
  ref_solve_triang_mat(s->band, s->vec, N, 2 * D - 1, 0);
  double term2 = 0.0;
  for (int64_t i = 0; i < N; ++i) term2 += s->vec[i] * s->vec[i];
  /* l_post.abs_log_det(): row 0 of the band (yet another repack in the reference) */
  double logdet_l = 0.0;
  for (int64_t i = 0; i < N; ++i) logdet_l += log(fabs(s->band[i]));
  double logdet_prior = 0.0; /* log_det_precision :343-373 */
  for (int j = 0; j < D; ++j) logdet_prior += log(chol_p0[j * D + j] * chol_p0[j * D + j]);
  for (int64_t k = 0; k + 1 < T; ++k)
    for (int j = 0; j < D; ++j) logdet_prior += log(cholQ[k * DD + j * D + j] * cholQ[k * DD + j * D + j]);
  logdet_prior = -logdet_prior;
  double lr[MAXD * MAXD], logdet_rinv = 0.0;
  chol_small(r_inv, lr, m);
  for (int o = 0; o < m; ++o) logdet_rinv += 2.0 * log(lr[o * m + o]);
  const double cst = -0.5 * log(2.0 * M_PI) * (double)(m * T);
  return cst - 0.5 * term1 + 0.5 * term2 + 0.5 * logdet_prior - logdet_l + 0.5 * (double)T * logdet_rinv;
}

int64_t ref_kalman_loglik_batch(const double* mu0, const double* chol_p0, const double* A, const double* b,
                                const double* cholQ, const double* H, const double* y, const double* cholR,
                                double* out, int64_t B, int64_t T, int D, int m, int64_t Hb, int nthreads) {
  int64_t nfail = 0;
  if (D > MAXD || m > MAXD) return -1;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel reduction(+ : nfail)
  {
    Scratch s = scratch_alloc(T, D);
#pragma omp for schedule(static)
    for (int64_t c = 0; c < B; ++c) {
      int fail = 0;
      out[c] = kalman_loglik_chain(mu0 + c * D, chol_p0 + c * D * D, A + c * (T - 1) * D * D, b + c * (T - 1) * D,
                                   cholQ + c * (T - 1) * D * D, H + (Hb == 1 ? 0 : c) * T * m * D, y + c * T * m,
                                   cholR, T, D, m, &s, &fail);
      nfail += fail;
    }
    scratch_free(&s);
  }
  return nfail;
}

/* ---- naturals_to_ssm_params, ssm_gaussian_transformations.py:332-511 -------------------------------- */
static int nat_to_ssm_chain(const double* th_lin, const double* th_diag, const double* th_sub, double* As,
                            double* offsets, double* chol_p0, double* cholQ, double* mu0, int64_t T, int D,
                            Scratch* s, double* pband, double* sband) {
  const int DD = D * D, KB = 2 * D - 1;
  const int64_t N = T * D;
  int fail = 0;
  /* precision = SymmetricBlockTriDiagonal(-2 theta_diag, -theta_subdiag) */
  for (int64_t i = 0; i < T * DD; ++i) s->diag[i] = -2.0 * th_diag[i];
  for (int64_t i = 0; i < (T - 1) * DD; ++i) s->sub[i] = -th_sub[i];
  /* precision.cholesky.as_band: block->band, Cholesky, band->block, block->band */
  ref_block_to_band(s->diag, s->sub, s->band, T, D);
  if (ref_cholesky_band(s->band, N, KB)) fail = 1;
  ref_band_to_block(s->band, s->ld, s->ls, T, D);
  ref_block_to_band(s->ld, s->ls, s->band, T, D);
  /* inverse_from_cholesky_band + band_to_block: marginal_covs (symmetric) and the lower sub-diagonal */
  inverse_from_cholesky_band(s->band, sband, N, KB);
  double cov[MAXD * MAXD], sd[MAXD * MAXD], cond[MAXD * MAXD], l1[MAXD * MAXD], l2[MAXD * MAXD];
  /* As = (marginal_covs^-1 sub_diag)^T, tf.linalg.solve (LU); sub_diag block k = Sigma_{k,k+1} */
  for (int64_t k = 0; k + 1 < T; ++k) {
    for (int i = 0; i < D; ++i)
      for (int j = 0; j < D; ++j) {
        const int lo = i >= j ? i : j, hi_ = i >= j ? j : i;
        cov[i * D + j] = sband[(int64_t)(lo - hi_) * N + k * D + hi_];
        /* Sigma_{k,k+1}[i][j] = Sigma[(k+1)D + j][kD + i] (lower band entry) */
        sd[i * D + j] = sband[(int64_t)(D + j - i) * N + k * D + i];
      }
    lu_solve_small(cov, sd, D, D);
    for (int i = 0; i < D; ++i)
      for (int j = 0; j < D; ++j) As[k * DD + i * D + j] = sd[j * D + i];
  }
  /* a_inv_block = LTBTD(eye, -As); tmp = solve_triang_band(a_inv_block.as_band, precision.as_band, transpose_left)
   * restricted to the lower band; its diagonal blocks are the conditional precisions */
  for (int64_t k = 0; k < T; ++k) eye_small(s->blk_d + k * DD, D);
  for (int64_t i = 0; i < (T - 1) * DD; ++i) s->blk_s[i] = -As[i];
  ref_block_to_band(s->blk_d, s->blk_s, s->band, T, D); /* band of A^-1 */
  ref_block_to_band(s->diag, s->sub, pband, T, D);       /* precision.as_band */
  for (int64_t j = 0; j < N; ++j) {
    const int64_t hi = (N - 1 < j + KB) ? N - 1 : j + KB;
    for (int64_t i = hi; i >= j; --i) {
      double v = pband[(i - j) * N + j];
      const int64_t pmax = (hi < i + KB) ? hi : i + KB;
      for (int64_t p = i + 1; p <= pmax; ++p) v -= s->band[(p - i) * N + i] * sband[(p - j) * N + j];
      sband[(i - j) * N + j] = v / s->band[i]; /* sband reused as the result band */
    }
  }
  /* conditional precisions -> cholesky -> cholesky_solve(eye) -> cholesky; offsets */
  memcpy(s->vec, th_lin, sizeof(double) * N);
  ref_solve_triang_mat(s->band, s->vec, N, KB, 1); /* a_inv_block.solve(theta_linear, transpose_left=True) */
  for (int64_t k = 0; k < T; ++k) {
    for (int i = 0; i < D; ++i)
      for (int j = 0; j <= i; ++j) {
        const double v = sband[(int64_t)(i - j) * N + k * D + j];
        cond[i * D + j] = v;
        cond[j * D + i] = v;
      }
    if (chol_small(cond, l1, D)) fail = 1;
    eye_small(cov, D);
    chol_solve_small(l1, cov, D, D); /* covariances */
    if (chol_small(cov, l2, D)) fail = 1;
    memcpy(k == 0 ? chol_p0 : cholQ + (k - 1) * DD, l2, sizeof(double) * DD);
    double* dst = k == 0 ? mu0 : offsets + (k - 1) * D;
    for (int i = 0; i < D; ++i) {
      double v = 0.0;
      for (int j = 0; j < D; ++j) v += cov[i * D + j] * s->vec[k * D + j];
      dst[i] = v;
    }
  }
  return fail;
}

int64_t ref_nat_to_ssm_batch(const double* th_lin, const double* th_diag, const double* th_sub, double* As,
                             double* offsets, double* chol_p0, double* cholQ, double* mu0, int64_t B, int64_t T,
                             int D, int nthreads) {
  int64_t nfail = 0;
  if (D > MAXD) return -1;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel reduction(+ : nfail)
  {
    Scratch s = scratch_alloc(T, D);
    double* pband = (double*)malloc(sizeof(double) * 2 * D * T * D);
    double* sband = (double*)malloc(sizeof(double) * 2 * D * T * D);
#pragma omp for schedule(static)
    for (int64_t c = 0; c < B; ++c)
      nfail += nat_to_ssm_chain(th_lin + c * T * D, th_diag + c * T * D * D, th_sub + c * (T - 1) * D * D,
                                As + c * (T - 1) * D * D, offsets + c * (T - 1) * D, chol_p0 + c * D * D,
                                cholQ + c * (T - 1) * D * D, mu0 + c * D, T, D, &s, pband, sband);
    free(pband);
    free(sband);
    scratch_free(&s);
  }
  return nfail;
}

/* ---- ssm_to_expectations, ssm_gaussian_transformations.py:31-89 ------------------------------------- */
int64_t ref_ssm_to_expectations_batch(const double* mu0, const double* chol_p0, const double* A, const double* b,
                                      const double* cholQ, double* eta_lin, double* eta_diag, double* eta_sub,
                                      int64_t B, int64_t T, int D, int nthreads) {
  int64_t nfail = 0;
  if (D > MAXD) return -1;
  const int DD = D * D, KB = 2 * D - 1;
  const int64_t N = T * D;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel reduction(+ : nfail)
  {
    Scratch s = scratch_alloc(T, D);
    double* sband = (double*)malloc(sizeof(double) * 2 * D * N);
#pragma omp for schedule(static)
    for (int64_t c = 0; c < B; ++c) {
      const double* Ac = A + c * (T - 1) * DD;
      double* el = eta_lin + c * N;
      double* ed = eta_diag + c * T * DD;
      double* es = eta_sub + c * (T - 1) * DD;
      marginal_means(mu0 + c * D, Ac, b + c * (T - 1) * D, el, s.band, s.blk_d, s.blk_s, T, D);
      /* marginal_covariances = precision.cholesky.block_diagonal_of_inverse() */
      build_precision(chol_p0 + c * DD, Ac, cholQ + c * (T - 1) * DD, NULL, NULL, s.diag, s.sub, T, D, 0);
      ref_block_to_band(s.diag, s.sub, s.band, T, D);
      if (ref_cholesky_band(s.band, N, KB)) ++nfail;
      ref_band_to_block(s.band, s.ld, s.ls, T, D);
      ref_block_to_band(s.ld, s.ls, s.band, T, D);
      inverse_from_cholesky_band(s.band, sband, N, KB);
      for (int64_t k = 0; k < T; ++k)
        for (int i = 0; i < D; ++i)
          for (int j = 0; j < D; ++j) {
            const int lo = i >= j ? i : j, hi_ = i >= j ? j : i;
            s.ld[k * DD + i * D + j] = sband[(int64_t)(lo - hi_) * N + k * D + hi_]; /* covariance block */
            ed[k * DD + i * D + j] = s.ld[k * DD + i * D + j] + el[k * D + i] * el[k * D + j];
          }
      for (int64_t k = 0; k + 1 < T; ++k)
        for (int i = 0; i < D; ++i)
          for (int j = 0; j < D; ++j) {
            double v = el[(k + 1) * D + i] * el[k * D + j];
            for (int p = 0; p < D; ++p) v += Ac[k * DD + i * D + p] * s.ld[k * DD + p * D + j];
            es[k * DD + i * D + j] = v;
          }
    }
    free(sband);
    scratch_free(&s);
  }
  return nfail;
}
