// Reverse-mode (vector-Jacobian product) kernels of the sequential recurrences -- SURVEY.md §8f-3.
//
// The reference differentiates through every banded op with TensorFlow's tape (gradients registered by
// banded-matrices; callers: ssm_natgrad.py:142-172, tests/integration/models/
// test_gaussian_process_regression.py:117-130).  Here each recurrence has an adjoint sweep that runs the
// chain in the opposite direction with the adjoint state in registers:
//
//   btd_cholesky_bwd_kernel   adjoint of  S_k = D_k - Ls_{k-1} Ls_{k-1}^T, Ld_k = chol(S_k), Ls_k = A_k Ld_k^-T
//                             (SymmetricBlockTriDiagonal.cholesky, block_tri_diag.py:423-436)
//   ssm_marginals_bwd_kernel  adjoint of  mu_{k+1} = A_k mu_k + b_k,  P_{k+1} = A_k P_k A_k^T + Lq_k Lq_k^T,
//                             sub_k = A_k P_k   (marginal_means / marginal_covariances / covariance_blocks,
//                             state_space_model.py:231-275,326-341)
//
// One thread per chain, D x D blocks in registers (smallmat.cuh), D <= MF_SMALL_D_MAX.
#pragma once
#include <cstdint>

#include "smallmat.cuh"

namespace mf {

// ---------------------------------------------------------------------------------------------
// Cholesky adjoint.  Inputs: the factor (ld [B,T,D,D], ls [B,T-1,D,D] or NULL) and the adjoints of its
// blocks (g_ld lower triangles read, g_ls; either may be NULL = zero).  Outputs: g_diag [B,T,D,D] --
// the gradient with respect to the entries the forward pass READS (lower triangles: off-diagonal
// entries carry both symmetric positions, upper triangles are zero) -- and g_sub [B,T-1,D,D].
//   per step, with S_bar the symmetric adjoint of S_{k+1} carried from the later step:
//     Gs  = g_ls_k - 2 S_bar Ls_k;   A_bar_k = Gs Ld_k^-1;   Gd = tril(g_ld_k) - tril(A_bar_k^T Ls_k)
//     P   = Phi(Ld_k^T Gd)  (lower triangle, diagonal halved);   S_bar = 1/2 Ld_k^-T (P + P^T) Ld_k^-1
// ---------------------------------------------------------------------------------------------
template <typename T, int D>
__global__ void __launch_bounds__(32)
btd_cholesky_bwd_kernel(const T* __restrict__ ld, const T* __restrict__ ls, const T* __restrict__ g_ld,
                        const T* __restrict__ g_ls, T* __restrict__ g_diag, T* __restrict__ g_sub,
                        int64_t B, int64_t Tn) {
  constexpr int DD = D * D;
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= B) return;
  const T* ldc = ld + c * Tn * DD;
  const T* lsc = ls ? ls + c * (Tn - 1) * DD : nullptr;
  const T* gdc = g_ld ? g_ld + c * Tn * DD : nullptr;
  const T* gsc = (g_ls && ls) ? g_ls + c * (Tn - 1) * DD : nullptr;
  T* odc = g_diag + c * Tn * DD;
  T* osc = (g_sub && ls) ? g_sub + c * (Tn - 1) * DD : nullptr;
  T Sb[DD];  // symmetric adjoint of S_{k+1}
#pragma unroll
  for (int i = 0; i < DD; ++i) Sb[i] = T(0);
  for (int64_t k = Tn - 1; k >= 0; --k) {
    T L[DD], rinv[D], Gd[DD];
    load_vec<T, DD>(L, ldc + k * DD);
#pragma unroll
    for (int j = 0; j < D; ++j) rinv[j] = Num<T>::rcp(L[j * D + j]);
    if (gdc) {
      load_vec<T, DD>(Gd, gdc + k * DD);
    } else {
#pragma unroll
      for (int i = 0; i < DD; ++i) Gd[i] = T(0);
    }
    if (lsc && k + 1 < Tn) {
      T Ls[DD], Gs[DD];
      load_vec<T, DD>(Ls, lsc + k * DD);
      if (gsc) {
        load_vec<T, DD>(Gs, gsc + k * DD);
      } else {
#pragma unroll
        for (int i = 0; i < DD; ++i) Gs[i] = T(0);
      }
      // Gs -= 2 S_bar Ls
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) {
          T v = Gs[i * D + j];
#pragma unroll
          for (int q = 0; q < D; ++q) v = Num<T>::fma(T(-2) * Sb[i * D + q], Ls[q * D + j], v);
          Gs[i * D + j] = v;
        }
      trsm_right_lower<T, D>(Gs, L, rinv);  // A_bar = Gs L^-1
      if (osc) store_vec<T, DD>(osc + k * DD, Gs);
      // Gd -= tril(A_bar^T Ls)
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
          T v = Gd[i * D + j];
#pragma unroll
          for (int q = 0; q < D; ++q) v = Num<T>::fma(-Gs[q * D + i], Ls[q * D + j], v);
          Gd[i * D + j] = v;
        }
    }
    // P = Phi(L^T tril(Gd)); symmetrise: M = P + P^T (diagonal: the un-halved value)
    T M[DD];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int j = 0; j <= i; ++j) {
        T v = T(0);
#pragma unroll
        for (int q = i; q < D; ++q) v = Num<T>::fma(L[q * D + i], Gd[q * D + j], v);  // (L^T Gd)_ij, Gd lower
        M[i * D + j] = v;
        M[j * D + i] = v;
      }
    // S_bar = 1/2 L^-T M L^-1
    trsm_left_lower_t<T, D>(L, rinv, M);
    trsm_right_lower<T, D>(M, L, rinv);
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int j = 0; j <= i; ++j) {
        const T v = T(0.5) * T(0.5) * (M[i * D + j] + M[j * D + i]);  // 1/2, and re-symmetrised against rounding
        Sb[i * D + j] = v;
        Sb[j * D + i] = v;
      }
    T out[DD];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int j = 0; j < D; ++j) out[i * D + j] = j < i ? T(2) * Sb[i * D + j] : (j == i ? Sb[i * D + j] : T(0));
    store_vec<T, DD>(odc + k * DD, out);
  }
}

// ---------------------------------------------------------------------------------------------
// Adjoint of the marginal-moment recursion.  Inputs: SSM parameters, the forward results mean [B,T,D] and
// cov [B,T,D,D], and the adjoints g_mean [B,T,D], g_cov [B,T,D,D], g_sub [B,T-1,D,D] (each may be NULL).
// Outputs: g_mu0 [B,D], g_l0 [B,D,D], g_a [B,T-1,D,D], g_b [B,T-1,D], g_lq [B,T-1,D,D] (lower triangles).
// ---------------------------------------------------------------------------------------------
template <typename T, int D>
__global__ void __launch_bounds__(32)
ssm_marginals_bwd_kernel(const T* __restrict__ chol_p0, const T* __restrict__ a, const T* __restrict__ chol_q,
                         const T* __restrict__ mean, const T* __restrict__ cov, const T* __restrict__ g_mean,
                         const T* __restrict__ g_cov, const T* __restrict__ g_sub, T* __restrict__ g_mu0,
                         T* __restrict__ g_l0, T* __restrict__ g_a, T* __restrict__ g_b, T* __restrict__ g_lq,
                         int64_t B, int64_t Tn) {
  constexpr int DD = D * D;
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= B) return;
  const int64_t n = Tn - 1;
  T mb[D], Pb[DD];  // adjoints of mu_{k+1}, P_{k+1} (all contributions from later steps included)
#pragma unroll
  for (int i = 0; i < D; ++i) mb[i] = g_mean ? g_mean[(c * Tn + n) * D + i] : T(0);
#pragma unroll
  for (int i = 0; i < DD; ++i) Pb[i] = g_cov ? g_cov[(c * Tn + n) * DD + i] : T(0);
  for (int64_t k = n - 1; k >= 0; --k) {
    T A[DD], Lq[DD], P[DD], mu[D], Gs[DD], Ps[DD], W[DD], out[DD];
    load_vec<T, DD>(A, a + (c * n + k) * DD);
    load_vec<T, DD>(Lq, chol_q + (c * n + k) * DD);
    load_vec<T, DD>(P, cov + (c * Tn + k) * DD);
    load_vec<T, D>(mu, mean + (c * Tn + k) * D);
    if (g_sub) {
      load_vec<T, DD>(Gs, g_sub + (c * n + k) * DD);
    } else {
#pragma unroll
      for (int i = 0; i < DD; ++i) Gs[i] = T(0);
    }
    store_vec<T, D>(g_b + (c * n + k) * D, mb);
    // Ps = P_bar + P_bar^T
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int j = 0; j < D; ++j) Ps[i * D + j] = Pb[i * D + j] + Pb[j * D + i];
    // g_lq = tril(Ps Lq)   (Lq lower)
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int j = 0; j < D; ++j) {
        T v = T(0);
        if (j <= i) {
#pragma unroll
          for (int q = j; q < D; ++q) v = Num<T>::fma(Ps[i * D + q], Lq[q * D + j], v);
        }
        out[i * D + j] = v;
      }
    store_vec<T, DD>(g_lq + (c * n + k) * DD, out);
    // g_a = m_bar mu^T + (Ps A + Gs) P        (P symmetric)
    gemm<T, D>(W, Ps, A);
#pragma unroll
    for (int i = 0; i < DD; ++i) W[i] += Gs[i];
    gemm<T, D>(out, W, P);
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int j = 0; j < D; ++j) out[i * D + j] = Num<T>::fma(mb[i], mu[j], out[i * D + j]);
    store_vec<T, DD>(g_a + (c * n + k) * DD, out);
    // P_bar_k = g_cov_k + A^T P_bar A + A^T Gs;   m_bar_k = g_mean_k + A^T m_bar
    gemm<T, D>(W, Pb, A);
#pragma unroll
    for (int i = 0; i < DD; ++i) W[i] += Gs[i];
    gemm_tn<T, D>(out, A, W);
    T mn[D];
#pragma unroll
    for (int i = 0; i < D; ++i) mn[i] = g_mean ? g_mean[(c * Tn + k) * D + i] : T(0);
    gemv_t_add<T, D>(mn, A, mb);
#pragma unroll
    for (int i = 0; i < D; ++i) mb[i] = mn[i];
#pragma unroll
    for (int i = 0; i < DD; ++i) Pb[i] = out[i] + (g_cov ? g_cov[(c * Tn + k) * DD + i] : T(0));
  }
  store_vec<T, D>(g_mu0 + c * D, mb);
  T L0[DD], out[DD];
  load_vec<T, DD>(L0, chol_p0 + c * DD);
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) {
      T v = T(0);
      if (j <= i) {
#pragma unroll
        for (int q = j; q < D; ++q) v = Num<T>::fma(Pb[i * D + q] + Pb[q * D + i], L0[q * D + j], v);
      }
      out[i * D + j] = v;
    }
  store_vec<T, DD>(g_l0 + c * DD, out);
}

}  // namespace mf
