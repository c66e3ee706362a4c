// Natural / expectation parameter transforms of a state-space model
// (reference markovflow/ssm_gaussian_transformations.py).
//
//   theta = (theta_lin [B,T,D], theta_diag [B,T,D,D], theta_sub [B,T-1,D,D])   natural params:
//           precision blocks P_kk = -2 theta_diag_k, P_{k+1,k} = -theta_sub_k, theta_lin = P mu
//   eta   = (eta_lin, eta_diag, eta_sub)  expectation params: E[x_k], E[x_k x_k^T], E[x_{k+1} x_k^T]
//   SSM outputs use the concatenated layout the reference returns slices of:
//           a [B,T-1,D,D], offsets [B,T,D] = [mu0, b_1..], chols [B,T,D,D] = [chol P0, chol Q_1..]
#pragma once
#include "ssm_kernels.cuh"

namespace mf {

// ---------------------------------------------------------------------------------------------
// naturals_to_ssm_params (ssm_gaussian_transformations.py:332-511) as ONE backward sweep.
//
// The precision of an SSM factors as P = U D U^T with U^T = A^{-1} (identity diagonal, -A_k below
// it) and D = blockdiag(P0^{-1}, Q_1^{-1}, ...).  Running that factorisation backwards gives every
// parameter directly:
//     D_{T-1} = P_{T-1,T-1}
//     A_k     = -D_{k+1}^{-1} P_{k+1,k}                 (= D_{k+1}^{-1} theta_sub_k)
//     D_k     = P_kk - P_{k+1,k}^T D_{k+1}^{-1} P_{k+1,k} = -2 theta_diag_k - theta_sub_k^T A_k
//     z_k     = theta_lin_k + A_k^T z_{k+1}             (A^T-solve of the linear term)
//     Q_k     = D_k^{-1},  offset_k = Q_k z_k,  chol_k = chol(Q_k)
// The reference reaches the same quantities through a banded Cholesky, the sparse inverse subset,
// a general solve, a banded triangular solve and two batched Choleskys (six sweeps).
// One thread per chain.
// ---------------------------------------------------------------------------------------------
template <typename T, int D>
__global__ void __launch_bounds__(32)
nat_to_ssm_kernel(const T* __restrict__ th_lin, const T* __restrict__ th_diag,
                  const T* __restrict__ th_sub, T* __restrict__ out_a, T* __restrict__ out_off,
                  T* __restrict__ out_chol, int32_t* __restrict__ info, int64_t B, int64_t Tn) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= B) return;
  constexpr int DD = D * D;
  const T* lp = th_lin + c * Tn * D;
  const T* dp = th_diag + c * Tn * DD;
  const T* sp = th_sub + c * (Tn - 1) * DD;
  T* ap = out_a + c * (Tn - 1) * DD;
  T* op = out_off + c * Tn * D;
  T* cp = out_chol + c * Tn * DD;
  T Dk[DD], rinv[D], z[D], S[DD], A[DD], Qc[DD], off[D];
  int32_t fail = 0;
#pragma unroll
  for (int i = 0; i < D; ++i) z[i] = T(0);
  for (int64_t k = Tn - 1; k >= 0; --k) {
    T th[D];
    load_vec<T, DD>(Dk, dp + k * DD);
    load_vec<T, D>(th, lp + k * D);
#pragma unroll
    for (int i = 0; i < DD; ++i) Dk[i] = T(-2) * Dk[i];
    if (k + 1 < Tn) {
      // (Dk still holds P_kk; S/rinv hold chol(D_{k+1}) from the previous iteration)
      load_vec<T, DD>(A, sp + k * DD);
      T Th[DD];
#pragma unroll
      for (int i = 0; i < DD; ++i) Th[i] = A[i];
      trsm_left_lower<T, D>(S, rinv, A);
      trsm_left_lower_t<T, D>(S, rinv, A);  // A_k = D_{k+1}^{-1} theta_sub_k
      store_vec<T, DD>(ap + k * DD, A);
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
          T v = Dk[i * D + j];
#pragma unroll
          for (int q = 0; q < D; ++q) v = Num<T>::fma(-Th[q * D + i], A[q * D + j], v);
          Dk[i * D + j] = v;
        }
      gemv_t_add<T, D>(th, A, z);  // z_k = theta_lin_k + A_k^T z_{k+1}
    }
#pragma unroll
    for (int i = 0; i < D; ++i) z[i] = th[i];
#pragma unroll
    for (int i = 0; i < DD; ++i) S[i] = Dk[i];
    const bool ok = chol_lower<T, D>(S, rinv);  // D_k = S S^T
    if (!ok && fail == 0) fail = (int32_t)(k + 1);
    // offset_k = D_k^{-1} z_k
#pragma unroll
    for (int i = 0; i < D; ++i) off[i] = z[i];
    trsv_lower<T, D>(S, rinv, off);
    trsv_lower_t<T, D>(S, rinv, off);
    store_vec<T, D>(op + k * D, off);
    // chol(Q_k) with Q_k = D_k^{-1}
    chol_inverse<T, D>(Qc, S, rinv);
    T r2[D];
    chol_lower<T, D>(Qc, r2);
    zero_upper<T, D>(Qc);
    store_vec<T, DD>(cp + k * DD, Qc);
  }
  if (info) info[c] = fail;
}

// ---------------------------------------------------------------------------------------------
// ssm_to_expectations (ssm_gaussian_transformations.py:31-89): forward sweep, one thread per chain.
//   eta_lin_k = mu_k, eta_diag_k = Sigma_kk + mu_k mu_k^T,
//   eta_sub_k = A_k Sigma_kk + mu_{k+1} mu_k^T
// ---------------------------------------------------------------------------------------------
template <typename T, int D>
__global__ void __launch_bounds__(32)
ssm_to_expectations_kernel(const T* __restrict__ mu0, const T* __restrict__ chol_p0,
                           const T* __restrict__ a, const T* __restrict__ b,
                           const T* __restrict__ chol_q, T* __restrict__ eta_lin,
                           T* __restrict__ eta_diag, T* __restrict__ eta_sub, int64_t B,
                           int64_t Tn) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= B) return;
  constexpr int DD = D * D;
  const T* ap = a + c * (Tn - 1) * DD;
  const T* bp = b + c * (Tn - 1) * D;
  const T* qp = chol_q + c * (Tn - 1) * DD;
  T mu[D], P[DD], A[DD], L[DD], AP[DD], E[DD], nmu[D];
  load_vec<T, D>(mu, mu0 + c * D);
  load_vec<T, DD>(L, chol_p0 + c * DD);
  llt<T, D>(P, L);
  for (int64_t k = 0; k < Tn; ++k) {
    store_vec<T, D>(eta_lin + (c * Tn + k) * D, mu);
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int j = 0; j < D; ++j) E[i * D + j] = Num<T>::fma(mu[i], mu[j], P[i * D + j]);
    store_vec<T, DD>(eta_diag + (c * Tn + k) * DD, E);
    if (k + 1 == Tn) break;
    load_vec<T, DD>(A, ap + k * DD);
    load_vec<T, D>(nmu, bp + k * D);
    load_vec<T, DD>(L, qp + k * DD);
    gemm<T, D>(AP, A, P);
    gemv_add<T, D>(nmu, A, mu);
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int j = 0; j < D; ++j) E[i * D + j] = Num<T>::fma(nmu[i], mu[j], AP[i * D + j]);
    store_vec<T, DD>(eta_sub + (c * (Tn - 1) + k) * DD, E);
    llt<T, D>(P, L);
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int j = 0; j <= i; ++j) {
        T v = P[i * D + j];
#pragma unroll
        for (int q = 0; q < D; ++q) v = Num<T>::fma(AP[i * D + q], A[j * D + q], v);
        P[i * D + j] = v;
        P[j * D + i] = v;
      }
#pragma unroll
    for (int i = 0; i < D; ++i) mu[i] = nmu[i];
  }
}

// ---------------------------------------------------------------------------------------------
// expectations_to_ssm_params (ssm_gaussian_transformations.py:92-178): one thread per (chain, k).
//   Sigma_k = eta_diag_k - eta_k eta_k^T,  Sigma_{k,k+1} = eta_sub_k^T - eta_k eta_{k+1}^T
//   A_k = (Sigma_k^{-1} Sigma_{k,k+1})^T,  b_k = eta_{k+1} - A_k eta_k,
//   chol_k = chol(Sigma_k) for k = 0, chol(Sigma_{k} - A_{k-1} Sigma_{k-1} A_{k-1}^T) for k >= 1
// ---------------------------------------------------------------------------------------------
template <typename T, int D>
__device__ __forceinline__ void cov_from_eta(T* __restrict__ S, const T* __restrict__ eta_diag,
                                             const T* __restrict__ eta) {
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) S[i * D + j] = Num<T>::fma(-eta[i], eta[j], eta_diag[i * D + j]);
}

template <typename T, int D>
__global__ void __launch_bounds__(128)
expectations_to_ssm_kernel(const T* __restrict__ eta_lin, const T* __restrict__ eta_diag,
                           const T* __restrict__ eta_sub, T* __restrict__ out_a,
                           T* __restrict__ out_off, T* __restrict__ out_chol,
                           int32_t* __restrict__ info, int64_t B, int64_t Tn) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= B * Tn) return;
  constexpr int DD = D * D;
  const int64_t c = idx / Tn, k = idx % Tn;
  T ek[D], Sk[DD], Ed[DD], rinv[D];
  load_vec<T, D>(ek, eta_lin + idx * D);
  load_vec<T, DD>(Ed, eta_diag + idx * DD);
  cov_from_eta<T, D>(Sk, Ed, ek);
  bool ok = true;
  if (k == 0) {
    store_vec<T, D>(out_off + idx * D, ek);
    ok = chol_lower<T, D>(Sk, rinv);
    zero_upper<T, D>(Sk);
    store_vec<T, DD>(out_chol + idx * DD, Sk);
  } else {
    // transition k-1 -> k
    T ep[D], Sp[DD], Lp[DD], X[DD], Es[DD], off[D];
    load_vec<T, D>(ep, eta_lin + (idx - 1) * D);
    load_vec<T, DD>(Ed, eta_diag + (idx - 1) * DD);
    cov_from_eta<T, D>(Sp, Ed, ep);
    load_vec<T, DD>(Es, eta_sub + (c * (Tn - 1) + k - 1) * DD);
    // X = Sigma_{k-1,k} = eta_sub^T - eta_{k-1} eta_k^T
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int j = 0; j < D; ++j) X[i * D + j] = Num<T>::fma(-ep[i], ek[j], Es[j * D + i]);
#pragma unroll
    for (int i = 0; i < DD; ++i) Lp[i] = Sp[i];
    ok = chol_lower<T, D>(Lp, rinv);
    trsm_left_lower<T, D>(Lp, rinv, X);
    trsm_left_lower_t<T, D>(Lp, rinv, X);  // Sigma_{k-1}^{-1} Sigma_{k-1,k} = A^T
    T A[DD];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int j = 0; j < D; ++j) A[i * D + j] = X[j * D + i];
    store_vec<T, DD>(out_a + (c * (Tn - 1) + k - 1) * DD, A);
#pragma unroll
    for (int i = 0; i < D; ++i) off[i] = ek[i];
    gemv_sub<T, D>(off, A, ep);
    store_vec<T, D>(out_off + idx * D, off);
    T ASp[DD];
    gemm<T, D>(ASp, A, Sp);
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int j = 0; j <= i; ++j) {
        T v = Sk[i * D + j];
#pragma unroll
        for (int q = 0; q < D; ++q) v = Num<T>::fma(-ASp[i * D + q], A[j * D + q], v);
        Sk[i * D + j] = v;
      }
    ok = chol_lower<T, D>(Sk, rinv) && ok;
    zero_upper<T, D>(Sk);
    store_vec<T, DD>(out_chol + idx * DD, Sk);
  }
  if (info && !ok) atomicMax(info + c, (int32_t)(k + 1));
}

// ---------------------------------------------------------------------------------------------
// ssm_to_naturals (ssm_gaussian_transformations.py:181-253) and its no-smoothing variant
// (:256-329): one thread per (chain, k).
//   smoothing:    theta_sub_k = Q_{k+1}^{-1} A_k,  theta_diag_k = -1/2 (Q_k^{-1} + A_k^T Q_{k+1}^{-1} A_k),
//                 theta_lin_k = Q_k^{-1} m_k - A_k^T Q_{k+1}^{-1} m_{k+1}
//   no smoothing: theta_sub_k = Q_{k+1}^{-1} A_k,  theta_diag_k = -1/2 Q_k^{-1},  theta_lin_k = Q_k^{-1} m_k
// (Q_0 = P0, m = [mu0, b_1, ...])
// ---------------------------------------------------------------------------------------------
template <typename T, int D>
__global__ void __launch_bounds__(128)
ssm_to_naturals_kernel(const T* __restrict__ mu0, const T* __restrict__ chol_p0,
                       const T* __restrict__ a, const T* __restrict__ b,
                       const T* __restrict__ chol_q, T* __restrict__ th_lin,
                       T* __restrict__ th_diag, T* __restrict__ th_sub, int64_t B, int64_t Tn,
                       int smoothing) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= B * Tn) return;
  constexpr int DD = D * D;
  const int64_t c = idx / Tn, k = idx % Tn;
  const int64_t tr = c * (Tn - 1);
  T L[DD], Qi[DD], rinv[D], lin[D];
  load_vec<T, DD>(L, k == 0 ? chol_p0 + c * DD : chol_q + (tr + k - 1) * DD);
  load_vec<T, D>(lin, k == 0 ? mu0 + c * D : b + (tr + k - 1) * D);
  diag_rcp<T, D>(L, rinv);
  chol_inverse<T, D>(Qi, L, rinv);
  trsv_lower<T, D>(L, rinv, lin);
  trsv_lower_t<T, D>(L, rinv, lin);  // Q_k^{-1} m_k
  if (k + 1 < Tn) {
    T A[DD], X[DD], nm[D];
    load_vec<T, DD>(L, chol_q + (tr + k) * DD);
    load_vec<T, DD>(A, a + (tr + k) * DD);
    load_vec<T, D>(nm, b + (tr + k) * D);
    diag_rcp<T, D>(L, rinv);
#pragma unroll
    for (int i = 0; i < DD; ++i) X[i] = A[i];
    trsm_left_lower<T, D>(L, rinv, X);
    trsm_left_lower_t<T, D>(L, rinv, X);  // Q_{k+1}^{-1} A_k
    store_vec<T, DD>(th_sub + (tr + k) * DD, X);
    if (smoothing) {
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
          T v = T(0);
#pragma unroll
          for (int q = 0; q < D; ++q) v = Num<T>::fma(A[q * D + i], X[q * D + j], v);
          Qi[i * D + j] += v;
          if (i != j) Qi[j * D + i] += v;
        }
      gemv_t_sub<T, D>(lin, X, nm);  // - A^T Q^{-1} m_{k+1} = - X^T m_{k+1}
    }
  }
#pragma unroll
  for (int i = 0; i < DD; ++i) Qi[i] = T(-0.5) * Qi[i];
  store_vec<T, DD>(th_diag + idx * DD, Qi);
  store_vec<T, D>(th_lin + idx * D, lin);
}

// ---------------------------------------------------------------------------------------------
// naturals_to_ssm_params_no_smoothing (ssm_gaussian_transformations.py:514-593): per (chain, k).
//   C_k = chol(-2 theta_diag_k); A_{k-1} = (C_k C_k^T)^{-1} theta_sub_{k-1};
//   offset_k = (C_k C_k^T)^{-1} theta_lin_k; chol_k = chol((C_k C_k^T)^{-1})
// ---------------------------------------------------------------------------------------------
template <typename T, int D>
__global__ void __launch_bounds__(128)
nat_to_ssm_no_smoothing_kernel(const T* __restrict__ th_lin, const T* __restrict__ th_diag,
                               const T* __restrict__ th_sub, T* __restrict__ out_a,
                               T* __restrict__ out_off, T* __restrict__ out_chol,
                               int32_t* __restrict__ info, int64_t B, int64_t Tn) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= B * Tn) return;
  constexpr int DD = D * D;
  const int64_t c = idx / Tn, k = idx % Tn;
  T S[DD], rinv[D], off[D], Qc[DD], r2[D];
  load_vec<T, DD>(S, th_diag + idx * DD);
#pragma unroll
  for (int i = 0; i < DD; ++i) S[i] = T(-2) * S[i];
  const bool ok = chol_lower<T, D>(S, rinv);
  load_vec<T, D>(off, th_lin + idx * D);
  trsv_lower<T, D>(S, rinv, off);
  trsv_lower_t<T, D>(S, rinv, off);
  store_vec<T, D>(out_off + idx * D, off);
  if (k > 0) {
    T A[DD];
    load_vec<T, DD>(A, th_sub + (c * (Tn - 1) + k - 1) * DD);
    trsm_left_lower<T, D>(S, rinv, A);
    trsm_left_lower_t<T, D>(S, rinv, A);
    store_vec<T, DD>(out_a + (c * (Tn - 1) + k - 1) * DD, A);
  }
  chol_inverse<T, D>(Qc, S, rinv);
  chol_lower<T, D>(Qc, r2);
  zero_upper<T, D>(Qc);
  store_vec<T, DD>(out_chol + idx * DD, Qc);
  if (info && !ok) atomicMax(info + c, (int32_t)(k + 1));
}

}  // namespace mf
