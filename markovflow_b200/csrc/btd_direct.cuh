// Thread-per-chain block-tridiagonal kernels, "direct" flavour: every thread streams its own
// chain straight from global memory (register double-buffered).  These are the simple, always
// correct kernels; the shared-memory staged flavour in btd_staged.cuh is the fast path for the
// Cholesky(+solve) sweep and falls back to these when its alignment/shape preconditions fail.
#pragma once
#include <cstdint>

#include "smallmat.cuh"

namespace mf {

// ---------------------------------------------------------------------------------------------
// Cholesky sweep, optionally fused with the forward solve and the log-determinant.
//   Ld_0 = chol(D_0);  Ls_k = A_k Ld_k^{-T};  Ld_{k+1} = chol(D_{k+1} - Ls_k Ls_k^T)
//   x_k  = Ld_k^{-1} (b_k - Ls_{k-1} x_{k-1})
// (block form of banded cholesky_band / solve_triang_mat, reference block_tri_diag.py:436,350)
// ---------------------------------------------------------------------------------------------
template <typename T, int D>
__global__ void __launch_bounds__(32)
btd_chol_direct_kernel(const T* __restrict__ diag, const T* __restrict__ sub,
                       const T* __restrict__ rhs, T* od, T* os, T* ox, T* __restrict__ logdet,
                       int32_t* __restrict__ info, int64_t B, int64_t Tn) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= B) return;
  constexpr int DD = D * D;
  const T* dp = diag + c * Tn * DD;
  const T* sp = sub ? sub + c * (Tn - 1) * DD : nullptr;
  const T* rp = rhs ? rhs + c * Tn * D : nullptr;
  T* odp = od + c * Tn * DD;
  T* osp = os ? os + c * (Tn - 1) * DD : nullptr;
  T* oxp = ox ? ox + c * Tn * D : nullptr;

  T S[DD], A[DD], r[D], nS[DD], nA[DD], nr[D], rinv[D];
  T acc = T(0);
  int32_t fail = 0;
  load_vec<T, DD>(S, dp);
  if (sp && Tn > 1) load_vec<T, DD>(A, sp);
  if (rp) load_vec<T, D>(r, rp);

  for (int64_t k = 0; k < Tn; ++k) {
    const bool has_next = k + 1 < Tn;
    if (has_next) {
      load_vec<T, DD>(nS, dp + (k + 1) * DD);
      if (sp && k + 2 < Tn) load_vec<T, DD>(nA, sp + (k + 1) * DD);
      if (rp) load_vec<T, D>(nr, rp + (k + 1) * D);
    }
    const bool ok = chol_lower<T, D>(S, rinv);
    if (!ok && fail == 0) fail = (int32_t)(k + 1);
    zero_upper<T, D>(S);
    store_vec<T, DD>(odp + k * DD, S);
    if (logdet) {
#pragma unroll
      for (int j = 0; j < D; ++j) acc += Num<T>::log(S[j * D + j]);
    }
    if (rp) {
      trsv_lower<T, D>(S, rinv, r);
      store_vec<T, D>(oxp + k * D, r);
    }
    if (has_next) {
      if (sp) {
        trsm_right_lower_t<T, D>(A, S, rinv);
        store_vec<T, DD>(osp + k * DD, A);
        syrk_sub_lower<T, D>(nS, A);
        if (rp) gemv_sub<T, D>(nr, A, r);
      }
#pragma unroll
      for (int i = 0; i < DD; ++i) { S[i] = nS[i]; A[i] = nA[i]; }
#pragma unroll
      for (int i = 0; i < D; ++i) r[i] = nr[i];
    }
  }
  if (logdet) logdet[c] = acc;
  if (info) info[c] = fail;
}

// ---------------------------------------------------------------------------------------------
// Triangular solve with a lower block-bidiagonal matrix (forward) or its transpose (backward).
// rhs chain c uses matrix chain c % Bm.
// ---------------------------------------------------------------------------------------------
template <typename T, int D>
__global__ void __launch_bounds__(32)
btd_solve_direct_kernel(const T* __restrict__ ld, const T* __restrict__ ls,
                        const T* __restrict__ rhs, T* out, int64_t n_rhs, int64_t Bm, int64_t Tn,
                        int transpose) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= n_rhs) return;
  constexpr int DD = D * D;
  const int64_t cm = c % Bm;
  const T* lp = ld ? ld + cm * Tn * DD : nullptr;
  const T* sp = ls ? ls + cm * (Tn - 1) * DD : nullptr;
  const T* rp = rhs + c * Tn * D;
  T* op = out + c * Tn * D;
  T L[DD], A[DD], x[D], r[D], rinv[D];
#pragma unroll
  for (int i = 0; i < D; ++i) x[i] = T(0);
  if (!transpose) {
    for (int64_t k = 0; k < Tn; ++k) {
      load_vec<T, D>(r, rp + k * D);
      if (sp && k > 0) {
        load_vec<T, DD>(A, sp + (k - 1) * DD);
        gemv_sub<T, D>(r, A, x);
      }
      if (ld) {  // ld == nullptr: identity diagonal blocks
        load_vec<T, DD>(L, lp + k * DD);
#pragma unroll
        for (int j = 0; j < D; ++j) rinv[j] = Num<T>::rcp(L[j * D + j]);
        trsv_lower<T, D>(L, rinv, r);
      }
#pragma unroll
      for (int i = 0; i < D; ++i) x[i] = r[i];
      store_vec<T, D>(op + k * D, x);
    }
  } else {
    for (int64_t k = Tn - 1; k >= 0; --k) {
      load_vec<T, D>(r, rp + k * D);
      if (sp && k + 1 < Tn) {
        load_vec<T, DD>(A, sp + k * DD);
        gemv_t_sub<T, D>(r, A, x);
      }
      if (ld) {
        load_vec<T, DD>(L, lp + k * DD);
#pragma unroll
        for (int j = 0; j < D; ++j) rinv[j] = Num<T>::rcp(L[j * D + j]);
        trsv_lower_t<T, D>(L, rinv, r);
      }
#pragma unroll
      for (int i = 0; i < D; ++i) x[i] = r[i];
      store_vec<T, D>(op + k * D, x);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Sparse inverse subset of (L L^T)^{-1}: diagonal (and optionally sub-diagonal) blocks.
//   Sigma_{T-1,T-1} = (Ld Ld^T)^{-1};  J_k = Ls_k Ld_k^{-1};  Sigma_{k+1,k} = -Sigma_{k+1,k+1} J_k;
//   Sigma_{kk} = (Ld_k Ld_k^T)^{-1} - J_k^T Sigma_{k+1,k}
// (block form of inverse_from_cholesky_band, reference block_tri_diag.py:331)
// ---------------------------------------------------------------------------------------------
template <typename T, int D>
__global__ void __launch_bounds__(32)
btd_inverse_subset_direct_kernel(const T* __restrict__ ld, const T* __restrict__ ls, T* od, T* os,
                                 int64_t B, int64_t Tn) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= B) return;
  constexpr int DD = D * D;
  const T* lp = ld + c * Tn * DD;
  const T* sp = ls ? ls + c * (Tn - 1) * DD : nullptr;
  T* odp = od + c * Tn * DD;
  T* osp = os ? os + c * (Tn - 1) * DD : nullptr;
  T L[DD], J[DD], sig[DD], loc[DD], ssub[DD], rinv[D];
#pragma unroll
  for (int i = 0; i < DD; ++i) sig[i] = T(0);
  for (int64_t k = Tn - 1; k >= 0; --k) {
    load_vec<T, DD>(L, lp + k * DD);
#pragma unroll
    for (int j = 0; j < D; ++j) rinv[j] = Num<T>::rcp(L[j * D + j]);
    chol_inverse<T, D>(loc, L, rinv);
    if (sp && k + 1 < Tn) {
      load_vec<T, DD>(J, sp + k * DD);
      trsm_right_lower<T, D>(J, L, rinv);  // J = Ls Ld^{-1}
      gemm<T, D>(ssub, sig, J);            // Sigma_{k+1,k+1} J
#pragma unroll
      for (int i = 0; i < DD; ++i) ssub[i] = -ssub[i];
      if (osp) store_vec<T, DD>(osp + k * DD, ssub);
      // loc(lower) -= J^T ssub
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
          T v = loc[i * D + j];
#pragma unroll
          for (int q = 0; q < D; ++q) v = Num<T>::fma(-J[q * D + i], ssub[q * D + j], v);
          loc[i * D + j] = v;
        }
      mirror_lower<T, D>(loc);
    }
    store_vec<T, DD>(odp + k * DD, loc);
#pragma unroll
    for (int i = 0; i < DD; ++i) sig[i] = loc[i];
  }
}

// ---------------------------------------------------------------------------------------------
// U D U^T factorisation, backwards (reference block_tri_diag.py:438-545):
//   cholD_{T-1} = chol(K_{T-1,T-1});  U_k^T = D_{k+1}^{-1} K_{k+1,k};
//   D_k = K_kk - K_{k+1,k}^T U_k^T;   cholD_k = chol(D_k)
// ---------------------------------------------------------------------------------------------
template <typename T, int D>
__global__ void __launch_bounds__(32)
btd_udu_direct_kernel(const T* __restrict__ diag, const T* __restrict__ sub, T* ou, T* ocd,
                      int32_t* __restrict__ info, int64_t B, int64_t Tn) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= B) return;
  constexpr int DD = D * D;
  const T* dp = diag + c * Tn * DD;
  const T* sp = sub + c * (Tn - 1) * DD;
  T* oup = ou + c * (Tn - 1) * DD;
  T* ocp = ocd + c * Tn * DD;
  T C[DD], K[DD], X[DD], rinv[D];
  int32_t fail = 0;
  load_vec<T, DD>(C, dp + (Tn - 1) * DD);
  bool ok = chol_lower<T, D>(C, rinv);
  if (!ok) fail = (int32_t)Tn;
  zero_upper<T, D>(C);
  store_vec<T, DD>(ocp + (Tn - 1) * DD, C);
  for (int64_t k = Tn - 2; k >= 0; --k) {
    load_vec<T, DD>(K, sp + k * DD);
#pragma unroll
    for (int i = 0; i < DD; ++i) X[i] = K[i];
    trsm_left_lower<T, D>(C, rinv, X);
    trsm_left_lower_t<T, D>(C, rinv, X);  // X = D_{k+1}^{-1} K_{k+1,k}
    store_vec<T, DD>(oup + k * DD, X);
    load_vec<T, DD>(C, dp + k * DD);
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int j = 0; j <= i; ++j) {
        T v = C[i * D + j];
#pragma unroll
        for (int q = 0; q < D; ++q) v = Num<T>::fma(-K[q * D + i], X[q * D + j], v);
        C[i * D + j] = v;
      }
    ok = chol_lower<T, D>(C, rinv);
    if (!ok && fail == 0) fail = (int32_t)(k + 1);
    zero_upper<T, D>(C);
    store_vec<T, DD>(ocp + k * DD, C);
  }
  if (info) info[c] = fail;
}

// ---------------------------------------------------------------------------------------------
// dense_mult: one thread per (rhs chain, block row); no recurrence.
// ---------------------------------------------------------------------------------------------
template <typename T, int D>
__global__ void __launch_bounds__(128)
btd_dense_mult_kernel(const T* __restrict__ diag, const T* __restrict__ sub,
                      const T* __restrict__ right, T* __restrict__ out, int64_t n_rhs, int64_t Bm,
                      int64_t Tn, int transpose, int symmetric) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= n_rhs * Tn) return;
  constexpr int DD = D * D;
  const int64_t c = idx / Tn, k = idx % Tn;
  const int64_t cm = c % Bm;
  const T* dp = diag + (cm * Tn + k) * DD;
  const T* xp = right + (c * Tn + k) * D;
  T M[DD], x[D], y[D];
  load_vec<T, DD>(M, dp);
  load_vec<T, D>(x, xp);
#pragma unroll
  for (int i = 0; i < D; ++i) y[i] = T(0);
  if (symmetric) {
    mirror_lower<T, D>(M);
    gemv_add<T, D>(y, M, x);
  } else {
    zero_upper<T, D>(M);
    if (transpose) gemv_t_add<T, D>(y, M, x); else gemv_add<T, D>(y, M, x);
  }
  if (sub) {
    const T* sp = sub + cm * (Tn - 1) * DD;
    if ((symmetric || !transpose) && k > 0) {  // + A_{k-1} x_{k-1}
      load_vec<T, DD>(M, sp + (k - 1) * DD);
      load_vec<T, D>(x, xp - D);
      gemv_add<T, D>(y, M, x);
    }
    if ((symmetric || transpose) && k + 1 < Tn) {  // + A_k^T x_{k+1}
      load_vec<T, DD>(M, sp + k * DD);
      load_vec<T, D>(x, xp + D);
      gemv_t_add<T, D>(y, M, x);
    }
  }
  store_vec<T, D>(out + (c * Tn + k) * D, y);
}

// abs_log_det: grid (nseg, B); each block reduces a time segment of one chain. Works for any D.
template <typename T>
__global__ void __launch_bounds__(256)
btd_abs_log_det_kernel(const T* __restrict__ ld, T* __restrict__ out, int64_t Tn, int D,
                       int64_t seg_len, int use_atomic) {
  const int64_t b = blockIdx.y;
  const int64_t k0 = blockIdx.x * seg_len;
  const int64_t k1 = (k0 + seg_len < Tn) ? k0 + seg_len : Tn;
  const int64_t n = (k1 - k0) * D;
  const T* base = ld + (b * Tn + k0) * D * D;
  T acc = T(0);
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const int64_t k = i / D;
    const int j = (int)(i % D);
    const T v = base[k * D * D + j * D + j];
    acc += T(0.5) * Num<T>::log(v * v);
  }
  __shared__ T red[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    T s = T(0);
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    if (use_atomic) atomicAdd(out + b, s); else out[b] = s;
  }
}

}  // namespace mf
