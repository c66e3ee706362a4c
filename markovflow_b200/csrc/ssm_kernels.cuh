// StateSpaceModel kernels (reference markovflow/state_space_model.py): precision blocks, affine
// recurrence (marginal means / sampling), marginal moments, log-density, KL divergence.
//
// SSM parameter layout (reference state_space_model.py:74-122), chain-contiguous:
//   mu0 [B,D], chol_p0 [B,D,D], a [B,T-1,D,D], b [B,T-1,D], chol_q [B,T-1,D,D]   (T = states)
// Sequential recurrences run one thread per chain with D x D blocks in registers; per-step maps
// run one thread per (chain, step).
#pragma once
#include <cstdint>

#include "smallmat.cuh"
#include "smallrng.cuh"

namespace mf {

// Running product with the binary exponent peeled off after every multiply: the logarithm of a
// long product costs ONE log at the end instead of one per factor (log is ~277 cycles in FP64).
template <typename T>
struct LogProd {
  T prod;
  int esum;
  __device__ __forceinline__ void init() { prod = T(1); esum = 0; }
  // multiply without renormalising: the caller must peel() before the product can overflow
  __device__ __forceinline__ void mul_lazy(T v) { prod *= v; }
  __device__ __forceinline__ void mul(T v) {
    prod *= v;
    peel();
  }
  __device__ __forceinline__ void peel() {
    if (sizeof(T) == 8) {
      const int hi = __double2hiint((double)prod);
      const int e = ((hi >> 20) & 0x7ff) - 1023;
      esum += e;
      prod = (T)__hiloint2double(hi - (e << 20), __double2loint((double)prod));
    } else {
      const int bits = __float_as_int((float)prod);
      const int e = ((bits >> 23) & 0xff) - 127;
      esum += e;
      prod = (T)__int_as_float(bits - (e << 23));
    }
  }
  // log |product|
  __device__ __forceinline__ T log_abs() const {
    return Num<T>::log(Num<T>::abs(prod)) + T(esum) * T(0.6931471805599453094);
  }
};

template <typename T, int D>
__device__ __forceinline__ void diag_rcp(const T* __restrict__ l, T* __restrict__ rinv) {
#pragma unroll
  for (int j = 0; j < D; ++j) rinv[j] = Num<T>::rcp(l[j * D + j]);
}

// out(full) = L L^T for a lower-triangular L (upper triangle of l ignored)
template <typename T, int D>
__device__ __forceinline__ void llt(T* __restrict__ out, const T* __restrict__ l) {
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      T v = T(0);
#pragma unroll
      for (int q = 0; q <= j; ++q) v = Num<T>::fma(l[i * D + q], l[j * D + q], v);
      out[i * D + j] = v;
      out[j * D + i] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// _build_precision (state_space_model.py:431-483), optionally + H^T R^-1 H (kalman_filter.py:85-101)
//   diag_k = Qcat_k^{-1} + A_k^T Q_k^{-1} A_k (last block: Q^{-1} only),  sub_k = -Q_k^{-1} A_k
// h [Bh,T,m,D] (Bh = B or 1), r_inv [Tr,m,m] (Tr = T or 1) are optional.
// ---------------------------------------------------------------------------------------------
template <typename T, int D>
__global__ void __launch_bounds__(128)
ssm_build_precision_kernel(const T* __restrict__ chol_p0, const T* __restrict__ a,
                           const T* __restrict__ chol_q, const T* __restrict__ h,
                           const T* __restrict__ r_inv, T* __restrict__ out_diag,
                           T* __restrict__ out_sub, int64_t B, int64_t Tn, int m, int64_t Bh,
                           int64_t Tr) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= B * Tn) return;
  constexpr int DD = D * D;
  const int64_t c = idx / Tn, k = idx % Tn;
  T L[DD], Qi[DD], rinv[D];
  load_vec<T, DD>(L, k == 0 ? chol_p0 + c * DD : chol_q + (c * (Tn - 1) + k - 1) * DD);
  diag_rcp<T, D>(L, rinv);
  chol_inverse<T, D>(Qi, L, rinv);
  if (k + 1 < Tn) {
    T A[DD], X[DD];
    load_vec<T, DD>(L, chol_q + (c * (Tn - 1) + k) * DD);
    load_vec<T, DD>(A, a + (c * (Tn - 1) + k) * DD);
    diag_rcp<T, D>(L, rinv);
#pragma unroll
    for (int i = 0; i < DD; ++i) X[i] = A[i];
    trsm_left_lower<T, D>(L, rinv, X);
    trsm_left_lower_t<T, D>(L, rinv, X);  // X = Q_k^{-1} A_k
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
#pragma unroll
    for (int i = 0; i < DD; ++i) X[i] = -X[i];
    store_vec<T, DD>(out_sub + (c * (Tn - 1) + k) * DD, X);
  }
  if (h) {
    const T* hp = h + (((Bh == 1) ? 0 : c) * Tn + k) * (int64_t)m * D;
    const T* rp = r_inv + ((Tr == 1) ? 0 : k) * (int64_t)m * m;
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < m; ++j) {
        const T rij = rp[i * m + j];
#pragma unroll
        for (int p = 0; p < D; ++p) {
          const T hr = hp[i * D + p] * rij;
#pragma unroll
          for (int q = 0; q < D; ++q) Qi[p * D + q] = Num<T>::fma(hr, hp[j * D + q], Qi[p * D + q]);
        }
      }
  }
  store_vec<T, DD>(out_diag + idx * DD, Qi);
}

// ---------------------------------------------------------------------------------------------
// Affine recurrence  x_0 = mu0 (+ L0 e_0),  x_k = A_{k-1} x_{k-1} + b_{k-1} (+ Lq_{k-1} e_k)
// = a_inv_block.solve(...) of marginal_means (state_space_model.py:231-251) and sample (:298-324).
// Output chain c uses SSM chain c % Bm (leading sample dims).  eps may be NULL (means).
// ---------------------------------------------------------------------------------------------
// the standard-normal stream of ChainRng written out: eps [n,T,D] (thread per (trajectory, step))
template <typename T, int D>
__global__ void __launch_bounds__(128)
philox_normal_kernel(T* __restrict__ out, int64_t n, int64_t Tn, unsigned long long seed) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= n * Tn) return;
  ChainRng rng;
  rng.init(seed, idx / Tn, idx % Tn, D);
  T e[D];
  rng.template draw<T, D>(e);
  store_vec<T, D>(out + idx * D, e);
}

// the same stream for a run-time state dimension (8 < D <= 32: the draws of mf_ssm_sample above the register kernels)
template <typename T>
__global__ void __launch_bounds__(128)
philox_normal_dyn_kernel(T* __restrict__ out, int64_t n, int64_t Tn, int d, unsigned long long seed) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= n * Tn) return;
  ChainRng rng;
  rng.init(seed, idx / Tn, idx % Tn, d);
  for (int i = 0; i < d; i += 2) {
    const double2 z = curand_normal2_double(&rng.st);
    out[idx * d + i] = (T)z.x;
    if (i + 1 < d) out[idx * d + i + 1] = (T)z.y;
  }
}

template <typename T, int D>
__global__ void __launch_bounds__(32)
ssm_affine_scan_kernel(const T* __restrict__ mu0, const T* __restrict__ chol_p0,
                       const T* __restrict__ a, const T* __restrict__ b,
                       const T* __restrict__ chol_q, const T* __restrict__ eps,
                       T* __restrict__ out, int64_t n, int64_t Bm, int64_t Tn, int use_rng,
                       unsigned long long seed) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= n) return;
  constexpr int DD = D * D;
  const int64_t cm = c % Bm;
  const T* ap = a + cm * (Tn - 1) * DD;
  const T* bp = b + cm * (Tn - 1) * D;
  const T* qp = chol_q + cm * (Tn - 1) * DD;
  const T* ep = eps ? eps + c * Tn * D : nullptr;
  T* op = out + c * Tn * D;
  T x[D], A[DD], L[DD], off[D], e[D];
  ChainRng rng;
  if (use_rng) rng.init(seed, c, 0, D);
  load_vec<T, D>(x, mu0 + cm * D);
  if (ep || use_rng) {
    load_vec<T, DD>(L, chol_p0 + cm * DD);
    zero_upper<T, D>(L);
    if (use_rng) rng.template draw<T, D>(e);
    else load_vec<T, D>(e, ep);
    gemv_add<T, D>(x, L, e);
  }
  store_vec<T, D>(op, x);
  for (int64_t k = 1; k < Tn; ++k) {
    load_vec<T, DD>(A, ap + (k - 1) * DD);
    load_vec<T, D>(off, bp + (k - 1) * D);
    if (ep || use_rng) {
      load_vec<T, DD>(L, qp + (k - 1) * DD);
      zero_upper<T, D>(L);
      if (use_rng) rng.template draw<T, D>(e);
      else load_vec<T, D>(e, ep + k * D);
      gemv_add<T, D>(off, L, e);
    }
    gemv_add<T, D>(off, A, x);
#pragma unroll
    for (int i = 0; i < D; ++i) x[i] = off[i];
    store_vec<T, D>(op + k * D, x);
  }
}

// ---------------------------------------------------------------------------------------------
// Marginal moments by the forward recursion  mu_{k+1} = A_k mu_k + b_k,
// P_{k+1} = A_k P_k A_k^T + Q_k  (what marginal_means/marginal_covariances equal,
// state_space_model.py:231-262; tests/unit/test_state_space_model.py:78-89) and the lag-one blocks
// Sigma_{k+1,k} = A_k P_k (subsequent_covariances :326-341).  Any output may be NULL.
// ---------------------------------------------------------------------------------------------
template <typename T, int D>
__global__ void __launch_bounds__(32)
ssm_marginals_kernel(const T* __restrict__ mu0, const T* __restrict__ chol_p0,
                     const T* __restrict__ a, const T* __restrict__ b,
                     const T* __restrict__ chol_q, T* __restrict__ out_mean,
                     T* __restrict__ out_cov, T* __restrict__ out_sub, int64_t B, int64_t Tn) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= B) return;
  constexpr int DD = D * D;
  const T* ap = a + c * (Tn - 1) * DD;
  const T* bp = b + c * (Tn - 1) * D;
  const T* qp = chol_q + c * (Tn - 1) * DD;
  T mu[D], P[DD], A[DD], L[DD], AP[DD], off[D];
  load_vec<T, D>(mu, mu0 + c * D);
  load_vec<T, DD>(L, chol_p0 + c * DD);
  llt<T, D>(P, L);
  if (out_mean) store_vec<T, D>(out_mean + c * Tn * D, mu);
  if (out_cov) store_vec<T, DD>(out_cov + c * Tn * DD, P);
  for (int64_t k = 0; k + 1 < Tn; ++k) {
    load_vec<T, DD>(A, ap + k * DD);
    load_vec<T, D>(off, bp + k * D);
    load_vec<T, DD>(L, qp + k * DD);
    gemm<T, D>(AP, A, P);
    if (out_sub) store_vec<T, DD>(out_sub + (c * (Tn - 1) + k) * DD, AP);
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
    gemv_add<T, D>(off, A, mu);
#pragma unroll
    for (int i = 0; i < D; ++i) mu[i] = off[i];
    if (out_mean) store_vec<T, D>(out_mean + (c * Tn + k + 1) * D, mu);
    if (out_cov) store_vec<T, DD>(out_cov + (c * Tn + k + 1) * DD, P);
  }
}

// ---------------------------------------------------------------------------------------------
// log_pdf (state_space_model.py:485-526): grid (nseg, n); each block sums the factors of one time
// segment of one trajectory; trajectories use SSM chain c % Bm.
// ---------------------------------------------------------------------------------------------
template <typename T, int D>
__global__ void __launch_bounds__(128)
ssm_log_pdf_kernel(const T* __restrict__ mu0, const T* __restrict__ chol_p0,
                   const T* __restrict__ a, const T* __restrict__ b, const T* __restrict__ chol_q,
                   const T* __restrict__ states, T* __restrict__ out, int64_t Bm, int64_t Tn,
                   int64_t seg_len, int use_atomic, int64_t c_base) {
  constexpr int DD = D * D;
  const int64_t c = c_base + blockIdx.y;
  const int64_t cm = c % Bm;
  const int64_t k0 = blockIdx.x * seg_len;
  const int64_t k1 = (k0 + seg_len < Tn) ? k0 + seg_len : Tn;
  const T* xp = states + c * Tn * D;
  T acc = T(0);
  for (int64_t k = k0 + threadIdx.x; k < k1; k += blockDim.x) {
    T L[DD], r[D], rinv[D];
    load_vec<T, D>(r, xp + k * D);
    if (k == 0) {
      T m0[D];
      load_vec<T, D>(m0, mu0 + cm * D);
#pragma unroll
      for (int i = 0; i < D; ++i) r[i] -= m0[i];
      load_vec<T, DD>(L, chol_p0 + cm * DD);
    } else {
      T A[DD], off[D], xprev[D];
      load_vec<T, DD>(A, a + (cm * (Tn - 1) + k - 1) * DD);
      load_vec<T, D>(off, b + (cm * (Tn - 1) + k - 1) * D);
      load_vec<T, D>(xprev, xp + (k - 1) * D);
#pragma unroll
      for (int i = 0; i < D; ++i) r[i] -= off[i];
      gemv_sub<T, D>(r, A, xprev);
      load_vec<T, DD>(L, chol_q + (cm * (Tn - 1) + k - 1) * DD);
    }
    diag_rcp<T, D>(L, rinv);
    trsv_lower<T, D>(L, rinv, r);
    T q = T(0), dprod = T(1);
#pragma unroll
    for (int i = 0; i < D; ++i) {
      q = Num<T>::fma(r[i], r[i], q);
      dprod *= L[i * D + i];
    }
    acc += T(-0.5) * q - Num<T>::log(Num<T>::abs(dprod)) - T(0.5 * D * 1.8378770664093454836);
  }
  __shared__ T red[4];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    T s = T(0);
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    if (use_atomic) atomicAdd(out + c, s); else out[c] = s;
  }
}

// ---------------------------------------------------------------------------------------------
// KL(q || p) between two state-space models (state_space_model.py:528-593), one pass, one thread
// per chain, as the chain-rule sum
//   KL = KL(q_0 || p_0) + sum_k E_{q(x_k)} KL( q(x_{k+1}|x_k) || p(x_{k+1}|x_k) )
// (identical in exact arithmetic to the reference's trace / Mahalanobis / log-det expression, with
// no cancellation between O(T D) terms).  q's marginal mean/covariance are carried along.
// ---------------------------------------------------------------------------------------------
template <typename T, int D>
__device__ __forceinline__ T kl_gauss_term(const T* __restrict__ Lp, const T* __restrict__ Lq,
                                           const T* __restrict__ dA, const T* __restrict__ dmean,
                                           const T* __restrict__ P, LogProd<T>& ratio) {
  // 0.5 [ |Lp^-1 Lq|_F^2 + tr(G P G^T) + |Lp^-1 dmean|^2 - D ],  G = Lp^-1 dA ; log-dets go to ratio
  constexpr int DD = D * D;
  T rinv[D], W[DD], e[D];
  diag_rcp<T, D>(Lp, rinv);
#pragma unroll
  for (int i = 0; i < DD; ++i) W[i] = Lq[i];
  zero_upper<T, D>(W);
  trsm_left_lower<T, D>(Lp, rinv, W);
  T s = T(0);
#pragma unroll
  for (int i = 0; i < DD; ++i) s = Num<T>::fma(W[i], W[i], s);
#pragma unroll
  for (int i = 0; i < D; ++i) e[i] = dmean[i];
  trsv_lower<T, D>(Lp, rinv, e);
#pragma unroll
  for (int i = 0; i < D; ++i) s = Num<T>::fma(e[i], e[i], s);
  if (dA) {
    T G[DD], GP[DD];
#pragma unroll
    for (int i = 0; i < DD; ++i) G[i] = dA[i];
    trsm_left_lower<T, D>(Lp, rinv, G);
    gemm<T, D>(GP, G, P);
#pragma unroll
    for (int i = 0; i < DD; ++i) s = Num<T>::fma(GP[i], G[i], s);
  }
#pragma unroll
  for (int i = 0; i < D; ++i) ratio.mul(Lp[i * D + i] * Num<T>::rcp(Lq[i * D + i]));
  return T(0.5) * (s - T(D));
}

template <typename T, int D>
__global__ void __launch_bounds__(32)
ssm_kl_kernel(const T* __restrict__ q_mu0, const T* __restrict__ q_chol_p0,
              const T* __restrict__ q_a, const T* __restrict__ q_b, const T* __restrict__ q_chol_q,
              const T* __restrict__ p_mu0, const T* __restrict__ p_chol_p0,
              const T* __restrict__ p_a, const T* __restrict__ p_b, const T* __restrict__ p_chol_q,
              T* __restrict__ out, int64_t B, int64_t Tn) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= B) return;
  constexpr int DD = D * D;
  T mu[D], P[DD], Lq[DD], Lp[DD], dm[D];
  LogProd<T> ratio;
  ratio.init();
  load_vec<T, D>(mu, q_mu0 + c * D);
  load_vec<T, DD>(Lq, q_chol_p0 + c * DD);
  load_vec<T, DD>(Lp, p_chol_p0 + c * DD);
  load_vec<T, D>(dm, p_mu0 + c * D);
#pragma unroll
  for (int i = 0; i < D; ++i) dm[i] = mu[i] - dm[i];
  T kl = kl_gauss_term<T, D>(Lp, Lq, nullptr, dm, nullptr, ratio);
  llt<T, D>(P, Lq);
  const int64_t off = c * (Tn - 1);
  for (int64_t k = 0; k + 1 < Tn; ++k) {
    T Aq[DD], Ap[DD], bq[D], bp[D], AP[DD];
    load_vec<T, DD>(Aq, q_a + (off + k) * DD);
    load_vec<T, DD>(Ap, p_a + (off + k) * DD);
    load_vec<T, D>(bq, q_b + (off + k) * D);
    load_vec<T, D>(bp, p_b + (off + k) * D);
    load_vec<T, DD>(Lq, q_chol_q + (off + k) * DD);
    load_vec<T, DD>(Lp, p_chol_q + (off + k) * DD);
#pragma unroll
    for (int i = 0; i < DD; ++i) Ap[i] = Aq[i] - Ap[i];  // dA
#pragma unroll
    for (int i = 0; i < D; ++i) dm[i] = bq[i] - bp[i];
    gemv_add<T, D>(dm, Ap, mu);  // dA mu + db
    kl += kl_gauss_term<T, D>(Lp, Lq, Ap, dm, P, ratio);
    // advance q's marginal
    gemm<T, D>(AP, Aq, P);
    llt<T, D>(P, Lq);
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int j = 0; j <= i; ++j) {
        T v = P[i * D + j];
#pragma unroll
        for (int q = 0; q < D; ++q) v = Num<T>::fma(AP[i * D + q], Aq[j * D + q], v);
        P[i * D + j] = v;
        P[j * D + i] = v;
      }
    gemv_add<T, D>(bq, Aq, mu);
#pragma unroll
    for (int i = 0; i < D; ++i) mu[i] = bq[i];
  }
  out[c] = kl + ratio.log_abs();  // + sum log(diag Lp / diag Lq) = 0.5 (log|Qp| - log|Qq|)
}

// Cholesky of every D x D block; an all-zero block maps to zero
// (state_space_model_from_covariances.cholesky_or_zero, state_space_model.py:634-656).
template <typename T, int D>
__global__ void __launch_bounds__(128)
block_cholesky_or_zero_kernel(const T* __restrict__ cov, T* __restrict__ out,
                              int32_t* __restrict__ info, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  constexpr int DD = D * D;
  T S[DD], rinv[D];
  load_vec<T, DD>(S, cov + i * DD);
  bool all_zero = true;
#pragma unroll
  for (int q = 0; q < DD; ++q) all_zero = all_zero && (S[q] == T(0));
  bool ok = true;
  if (!all_zero) {
    ok = chol_lower<T, D>(S, rinv);
    zero_upper<T, D>(S);
  }
  store_vec<T, DD>(out + i * DD, S);
  if (info && !ok) atomicMax(info, (int32_t)(i < 2147483647 ? i + 1 : 2147483647));
}

// out_k = chol( (L_k L_k^T)^{-1} ) for every block: the process-noise factors of the posterior
// state-space model from the factors of D in U D U^T (kalman_filter.py:170-174).
template <typename T, int D>
__global__ void __launch_bounds__(128)
block_chol_of_inverse_kernel(const T* __restrict__ l, T* __restrict__ out, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  constexpr int DD = D * D;
  T L[DD], S[DD], rinv[D];
  load_vec<T, DD>(L, l + i * DD);
  diag_rcp<T, D>(L, rinv);
  chol_inverse<T, D>(S, L, rinv);
  chol_lower<T, D>(S, rinv);
  zero_upper<T, D>(S);
  store_vec<T, DD>(out + i * DD, S);
}

}  // namespace mf
