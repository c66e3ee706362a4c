// C-ABI entry points for the conditionals on top of the Gauss-Markov marginals
// (include/markovflow_b200.h; reference markovflow/conditionals.py): pairwise marginals of subsequent
// states, conditional statistics p(x_t | x_-, x_+) from transition statistics, and the prediction at
// new time points.  All three are per-(chain, point) maps without recursion.
#include "dispatch.cuh"
#include "mid_api.h"
#include "smallmat.cuh"

using namespace mf;

namespace {

// joint of (x_{k-1}, x_k), k = 0..T, with the initial state at both ends (conditionals.py:423-485):
//   mean_k = [m_{k-1}, m_k],  cov_k = [[S_{k-1}, C_{k-1}^T], [C_{k-1}, S_k]],  C_k = A_k S_k (lag one),
//   C = 0 for the two pairs that involve the initial state.  One thread per (chain, pair).
template <typename T, int D>
__global__ void __launch_bounds__(128)
pairwise_marginals_kernel(const T* __restrict__ mean, const T* __restrict__ cov,
                          const T* __restrict__ sub, const T* __restrict__ init_mean,
                          const T* __restrict__ init_cov, int64_t init_batch, T* __restrict__ o_mean,
                          T* __restrict__ o_cov, int64_t B, int64_t Tn) {
  constexpr int DD = D * D, D2 = 2 * D;
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= B * (Tn + 1)) return;
  const int64_t c = idx / (Tn + 1), k = idx % (Tn + 1);
  const int64_t ci = init_batch == 1 ? 0 : c;
  T m0[D], m1[D], S0[DD], S1[DD], C[DD];
  if (k == 0) {
    load_vec<T, D>(m0, init_mean + ci * D);
    load_vec<T, DD>(S0, init_cov + ci * DD);
  } else {
    load_vec<T, D>(m0, mean + (c * Tn + k - 1) * D);
    load_vec<T, DD>(S0, cov + (c * Tn + k - 1) * DD);
  }
  if (k == Tn) {
    load_vec<T, D>(m1, init_mean + ci * D);
    load_vec<T, DD>(S1, init_cov + ci * DD);
  } else {
    load_vec<T, D>(m1, mean + (c * Tn + k) * D);
    load_vec<T, DD>(S1, cov + (c * Tn + k) * DD);
  }
  if (k >= 1 && k < Tn) {
    load_vec<T, DD>(C, sub + (c * (Tn - 1) + k - 1) * DD);
  } else {
#pragma unroll
    for (int i = 0; i < DD; ++i) C[i] = T(0);
  }
  T* om = o_mean + idx * D2;
  T* oc = o_cov + idx * D2 * D2;
#pragma unroll
  for (int i = 0; i < D; ++i) {
    om[i] = m0[i];
    om[D + i] = m1[i];
  }
#pragma unroll
  for (int r = 0; r < D; ++r)
#pragma unroll
    for (int q = 0; q < D; ++q) {
      oc[r * D2 + q] = S0[r * D + q];
      oc[r * D2 + D + q] = C[q * D + r];
      oc[(D + r) * D2 + q] = C[r * D + q];
      oc[(D + r) * D2 + D + q] = S1[r * D + q];
    }
}

// p(x_t | x_-, x_+) = N(D_t x_- + E_t x_+, T_t) from the transitions (A_mt, Q_mt) into t and
// (A_tp, Q_tp) out of t (conditionals.py:128-205):  Q_mp = Q_tp + A_tp Q_mt A_tp^T = L L^T,
// V = L^{-1} A_tp Q_mt,  E = (L^{-T} V)^T,  D = A_mt - E A_tp A_mt,  T = Q_mt - V^T V
// (or the precision T^{-1} = Q_mt^{-1} + A_tp^T Q_tp^{-1} A_tp).  One thread per point.
template <typename T, int D>
__global__ void __launch_bounds__(128)
conditional_statistics_kernel(const T* __restrict__ a_mt, const T* __restrict__ q_mt,
                              const T* __restrict__ a_tp, const T* __restrict__ q_tp,
                              T* __restrict__ o_p, T* __restrict__ o_t, int32_t* __restrict__ info,
                              int return_precision, int64_t N) {
  constexpr int DD = D * D, D2 = 2 * D;
  const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (n >= N) return;
  T Am[DD], Qm[DD], Ap[DD], Qp[DD], AQ[DD], L[DD], rinv[D], V[DD], E[DD], Tm[DD];
  load_vec<T, DD>(Am, a_mt + n * DD);
  load_vec<T, DD>(Qm, q_mt + n * DD);
  load_vec<T, DD>(Ap, a_tp + n * DD);
  load_vec<T, DD>(Qp, q_tp + n * DD);
  gemm<T, D>(AQ, Ap, Qm);       // A_tp Q_mt
  gemm_nt<T, D>(L, AQ, Ap);     // A_tp Q_mt A_tp^T
#pragma unroll
  for (int i = 0; i < DD; ++i) L[i] += Qp[i];
  bool ok = chol_lower<T, D>(L, rinv);
#pragma unroll
  for (int i = 0; i < DD; ++i) V[i] = AQ[i];
  trsm_left_lower<T, D>(L, rinv, V);  // V = L^{-1} A_tp Q_mt
#pragma unroll
  for (int i = 0; i < DD; ++i) E[i] = V[i];
  trsm_left_lower_t<T, D>(L, rinv, E);  // L^{-T} V = E^T
  T Et[DD], EA[DD], Dm[DD];
#pragma unroll
  for (int r = 0; r < D; ++r)
#pragma unroll
    for (int q = 0; q < D; ++q) Et[r * D + q] = E[q * D + r];
  gemm<T, D>(EA, Et, Ap);
  gemm<T, D>(Dm, EA, Am);
#pragma unroll
  for (int i = 0; i < DD; ++i) Dm[i] = Am[i] - Dm[i];
  if (return_precision) {
    T Lm[DD], Lp[DD], r1[D], r2[D], W[DD];
#pragma unroll
    for (int i = 0; i < DD; ++i) {
      Lm[i] = Qm[i];
      Lp[i] = Qp[i];
    }
    ok = chol_lower<T, D>(Lm, r1) && ok;
    ok = chol_lower<T, D>(Lp, r2) && ok;
    chol_inverse<T, D>(Tm, Lm, r1);  // Q_mt^{-1}
#pragma unroll
    for (int i = 0; i < DD; ++i) W[i] = Ap[i];
    trsm_left_lower<T, D>(Lp, r2, W);  // L_tp^{-1} A_tp
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
      for (int q = 0; q < D; ++q) {
        T v = Tm[r * D + q];
#pragma unroll
        for (int s = 0; s < D; ++s) v = Num<T>::fma(W[s * D + r], W[s * D + q], v);
        Tm[r * D + q] = v;
      }
  } else {
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
      for (int q = 0; q < D; ++q) {
        T v = Qm[r * D + q];
#pragma unroll
        for (int s = 0; s < D; ++s) v = Num<T>::fma(-V[s * D + r], V[s * D + q], v);
        Tm[r * D + q] = v;
      }
  }
  T* op = o_p + n * D * D2;
#pragma unroll
  for (int r = 0; r < D; ++r)
#pragma unroll
    for (int q = 0; q < D; ++q) {
      op[r * D2 + q] = Dm[r * D + q];
      op[r * D2 + D + q] = Et[r * D + q];
    }
  store_vec<T, DD>(o_t + n * DD, Tm);
  if (info) info[n] = ok ? 0 : 1;
}

// mean = P m[idx],  cov = T (+ P S[idx] P^T)  (conditionals.py:29-83, base_conditional_predict
// :380-420 with the gather by insertion index fused in).  One thread per (chain, point).
template <typename T, int D>
__global__ void __launch_bounds__(128)
conditional_predict_kernel(const T* __restrict__ proj, const T* __restrict__ tcov,
                           const T* __restrict__ pair_means, const T* __restrict__ pair_covs,
                           const int64_t* __restrict__ indices, T* __restrict__ o_mean,
                           T* __restrict__ o_cov, int64_t B, int64_t N, int64_t M) {
  constexpr int DD = D * D, D2 = 2 * D;
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= B * N) return;
  const int64_t c = idx / N;
  int64_t j = indices ? indices[idx] : idx % N;
  if (j < 0) j = 0;
  if (j > M - 1) j = M - 1;
  const T* P = proj + idx * D * D2;
  const T* m = pair_means + (c * M + j) * D2;
  T mean[D];
#pragma unroll
  for (int r = 0; r < D; ++r) {
    T v = T(0);
#pragma unroll
    for (int q = 0; q < D2; ++q) v = Num<T>::fma(P[r * D2 + q], m[q], v);
    mean[r] = v;
  }
  store_vec<T, D>(o_mean + idx * D, mean);
  T cov[DD];
  load_vec<T, DD>(cov, tcov + idx * DD);
  if (pair_covs) {
    const T* S = pair_covs + (c * M + j) * D2 * D2;
    T PS[D * D2];
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
      for (int q = 0; q < D2; ++q) {
        T v = T(0);
#pragma unroll
        for (int s = 0; s < D2; ++s) v = Num<T>::fma(P[r * D2 + s], S[s * D2 + q], v);
        PS[r * D2 + q] = v;
      }
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
      for (int q = 0; q < D; ++q) {
        T v = cov[r * D + q];
#pragma unroll
        for (int s = 0; s < D2; ++s) v = Num<T>::fma(PS[r * D2 + s], P[q * D2 + s], v);
        cov[r * D + q] = v;
      }
  }
  store_vec<T, DD>(o_cov + idx * DD, cov);
}

}  // namespace

extern "C" {

int mf_pairwise_marginals(int dtype, const void* mean, const void* cov, const void* sub,
                          const void* init_mean, const void* init_cov, int64_t init_batch,
                          void* out_mean, void* out_cov, int64_t B, int64_t T, int64_t D, void* stream) {
  if (B < 0 || T < 1 || D < 1 || (init_batch != 1 && init_batch != B)) return MF_ERR_BAD_ARG;
  if (B == 0) return MF_OK;
  if (!mean || !cov || (T > 1 && !sub) || !init_mean || !init_cov || !out_mean || !out_cov)
    return MF_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  if (mid_dim(D))
    return mid_pairwise_marginals(dtype, mean, cov, sub, init_mean, init_cov, init_batch, out_mean, out_cov, B, T, D, s);
  return dispatch_small(dtype, D, [&](auto tt, auto dd) {
    using Tp = typename decltype(tt)::type;
    constexpr int kD = decltype(dd)::value;
    pairwise_marginals_kernel<Tp, kD><<<grid_for(B * (T + 1), 128), 128, 0, s>>>(
        (const Tp*)mean, (const Tp*)cov, (const Tp*)sub, (const Tp*)init_mean, (const Tp*)init_cov,
        init_batch, (Tp*)out_mean, (Tp*)out_cov, B, T);
    return check_launch();
  });
}

int mf_conditional_statistics(int dtype, const void* a_mt, const void* q_mt, const void* a_tp,
                              const void* q_tp, void* out_p, void* out_t, int32_t* info,
                              int return_precision, int64_t N, int64_t D, void* stream) {
  if (N < 0 || D < 1) return MF_ERR_BAD_ARG;
  if (N == 0) return MF_OK;
  if (!a_mt || !q_mt || !a_tp || !q_tp || !out_p || !out_t) return MF_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  if (mid_dim(D))
    return mid_conditional_statistics(dtype, a_mt, q_mt, a_tp, q_tp, out_p, out_t, info, return_precision, N, D, s);
  return dispatch_small(dtype, D, [&](auto tt, auto dd) {
    using Tp = typename decltype(tt)::type;
    constexpr int kD = decltype(dd)::value;
    conditional_statistics_kernel<Tp, kD><<<grid_for(N, 128), 128, 0, s>>>(
        (const Tp*)a_mt, (const Tp*)q_mt, (const Tp*)a_tp, (const Tp*)q_tp, (Tp*)out_p, (Tp*)out_t,
        info, return_precision, N);
    return check_launch();
  });
}

int mf_conditional_predict(int dtype, const void* proj, const void* tcov, const void* pair_means,
                           const void* pair_covs, const int64_t* indices, void* out_mean,
                           void* out_cov, int64_t B, int64_t N, int64_t M, int64_t D, void* stream) {
  if (B < 0 || N < 0 || M < 1 || D < 1) return MF_ERR_BAD_ARG;
  if (B == 0 || N == 0) return MF_OK;
  if (!proj || !tcov || !pair_means || !out_mean || !out_cov) return MF_ERR_BAD_ARG;
  if (!indices && N != M) return MF_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  if (mid_dim(D))
    return mid_conditional_predict(dtype, proj, tcov, pair_means, pair_covs, indices, out_mean, out_cov, B, N, M, D, s);
  return dispatch_small(dtype, D, [&](auto tt, auto dd) {
    using Tp = typename decltype(tt)::type;
    constexpr int kD = decltype(dd)::value;
    conditional_predict_kernel<Tp, kD><<<grid_for(B * N, 128), 128, 0, s>>>(
        (const Tp*)proj, (const Tp*)tcov, (const Tp*)pair_means, (const Tp*)pair_covs, indices,
        (Tp*)out_mean, (Tp*)out_cov, B, N, M);
    return check_launch();
  });
}

}  // extern "C"
