// Per-thread Kalman-filter arithmetic for the log-likelihood kernels (KalmanFilter.log_likelihood,
// reference markovflow/kalman_filter.py:184-255).
//
// The reference evaluates the marginal likelihood in SpInGP (banded precision) form; the value is
// that of the classical predict/update filter
//     log p(y) = sum_k -1/2 ( v_k^T S_k^{-1} v_k + log|S_k| + m log 2 pi )
// which is what these kernels compute (identity checked to 1e-13 by the oracle tests).
//
// Observations are whitened with W = chol(R)^{-1} (y' = W y, H' = W H, unit noise) and then absorbed
// ONE SCALAR AT A TIME, so no m x m system is ever solved and m is a run-time quantity.
//
// Two carried objects:
//   FilterState  (mean, cov)                      -- the sequential filter
//   ScanElem     (A, b, C, eta, J)                -- the associative element of the parallel-in-time
//       scan (Sarkka & Garcia-Fernandez 2021): over a range of steps i..j
//         x_j | x_{i-1}, y_{i:j} ~ N(A x_{i-1} + b, C),   p(y_{i:j} | x_{i-1}) ∝ N_info(x_{i-1}; eta, J)
//       plus the scalar ell with  p(y_{i:j} | x_{i-1}) = exp(ell - x^T J x / 2 + eta^T x),  so that the
//       ell of the join of ALL elements of a series (whose first element is the prior) IS the
//       marginal log-likelihood: one pass over the data and an ordered reduction, no second sweep.
//       Extending a range by one step is a transition + scalar absorptions (cheap, no inverse);
//       joining two ranges is elem_combine (one Cholesky of I + L^T J L).
#pragma once
#include "ssm_kernels.cuh"

namespace mf {

constexpr int kMaxObsDim = 4;  // run-time output_dim m <= kMaxObsDim

template <typename T, int D>
struct FilterState {
  T m[D];
  T P[D * D];  // full symmetric
};

template <typename T, int D>
struct ScanElem {
  // values per element in memory: A, b, C, eta, J, ell
  static constexpr int N = 3 * D * D + 2 * D + 1;
  T A[D * D], b[D], C[D * D], eta[D], J[D * D];
  T ell;  // log-normaliser: p(y_{i:j} | x_{i-1}) = exp(ell - x^T J x / 2 + eta^T x)
};

// ---- symmetric helpers --------------------------------------------------------------------------

// P <- F P F^T + Lq Lq^T   (P full symmetric in/out)
template <typename T, int D>
__device__ __forceinline__ void cov_predict(T* __restrict__ P, const T* __restrict__ F,
                                            const T* __restrict__ Lq) {
  constexpr int DD = D * D;
  T FP[DD], Q[DD];
  gemm<T, D>(FP, F, P);
  llt<T, D>(Q, Lq);
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      T v = Q[i * D + j];
#pragma unroll
      for (int q = 0; q < D; ++q) v = Num<T>::fma(FP[i * D + q], F[j * D + q], v);
      P[i * D + j] = v;
      P[j * D + i] = v;
    }
}

// v <- F v + u
template <typename T, int D>
__device__ __forceinline__ void mean_predict(T* __restrict__ v, const T* __restrict__ F,
                                             const T* __restrict__ u) {
  T t[D];
#pragma unroll
  for (int i = 0; i < D; ++i) {
    T a = u[i];
#pragma unroll
    for (int q = 0; q < D; ++q) a = Num<T>::fma(F[i * D + q], v[q], a);
    t[i] = a;
  }
#pragma unroll
  for (int i = 0; i < D; ++i) v[i] = t[i];
}

// ---- sequential filter --------------------------------------------------------------------------

template <typename T, int D>
__device__ __forceinline__ void filter_init(FilterState<T, D>& st, const T* __restrict__ mu0,
                                            const T* __restrict__ chol_p0) {
#pragma unroll
  for (int i = 0; i < D; ++i) st.m[i] = mu0[i];
  llt<T, D>(st.P, chol_p0);
}

template <typename T, int D>
__device__ __forceinline__ void filter_predict(FilterState<T, D>& st, const T* __restrict__ F,
                                               const T* __restrict__ u, const T* __restrict__ Lq) {
  mean_predict<T, D>(st.m, F, u);
  cov_predict<T, D>(st.P, F, Lq);
}

// Absorb one whitened scalar observation y = h.x + N(0,1).  quad += v^2/s, det *= s.
template <typename T, int D>
__device__ __forceinline__ void filter_absorb(FilterState<T, D>& st, const T* __restrict__ h, T y,
                                              T& quad, LogProd<T>& det) {
  T g[D];
  T s = T(1), v = y;
#pragma unroll
  for (int i = 0; i < D; ++i) {
    T a = T(0);
#pragma unroll
    for (int q = 0; q < D; ++q) a = Num<T>::fma(st.P[i * D + q], h[q], a);
    g[i] = a;
  }
#pragma unroll
  for (int i = 0; i < D; ++i) {
    s = Num<T>::fma(h[i], g[i], s);
    v = Num<T>::fma(-h[i], st.m[i], v);
  }
  const T rs = Num<T>::rcp(s);
  const T vs = v * rs;
  quad = Num<T>::fma(v, vs, quad);
  det.mul_lazy(s);  // caller peels (at least every few absorptions)
#pragma unroll
  for (int i = 0; i < D; ++i) {
    st.m[i] = Num<T>::fma(g[i], vs, st.m[i]);
    const T ki = g[i] * rs;
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      const T p = Num<T>::fma(-ki, g[j], st.P[i * D + j]);
      st.P[i * D + j] = p;
      st.P[j * D + i] = p;
    }
  }
}

// ---- scan elements ------------------------------------------------------------------------------

template <typename T, int D>
__device__ __forceinline__ void elem_identity(ScanElem<T, D>& e) {
  e.ell = T(0);
#pragma unroll
  for (int i = 0; i < D; ++i) {
    e.b[i] = T(0);
    e.eta[i] = T(0);
#pragma unroll
    for (int j = 0; j < D; ++j) {
      e.A[i * D + j] = (i == j) ? T(1) : T(0);
      e.C[i * D + j] = T(0);
      e.J[i * D + j] = T(0);
    }
  }
}

// The prior x_0 ~ N(mu0, P0) as an element (no dependence on anything earlier).
template <typename T, int D>
__device__ __forceinline__ void elem_prior(ScanElem<T, D>& e, const T* __restrict__ mu0,
                                           const T* __restrict__ chol_p0) {
  e.ell = T(0);
#pragma unroll
  for (int i = 0; i < D; ++i) {
    e.b[i] = mu0[i];
    e.eta[i] = T(0);
  }
#pragma unroll
  for (int i = 0; i < D * D; ++i) {
    e.A[i] = T(0);
    e.J[i] = T(0);
  }
  llt<T, D>(e.C, chol_p0);
}

template <typename T, int D>
__device__ __forceinline__ void elem_transition(ScanElem<T, D>& e, const T* __restrict__ F,
                                                const T* __restrict__ u, const T* __restrict__ Lq) {
  T FA[D * D];
  gemm<T, D>(FA, F, e.A);
#pragma unroll
  for (int i = 0; i < D * D; ++i) e.A[i] = FA[i];
  mean_predict<T, D>(e.b, F, u);
  cov_predict<T, D>(e.C, F, Lq);
}

// Absorb one whitened scalar observation into the range element; the element's ell gains
// -(v^2/s + log s + log 2 pi)/2, accumulated by the caller as quad += v^2/s, det *= s.
template <typename T, int D>
__device__ __forceinline__ void elem_absorb(ScanElem<T, D>& e, const T* __restrict__ h, T y,
                                            T& quad, LogProd<T>& det) {
  T g[D], w[D];
  T s = T(1), v = y;
#pragma unroll
  for (int i = 0; i < D; ++i) {
    T a = T(0), c = T(0);
#pragma unroll
    for (int q = 0; q < D; ++q) {
      a = Num<T>::fma(e.C[i * D + q], h[q], a);
      c = Num<T>::fma(e.A[q * D + i], h[q], c);
    }
    g[i] = a;  // C h
    w[i] = c;  // A^T h
  }
#pragma unroll
  for (int i = 0; i < D; ++i) {
    s = Num<T>::fma(h[i], g[i], s);
    v = Num<T>::fma(-h[i], e.b[i], v);
  }
  const T rs = Num<T>::rcp(s);
  const T vs = v * rs;
  quad = Num<T>::fma(v, vs, quad);
  det.mul_lazy(s);  // caller peels (at least every few absorptions)
#pragma unroll
  for (int i = 0; i < D; ++i) {
    const T ki = g[i] * rs;
    const T wi = w[i] * rs;
    e.b[i] = Num<T>::fma(g[i], vs, e.b[i]);
    e.eta[i] = Num<T>::fma(w[i], vs, e.eta[i]);
#pragma unroll
    for (int j = 0; j < D; ++j) e.A[i * D + j] = Num<T>::fma(-ki, w[j], e.A[i * D + j]);
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      const T c = Num<T>::fma(-ki, g[j], e.C[i * D + j]);
      e.C[i * D + j] = c;
      e.C[j * D + i] = c;
      const T jj = Num<T>::fma(wi, w[j], e.J[i * D + j]);
      e.J[i * D + j] = jj;
      e.J[j * D + i] = jj;
    }
  }
}

// Cholesky of a positive SEMI-definite matrix: a non-positive (or tiny) pivot zeroes its column.
template <typename T, int D>
__device__ __forceinline__ void chol_psd(T* __restrict__ l, const T* __restrict__ c) {
  T tr = T(0);
#pragma unroll
  for (int i = 0; i < D; ++i) tr += Num<T>::abs(c[i * D + i]);
  const T tiny = tr * (sizeof(T) == 8 ? T(1e-28) : T(1e-12));
#pragma unroll
  for (int j = 0; j < D; ++j) {
    T p = c[j * D + j];
#pragma unroll
    for (int q = 0; q < j; ++q) p = Num<T>::fma(-l[j * D + q], l[j * D + q], p);
    const bool ok = p > tiny;
    const T r = ok ? Num<T>::rsqrt(p) : T(0);
    l[j * D + j] = ok ? p * r : T(0);
#pragma unroll
    for (int i = j + 1; i < D; ++i) {
      T v = c[i * D + j];
#pragma unroll
      for (int q = 0; q < j; ++q) v = Num<T>::fma(-l[i * D + q], l[j * D + q], v);
      l[i * D + j] = v * r;
    }
#pragma unroll
    for (int i = 0; i < j; ++i) l[i * D + j] = T(0);
  }
}

// out = ei (earlier range) joined with ej (later range).  out may alias neither input.
//   M = (I + Ci Jj)^{-1} = I - L (I + L^T Jj L)^{-1} L^T Jj,   Ci = L L^T  (I + L^T Jj L is SPD >= I)
template <typename T, int D>
__device__ __forceinline__ void elem_combine_inl(ScanElem<T, D>& out, const ScanElem<T, D>& ei,
                                                 const ScanElem<T, D>& ej) {
  constexpr int DD = D * D;
  T L[DD], G[DD], R[DD], rinv[D], tmp[DD];
  chol_psd<T, D>(L, ei.C);
  gemm<T, D>(tmp, ej.J, L);   // Jj L
  gemm_tn<T, D>(G, L, tmp);   // L^T Jj L
#pragma unroll
  for (int i = 0; i < D; ++i) G[i * D + i] += T(1);
#pragma unroll
  for (int i = 0; i < DD; ++i) R[i] = G[i];
  chol_lower<T, D>(R, rinv);
  // K = L (I+G)^{-1} L^T  (symmetric PSD);  M X = X - K (Jj X),  M^T Y = Y - Jj (K Y)
  T Kk[DD];
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) tmp[i * D + j] = L[j * D + i];  // L^T
  trsm_left_lower<T, D>(R, rinv, tmp);                          // R^{-1} L^T
  gemm_tn<T, D>(Kk, tmp, tmp);                                  // L R^{-T} R^{-1} L^T
  // X-type products
  T JA[DD], MA[DD], MC[DD], Mb[D], t1[D], t2[D];
  gemm<T, D>(JA, ej.J, ei.A);  // Jj Ai
  gemm<T, D>(tmp, Kk, JA);
#pragma unroll
  for (int i = 0; i < DD; ++i) MA[i] = ei.A[i] - tmp[i];  // M Ai
  gemm<T, D>(tmp, ej.J, ei.C);
  gemm<T, D>(MC, Kk, tmp);
#pragma unroll
  for (int i = 0; i < DD; ++i) MC[i] = ei.C[i] - MC[i];  // M Ci
  // b: Aj M (bi + Ci eta_j) + bj
#pragma unroll
  for (int i = 0; i < D; ++i) t1[i] = ei.b[i];
  gemv_add<T, D>(t1, ei.C, ej.eta);
#pragma unroll
  for (int i = 0; i < D; ++i) t2[i] = T(0);
  gemv_add<T, D>(t2, ej.J, t1);
#pragma unroll
  for (int i = 0; i < D; ++i) Mb[i] = t1[i];
  gemv_sub<T, D>(Mb, Kk, t2);
#pragma unroll
  for (int i = 0; i < D; ++i) out.b[i] = ej.b[i];
  gemv_add<T, D>(out.b, ej.A, Mb);
  // A, C
  gemm<T, D>(out.A, ej.A, MA);
  gemm<T, D>(tmp, ej.A, MC);
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      T v = ej.C[i * D + j];
#pragma unroll
      for (int q = 0; q < D; ++q) v = Num<T>::fma(tmp[i * D + q], ej.A[j * D + q], v);
      out.C[i * D + j] = v;
      out.C[j * D + i] = v;
    }
  // eta: Ai^T M^T (eta_j - Jj bi) + eta_i
#pragma unroll
  for (int i = 0; i < D; ++i) t1[i] = ej.eta[i];
  gemv_sub<T, D>(t1, ej.J, ei.b);
  {
    // ell = ell_i + ell_j + log Int exp(-x^T Jj x/2 + eta_j^T x) N(x; b_i, C_i) dx
    //     = ... + b_i.(eta_j + t1)/2 - sum log R_kk + |R^{-1} L^T t1|^2 / 2
    T cz[D];
#pragma unroll
    for (int i = 0; i < D; ++i) cz[i] = T(0);
    gemv_t_add<T, D>(cz, L, t1);
    trsv_lower<T, D>(R, rinv, cz);
    T f = T(0), rprod = T(1);
#pragma unroll
    for (int i = 0; i < D; ++i) {
      f = Num<T>::fma(ei.b[i], ej.eta[i] + t1[i], f);
      f = Num<T>::fma(cz[i], cz[i], f);
      rprod *= rinv[i];
    }
    out.ell = ei.ell + ej.ell + T(0.5) * f + Num<T>::log(rprod);
  }
#pragma unroll
  for (int i = 0; i < D; ++i) t2[i] = T(0);
  gemv_add<T, D>(t2, Kk, t1);
  gemv_sub<T, D>(t1, ej.J, t2);  // M^T (eta_j - Jj bi)
#pragma unroll
  for (int i = 0; i < D; ++i) out.eta[i] = ei.eta[i];
  gemv_t_add<T, D>(out.eta, ei.A, t1);
  // J: Ai^T M^T Jj Ai + Ji  = Ai^T (JA - Jj K JA) + Ji
  gemm<T, D>(tmp, Kk, JA);
  T NJA[DD];
  gemm<T, D>(NJA, ej.J, tmp);
#pragma unroll
  for (int i = 0; i < DD; ++i) NJA[i] = JA[i] - NJA[i];
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      T v = ei.J[i * D + j];
#pragma unroll
      for (int q = 0; q < D; ++q) v = Num<T>::fma(ei.A[q * D + i], NJA[q * D + j], v);
      out.J[i * D + j] = v;
      out.J[j * D + i] = v;
    }
}

// out-of-line flavour for call sites where code size / compile time matter more than latency
template <typename T, int D>
__device__ __noinline__ void elem_combine(ScanElem<T, D>& out, const ScanElem<T, D>& ei,
                                          const ScanElem<T, D>& ej) {
  elem_combine_inl<T, D>(out, ei, ej);
}

template <typename T, int D>
__device__ __forceinline__ void elem_store(T* __restrict__ p, const ScanElem<T, D>& e) {
  constexpr int DD = D * D;
#pragma unroll
  for (int i = 0; i < DD; ++i) { p[i] = e.A[i]; p[DD + D + i] = e.C[i]; p[2 * DD + 2 * D + i] = e.J[i]; }
#pragma unroll
  for (int i = 0; i < D; ++i) { p[DD + i] = e.b[i]; p[2 * DD + D + i] = e.eta[i]; }
  p[3 * DD + 2 * D] = e.ell;
}

template <typename T, int D>
__device__ __forceinline__ void elem_load(ScanElem<T, D>& e, const T* __restrict__ p) {
  constexpr int DD = D * D;
#pragma unroll
  for (int i = 0; i < DD; ++i) { e.A[i] = p[i]; e.C[i] = p[DD + D + i]; e.J[i] = p[2 * DD + 2 * D + i]; }
#pragma unroll
  for (int i = 0; i < D; ++i) { e.b[i] = p[DD + i]; e.eta[i] = p[2 * DD + D + i]; }
  e.ell = p[3 * DD + 2 * D];
}

// ---- observation access: whitening with W = chol(R)^{-1} -------------------------------------------

// Whitener of one m x m lower Cholesky factor (m <= kMaxObsDim), and sum_i log W_ii.
template <typename T>
struct Whitener {
  T W[kMaxObsDim * kMaxObsDim];
  __device__ __forceinline__ void set(const T* __restrict__ chol_r, int m) {
    for (int c = 0; c < m; ++c)
      for (int i = 0; i < m; ++i) {
        if (i < c) { W[i * kMaxObsDim + c] = T(0); continue; }
        T v = (i == c) ? T(1) : T(0);
        for (int q = c; q < i; ++q) v -= chol_r[i * m + q] * W[q * kMaxObsDim + c];
        W[i * kMaxObsDim + c] = v / chol_r[i * m + i];
      }
  }
};

}  // namespace mf
