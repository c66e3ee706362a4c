// Register-resident D x D block arithmetic for the thread-per-chain kernels (D <= 8).
//
// Every function is fully unrolled on compile-time D so that blocks live in registers; a chain's
// recurrence state never touches memory.  Matrices are row-major T[D*D].  "Lower" functions read
// and write only i >= j entries.
#pragma once
#include <cuda_runtime.h>

namespace mf {

template <typename T> struct Num;
template <> struct Num<double> {
  // Branch-free 1/sqrt(x) for positive normal x: MUFU.RSQ64H seed (rel. error <= 2^-20 with the low word it
  // ignores) + one third-order step y(1 + e/2 + 3e^2/8), e = 1 - x y^2, error O(e^3) < 2^-60 before rounding.
  // ::rsqrt() wraps the same seed in a slow-path BRANCH (denormals, 0, inf) that ends the basic block, so
  // independent rsqrt's of one step could not overlap (ncu: 3 x ~70 cycles serialised in the D=3 Cholesky
  // step).  x <= 0 or NaN gives NaN/inf here as well; callers test the pivot's sign themselves.
  static __device__ __forceinline__ double rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double h = x * y;
    const double e = ::fma(-h, y, 1.0);
    const double t = ::fma(0.375, e, 0.5);
    // x = 0, inf (seed inf, 0: h is NaN) and NaN seeds fall through as the seed itself -- a select, not a branch
    return ::fabs(e) < 0.5 ? ::fma(y * e, t, y) : y;
  }
  // the library routine, for pivots that depend on each other anyway (large-block kernels: one rsqrt per
  // column, nothing to overlap it with): its special-case test is an integer compare off the FP64 chain
  static __device__ __forceinline__ double rsqrt_seq(double x) { return ::rsqrt(x); }
  static __device__ __forceinline__ double log(double x) { return ::log(x); }
  static __device__ __forceinline__ double fma(double a, double b, double c) { return ::fma(a, b, c); }
  static __device__ __forceinline__ double abs(double x) { return ::fabs(x); }
  // Branch-free 1/x for normal x (no division subroutine): MUFU.RCP64H seed + y(1 + e + e^2), e = 1 - x y.
  static __device__ __forceinline__ double rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = ::fma(-x, y, 1.0);
    const double t = ::fma(e, e, e);
    return ::fabs(e) < 0.5 ? ::fma(y, t, y) : y;  // 1/inf = 0, 1/0 = inf (an "infinite" noise variance = missing observation)
  }
};
template <> struct Num<float> {
  static __device__ __forceinline__ float rsqrt(float x) { return ::rsqrtf(x); }
  static __device__ __forceinline__ float rsqrt_seq(float x) { return ::rsqrtf(x); }
  static __device__ __forceinline__ float log(float x) { return ::logf(x); }
  static __device__ __forceinline__ float fma(float a, float b, float c) { return ::fmaf(a, b, c); }
  static __device__ __forceinline__ float abs(float x) { return ::fabsf(x); }
  static __device__ __forceinline__ float rcp(float x) { return 1.0f / x; }
};

// In-place Cholesky of the lower triangle of s.  rinv[j] = 1 / L[j][j].
// Returns false if a pivot is not strictly positive (NaN counts as failure); the factor then holds
// NaNs from that column on, as a LAPACK-style failure would leave it undefined.
template <typename T, int D>
__device__ __forceinline__ bool chol_lower(T* __restrict__ s, T* __restrict__ rinv) {
  bool ok = true;
#pragma unroll
  for (int j = 0; j < D; ++j) {
    T p = s[j * D + j];
#pragma unroll
    for (int q = 0; q < j; ++q) p = Num<T>::fma(-s[j * D + q], s[j * D + q], p);
    ok = ok && (p > T(0));
    const T r = Num<T>::rsqrt(p);
    rinv[j] = r;
    s[j * D + j] = p * r;
#pragma unroll
    for (int i = j + 1; i < D; ++i) {
      T v = s[i * D + j];
#pragma unroll
      for (int q = 0; q < j; ++q) v = Num<T>::fma(-s[i * D + q], s[j * D + q], v);
      s[i * D + j] = v * r;
    }
  }
  return ok;
}

// x <- L^{-1} x   (forward substitution, L lower with reciprocal diagonal rinv)
template <typename T, int D>
__device__ __forceinline__ void trsv_lower(const T* __restrict__ l, const T* __restrict__ rinv,
                                           T* __restrict__ x) {
#pragma unroll
  for (int i = 0; i < D; ++i) {
    T v = x[i];
#pragma unroll
    for (int q = 0; q < i; ++q) v = Num<T>::fma(-l[i * D + q], x[q], v);
    x[i] = v * rinv[i];
  }
}

// x <- L^{-T} x   (back substitution with the transpose of a lower factor)
template <typename T, int D>
__device__ __forceinline__ void trsv_lower_t(const T* __restrict__ l, const T* __restrict__ rinv,
                                             T* __restrict__ x) {
#pragma unroll
  for (int i = D - 1; i >= 0; --i) {
    T v = x[i];
#pragma unroll
    for (int q = i + 1; q < D; ++q) v = Num<T>::fma(-l[q * D + i], x[q], v);
    x[i] = v * rinv[i];
  }
}

// a <- a L^{-T}   (row-wise: each row r solves  x L^T = a_r, i.e. L x^T = a_r^T)
template <typename T, int D>
__device__ __forceinline__ void trsm_right_lower_t(T* __restrict__ a, const T* __restrict__ l,
                                                   const T* __restrict__ rinv) {
#pragma unroll
  for (int r = 0; r < D; ++r) trsv_lower<T, D>(l, rinv, a + r * D);
}

// a <- a L^{-1}   (row-wise: x L = a_r, i.e. L^T x^T = a_r^T)
template <typename T, int D>
__device__ __forceinline__ void trsm_right_lower(T* __restrict__ a, const T* __restrict__ l,
                                                 const T* __restrict__ rinv) {
#pragma unroll
  for (int r = 0; r < D; ++r) trsv_lower_t<T, D>(l, rinv, a + r * D);
}

// a <- L^{-1} a   (column-wise forward substitution on a full D x D right-hand side)
template <typename T, int D>
__device__ __forceinline__ void trsm_left_lower(const T* __restrict__ l, const T* __restrict__ rinv,
                                                T* __restrict__ a) {
#pragma unroll
  for (int i = 0; i < D; ++i) {
#pragma unroll
    for (int c = 0; c < D; ++c) {
      T v = a[i * D + c];
#pragma unroll
      for (int q = 0; q < i; ++q) v = Num<T>::fma(-l[i * D + q], a[q * D + c], v);
      a[i * D + c] = v * rinv[i];
    }
  }
}

// a <- L^{-T} a
template <typename T, int D>
__device__ __forceinline__ void trsm_left_lower_t(const T* __restrict__ l, const T* __restrict__ rinv,
                                                  T* __restrict__ a) {
#pragma unroll
  for (int i = D - 1; i >= 0; --i) {
#pragma unroll
    for (int c = 0; c < D; ++c) {
      T v = a[i * D + c];
#pragma unroll
      for (int q = i + 1; q < D; ++q) v = Num<T>::fma(-l[q * D + i], a[q * D + c], v);
      a[i * D + c] = v * rinv[i];
    }
  }
}

// s(lower) <- s - x x^T
template <typename T, int D>
__device__ __forceinline__ void syrk_sub_lower(T* __restrict__ s, const T* __restrict__ x) {
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      T v = s[i * D + j];
#pragma unroll
      for (int q = 0; q < D; ++q) v = Num<T>::fma(-x[i * D + q], x[j * D + q], v);
      s[i * D + j] = v;
    }
}

// y <- y - A x
template <typename T, int D>
__device__ __forceinline__ void gemv_sub(T* __restrict__ y, const T* __restrict__ a,
                                         const T* __restrict__ x) {
#pragma unroll
  for (int i = 0; i < D; ++i) {
    T v = y[i];
#pragma unroll
    for (int q = 0; q < D; ++q) v = Num<T>::fma(-a[i * D + q], x[q], v);
    y[i] = v;
  }
}

// y <- y - A^T x
template <typename T, int D>
__device__ __forceinline__ void gemv_t_sub(T* __restrict__ y, const T* __restrict__ a,
                                           const T* __restrict__ x) {
#pragma unroll
  for (int i = 0; i < D; ++i) {
    T v = y[i];
#pragma unroll
    for (int q = 0; q < D; ++q) v = Num<T>::fma(-a[q * D + i], x[q], v);
    y[i] = v;
  }
}

// y <- y + A x
template <typename T, int D>
__device__ __forceinline__ void gemv_add(T* __restrict__ y, const T* __restrict__ a,
                                         const T* __restrict__ x) {
#pragma unroll
  for (int i = 0; i < D; ++i) {
    T v = y[i];
#pragma unroll
    for (int q = 0; q < D; ++q) v = Num<T>::fma(a[i * D + q], x[q], v);
    y[i] = v;
  }
}

// y <- y + A^T x
template <typename T, int D>
__device__ __forceinline__ void gemv_t_add(T* __restrict__ y, const T* __restrict__ a,
                                           const T* __restrict__ x) {
#pragma unroll
  for (int i = 0; i < D; ++i) {
    T v = y[i];
#pragma unroll
    for (int q = 0; q < D; ++q) v = Num<T>::fma(a[q * D + i], x[q], v);
    y[i] = v;
  }
}

// c <- a b      (full D x D)
template <typename T, int D>
__device__ __forceinline__ void gemm(T* __restrict__ c, const T* __restrict__ a,
                                     const T* __restrict__ b) {
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) {
      T v = T(0);
#pragma unroll
      for (int q = 0; q < D; ++q) v = Num<T>::fma(a[i * D + q], b[q * D + j], v);
      c[i * D + j] = v;
    }
}

// c <- a^T b
template <typename T, int D>
__device__ __forceinline__ void gemm_tn(T* __restrict__ c, const T* __restrict__ a,
                                        const T* __restrict__ b) {
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) {
      T v = T(0);
#pragma unroll
      for (int q = 0; q < D; ++q) v = Num<T>::fma(a[q * D + i], b[q * D + j], v);
      c[i * D + j] = v;
    }
}

// c <- a b^T
template <typename T, int D>
__device__ __forceinline__ void gemm_nt(T* __restrict__ c, const T* __restrict__ a,
                                        const T* __restrict__ b) {
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) {
      T v = T(0);
#pragma unroll
      for (int q = 0; q < D; ++q) v = Num<T>::fma(a[i * D + q], b[j * D + q], v);
      c[i * D + j] = v;
    }
}

// Full symmetric  (L L^T)^{-1}  from a lower factor: inverse = L^{-T} L^{-1}.
// w is D*D scratch; out receives the full symmetric matrix.
template <typename T, int D>
__device__ __forceinline__ void chol_inverse(T* __restrict__ out, const T* __restrict__ l,
                                             const T* __restrict__ rinv) {
  T w[D * D];  // w = L^{-1}  (lower)
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int c = 0; c < D; ++c) {
      if (c > i) { w[i * D + c] = T(0); continue; }
      T v = (i == c) ? T(1) : T(0);
#pragma unroll
      for (int q = c; q < i; ++q) v = Num<T>::fma(-l[i * D + q], w[q * D + c], v);
      w[i * D + c] = v * rinv[i];
    }
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      T v = T(0);
#pragma unroll
      for (int q = i; q < D; ++q) v = Num<T>::fma(w[q * D + i], w[q * D + j], v);
      out[i * D + j] = v;
      out[j * D + i] = v;
    }
}

// s(lower) <- s - t w^T     (t and w full D x D)
template <typename T, int D>
__device__ __forceinline__ void gemm_nt_sub_lower(T* __restrict__ s, const T* __restrict__ t,
                                                  const T* __restrict__ w) {
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      T v = s[i * D + j];
#pragma unroll
      for (int q = 0; q < D; ++q) v = Num<T>::fma(-t[i * D + q], w[j * D + q], v);
      s[i * D + j] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// Short-critical-path factorisation step for D <= 3.
//
// The textbook Cholesky of a D x D block chains D reciprocal-square-roots (each ~66 cycles of
// dependent FP64 latency on B200).  Here the pivots d_j of S = L_u diag(d) L_u^T are obtained from
// the leading principal minors m_j (d_j = m_j / m_{j-1}), so the D rsqrt's q_j = rsqrt(m_j) are
// INDEPENDENT of each other:
//     sqrt(d_j) = (m_j q_j) q_{j-1},    1/sqrt(d_j) = q_j (m_{j-1} q_{j-1}),    prod_j d_j = m_D
// Outputs are the ordinary Cholesky quantities (identical in exact arithmetic; accuracy against a
// long-double factorisation is the same as the chained form):
//   lu  unit-lower factor L_u (strict lower part), rs[j] = 1/sqrt(d_j), sq[j] = sqrt(d_j),
//   det = m_D.  Returns false on a non-positive pivot.
// ---------------------------------------------------------------------------------------------
template <typename T, int D>
__device__ __forceinline__ bool ldl_minors(const T* __restrict__ s, T* __restrict__ lu,
                                           T* __restrict__ rs, T* __restrict__ sq, T& det) {
  static_assert(D >= 1 && D <= 3, "minor-based pivots are implemented for D <= 3");
  const T m1 = s[0];
  const T q0 = Num<T>::rsqrt(m1);
  const T sm1 = m1 * q0;  // sqrt(m1)
  rs[0] = q0;
  sq[0] = sm1;
  det = m1;
  bool ok = m1 > T(0);
  if (D >= 2) {
    const T s10 = s[1 * D + 0], s11 = s[1 * D + 1];
    const T m2 = Num<T>::fma(m1, s11, -(s10 * s10));
    const T q1 = Num<T>::rsqrt(m2);
    const T sm2 = m2 * q1;  // sqrt(m2)
    ok = ok && (m2 > T(0));
    const T rd0 = q0 * q0;  // 1/d_0
    rs[1] = q1 * sm1;
    sq[1] = sm2 * q0;
    det = m2;
    lu[1 * D + 0] = s10 * rd0;
    if (D >= 3) {
      const T s20 = s[2 * D + 0], s21 = s[2 * D + 1], s22 = s[2 * D + 2];
      const T c0 = Num<T>::fma(s11, s22, -(s21 * s21));
      const T c1 = Num<T>::fma(s10, s22, -(s21 * s20));
      const T c2 = Num<T>::fma(s10, s21, -(s11 * s20));
      const T m3 = Num<T>::fma(s20, c2, Num<T>::fma(-s10, c1, m1 * c0));
      const T q2 = Num<T>::rsqrt(m3);
      ok = ok && (m3 > T(0));
      rs[2] = q2 * sm2;
      sq[2] = (m3 * q2) * q1;
      det = m3;
      const T rd1 = rs[1] * rs[1];  // 1/d_1
      lu[2 * D + 0] = s20 * rd0;
      lu[2 * D + 1] = Num<T>::fma(-lu[2 * D + 0], s10, s21) * rd1;
    }
  }
  return ok;
}

template <typename T, int D>
__device__ __forceinline__ void zero_upper(T* __restrict__ a) {
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = i + 1; j < D; ++j) a[i * D + j] = T(0);
}

template <typename T, int D>
__device__ __forceinline__ void mirror_lower(T* __restrict__ a) {
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = i + 1; j < D; ++j) a[i * D + j] = a[j * D + i];
}

template <typename T, int N>
__device__ __forceinline__ void load_vec(T* __restrict__ r, const T* __restrict__ g) {
#pragma unroll
  for (int i = 0; i < N; ++i) r[i] = __ldg(g + i);
}

// plain (coherent, ordered) loads: for memory that the same kernel also writes -- the parked
// elements / seeds of the parallel-in-time passes live in OUTPUT arrays.  load_vec's __ldg
// (ld.global.nc) assumes read-only data and may be reordered past a later store to the same address.
template <typename T, int N>
__device__ __forceinline__ void load_vec_rw(T* r, const T* g) {
#pragma unroll
  for (int i = 0; i < N; ++i) r[i] = g[i];
}

template <typename T, int N>
__device__ __forceinline__ void store_vec(T* __restrict__ g, const T* __restrict__ r) {
#pragma unroll
  for (int i = 0; i < N; ++i) g[i] = r[i];
}

}  // namespace mf
