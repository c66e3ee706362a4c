// Warp-cooperative dense algebra on d x d blocks (runtime d <= 32) held in shared memory: the arithmetic layer of
// the mid-size kernels (8 < D <= 32, mid_kernels.cuh), where a block no longer fits one thread's registers.
//
// A matrix is stored row-major with leading dimension MID_LD = 33 elements (odd: lanes walking a column hit
// different banks).  One WARP owns the matrices it works on; every primitive is called by all 32 lanes and
// starts and ends with __syncwarp(): it may overwrite what earlier primitives read and the next one may read what
// it wrote.  Work split: lane = row
// (products, factorisation) or lane = column (triangular solves of matrices: the columns are independent, so
// a solve needs no synchronisation inside).  The operation ORDER inside every dot product is that of the
// one-thread-per-chain primitives of smallmat.cuh (ascending index, fused multiply-add), so both paths
// round alike.
#pragma once
#include "smallmat.cuh"

namespace mf {

constexpr int MID_LD = 33;
constexpr int MID_MAT = 32 * MID_LD;  // elements of one matrix slot

template <typename T>
__device__ __forceinline__ T& mid_at(T* m, int i, int j) { return m[i * MID_LD + j]; }

// global (row-major d x d, contiguous) -> shared
template <typename T>
__device__ __forceinline__ void mid_load(T* __restrict__ m, const T* __restrict__ g, int d, int lane) {
  __syncwarp();  // earlier reads of the slot (e.g. by a vector solve, which ends on a shuffle) are done
  for (int idx = lane; idx < d * d; idx += 32) m[(idx / d) * MID_LD + (idx % d)] = g[idx];
  __syncwarp();
}
template <typename T>
__device__ __forceinline__ void mid_store(T* __restrict__ g, const T* __restrict__ m, int d, int lane) {
  __syncwarp();
  for (int idx = lane; idx < d * d; idx += 32) g[idx] = m[(idx / d) * MID_LD + (idx % d)];
}
// the same with a row pitch on the global side (a d x d window of a wider row-major matrix)
template <typename T>
__device__ __forceinline__ void mid_load_pitched(T* __restrict__ m, const T* __restrict__ g, int d, int pitch,
                                                 int lane) {
  __syncwarp();
  for (int idx = lane; idx < d * d; idx += 32) m[(idx / d) * MID_LD + (idx % d)] = g[(idx / d) * pitch + (idx % d)];
  __syncwarp();
}
template <typename T>
__device__ __forceinline__ void mid_store_pitched(T* __restrict__ g, const T* __restrict__ m, int d, int pitch,
                                                  int lane) {
  __syncwarp();
  for (int idx = lane; idx < d * d; idx += 32) g[(idx / d) * pitch + (idx % d)] = m[(idx / d) * MID_LD + (idx % d)];
}
// a (op)= b, elementwise: SIGN +1 / -1;  a = b - a for SIGN 0
template <typename T, int SIGN>
__device__ __forceinline__ void mid_axpy(T* __restrict__ a, const T* __restrict__ b, int d, int lane) {
  __syncwarp();
  if (lane < d)
    for (int j = 0; j < d; ++j) {
      const T x = a[lane * MID_LD + j], y = b[lane * MID_LD + j];
      a[lane * MID_LD + j] = SIGN > 0 ? x + y : (SIGN < 0 ? x - y : y - x);
    }
  __syncwarp();
}
// lower triangle only, zeros above (Cholesky factors as the reference returns them)
template <typename T>
__device__ __forceinline__ void mid_store_lower(T* __restrict__ g, const T* __restrict__ m, int d, int lane) {
  __syncwarp();
  for (int idx = lane; idx < d * d; idx += 32) {
    const int i = idx / d, j = idx % d;
    g[idx] = j <= i ? m[i * MID_LD + j] : T(0);
  }
}
template <typename T>
__device__ __forceinline__ void mid_copy(T* __restrict__ dst, const T* __restrict__ src, int d, int lane) {
  __syncwarp();
  if (lane < d)
    for (int j = 0; j < d; ++j) dst[lane * MID_LD + j] = src[lane * MID_LD + j];
  __syncwarp();
}
template <typename T>
__device__ __forceinline__ void mid_scale(T* m, T alpha, int d, int lane) {
  __syncwarp();
  if (lane < d)
    for (int j = 0; j < d; ++j) m[lane * MID_LD + j] *= alpha;
  __syncwarp();
}
template <typename T>
__device__ __forceinline__ void mid_transpose(T* __restrict__ dst, const T* __restrict__ src, int d, int lane) {
  __syncwarp();
  if (lane < d)
    for (int j = 0; j < d; ++j) dst[lane * MID_LD + j] = src[j * MID_LD + lane];
  __syncwarp();
}

// C (op)= A' B'  with A' = A or A^T, B' = B or B^T;  MODE 0: C = , 1: C += , -1: C -=.  Lane = row of C.
template <typename T, bool TA, bool TB, int MODE>
__device__ __forceinline__ void mid_gemm(T* __restrict__ c, const T* __restrict__ a, const T* __restrict__ b, int d,
                                         int lane) {
  __syncwarp();
  if (lane < d) {
    for (int j = 0; j < d; ++j) {
      T v = MODE == 0 ? T(0) : c[lane * MID_LD + j];
      for (int q = 0; q < d; ++q) {
        const T x = TA ? a[q * MID_LD + lane] : a[lane * MID_LD + q];
        const T y = TB ? b[j * MID_LD + q] : b[q * MID_LD + j];
        v = Num<T>::fma(MODE < 0 ? -x : x, y, v);
      }
      c[lane * MID_LD + j] = v;
    }
  }
  __syncwarp();
}

// y (op)= A' x for vectors held one element per lane (registers): returns the lane's entry.
// MODE 0: A'x, 1: y0 + A'x, -1: y0 - A'x.
template <typename T, bool TA, int MODE>
__device__ __forceinline__ T mid_gemv(const T* __restrict__ a, T x, T y0, int d, int lane) {
  __syncwarp();
  T v = MODE == 0 ? T(0) : y0;
  for (int q = 0; q < d; ++q) {
    const T xq = __shfl_sync(0xffffffffu, x, q);
    if (lane < d) {
      const T aq = TA ? a[q * MID_LD + lane] : a[lane * MID_LD + q];
      v = Num<T>::fma(MODE < 0 ? -aq : aq, xq, v);
    }
  }
  return v;
}

// In-place Cholesky of the lower triangle (left-looking by columns, as chol_lower): rinv[j] = 1 / L[j][j].
// Returns false (warp-uniform) if a pivot is not strictly positive.
template <typename T>
__device__ __forceinline__ bool mid_chol(T* __restrict__ s, T* __restrict__ rinv, int d, int lane) {
  __syncwarp();
  bool ok = true;
  for (int j = 0; j < d; ++j) {
    // column j: every lane i >= j forms s[i][j] - sum_{q<j} L[i][q] L[j][q]
    T v = T(0);
    if (lane >= j && lane < d) {
      v = s[lane * MID_LD + j];
      for (int q = 0; q < j; ++q) v = Num<T>::fma(-s[lane * MID_LD + q], s[j * MID_LD + q], v);
    }
    const T p = __shfl_sync(0xffffffffu, v, j);
    ok = ok && (p > T(0));
    const T r = Num<T>::rsqrt_seq(p);
    if (lane == j) {
      s[j * MID_LD + j] = p * r;
      rinv[j] = r;
    } else if (lane > j && lane < d) {
      s[lane * MID_LD + j] = v * r;
    }
    __syncwarp();
  }
  return ok;
}
template <typename T>
__device__ __forceinline__ void mid_diag_rcp(const T* __restrict__ l, T* __restrict__ rinv, int d, int lane) {
  __syncwarp();
  if (lane < d) rinv[lane] = Num<T>::rcp(l[lane * MID_LD + lane]);
  __syncwarp();
}

// B <- L^{-1} B (matrix, d columns): lane = column, forward substitution down the column
template <typename T>
__device__ __forceinline__ void mid_trsm_l(const T* __restrict__ l, const T* __restrict__ rinv, T* __restrict__ b,
                                           int d, int lane) {
  __syncwarp();
  if (lane < d) {
    for (int i = 0; i < d; ++i) {
      T v = b[i * MID_LD + lane];
      for (int q = 0; q < i; ++q) v = Num<T>::fma(-l[i * MID_LD + q], b[q * MID_LD + lane], v);
      b[i * MID_LD + lane] = v * rinv[i];
    }
  }
  __syncwarp();
}
// B <- L^{-T} B: backward substitution up the column
template <typename T>
__device__ __forceinline__ void mid_trsm_lt(const T* __restrict__ l, const T* __restrict__ rinv, T* __restrict__ b,
                                            int d, int lane) {
  __syncwarp();
  if (lane < d) {
    for (int i = d - 1; i >= 0; --i) {
      T v = b[i * MID_LD + lane];
      for (int q = i + 1; q < d; ++q) v = Num<T>::fma(-l[q * MID_LD + i], b[q * MID_LD + lane], v);
      b[i * MID_LD + lane] = v * rinv[i];
    }
  }
  __syncwarp();
}
// B <- B L^{-1}: lane = row of B, backward over the columns
template <typename T>
__device__ __forceinline__ void mid_trsm_right_l(T* __restrict__ b, const T* __restrict__ l, const T* __restrict__ rinv,
                                                 int d, int lane) {
  __syncwarp();
  if (lane < d) {
    for (int j = d - 1; j >= 0; --j) {
      T v = b[lane * MID_LD + j];
      for (int q = j + 1; q < d; ++q) v = Num<T>::fma(-b[lane * MID_LD + q], l[q * MID_LD + j], v);
      b[lane * MID_LD + j] = v * rinv[j];
    }
  }
  __syncwarp();
}
// sum_q x_q y_q, accumulated in ascending order on top of `init` (identical on every lane)
template <typename T>
__device__ __forceinline__ T mid_dot(T x, T y, T init, int d) {
  T s = init;
  for (int q = 0; q < d; ++q) s = Num<T>::fma(__shfl_sync(0xffffffffu, x, q), __shfl_sync(0xffffffffu, y, q), s);
  return s;
}
// x <- L^{-1} x, x <- L^{-T} x for a vector held one element per lane
template <typename T>
__device__ __forceinline__ T mid_trsv_l(const T* __restrict__ l, const T* __restrict__ rinv, T x, int d, int lane) {
  __syncwarp();
  for (int j = 0; j < d; ++j) {
    const T xj = __shfl_sync(0xffffffffu, x, j) * rinv[j];
    if (lane == j) x = xj;
    else if (lane > j && lane < d) x = Num<T>::fma(-l[lane * MID_LD + j], xj, x);
  }
  return x;
}
template <typename T>
__device__ __forceinline__ T mid_trsv_lt(const T* __restrict__ l, const T* __restrict__ rinv, T x, int d, int lane) {
  __syncwarp();
  for (int j = d - 1; j >= 0; --j) {
    const T xj = __shfl_sync(0xffffffffu, x, j) * rinv[j];
    if (lane == j) x = xj;
    else if (lane < j) x = Num<T>::fma(-l[j * MID_LD + lane], xj, x);
  }
  return x;
}

// q = (L L^T)^{-1} (full symmetric), w = scratch matrix
template <typename T>
__device__ __forceinline__ void mid_chol_inverse(T* __restrict__ q, const T* __restrict__ l, const T* __restrict__ rinv,
                                                 T* __restrict__ w, int d, int lane) {
  __syncwarp();
  if (lane < d)
    for (int j = 0; j < d; ++j) w[lane * MID_LD + j] = lane == j ? T(1) : T(0);
  mid_trsm_l<T>(l, rinv, w, d, lane);            // w = L^{-1}
  mid_gemm<T, true, false, 0>(q, w, w, d, lane);  // q = L^{-T} L^{-1}
}

}  // namespace mf
