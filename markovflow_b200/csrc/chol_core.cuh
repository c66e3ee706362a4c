// The per-chain Cholesky(+solve) recurrence, independent of how step records reach the thread.
//
// A "record" for step k holds  D_k (D*D, lower triangle read), A_k (D*D, absent for the last
// step) and optionally b_k (D); the thread emits  Ld_k, Ls_k and x_k.  The layout policy says where
// element e of step s of each stream lives relative to the per-thread base pointers:
//     addr = base + s * L::SS_x + e * L::ES_x        (x in {M: matrix streams, V: vector streams})
//
// Carried between steps: Lp = Ls_{k-1} and xp = x_{k-1} (registers).
#pragma once
#include "smallmat.cuh"

namespace mf {

template <typename T, int D, bool RHS, class L>
struct CholCore {
  static constexpr int DD = D * D;
  static constexpr bool FAST = (D <= 3);
  struct Rec {  // one step's inputs, prefetched into registers
    T S[DD], A[DD], r[D];
  };
  T Lp[DD], xp[D];
  T prod;
  int esum;
  int32_t fail;

  __device__ __forceinline__ void init() {
#pragma unroll
    for (int i = 0; i < DD; ++i) Lp[i] = T(0);
#pragma unroll
    for (int i = 0; i < D; ++i) xp[i] = T(0);
    prod = T(1);
    esum = 0;
    fail = 0;
  }

  // sum_k sum_j log L_jj = 0.5 * log prod_k prod_j d_j
  __device__ __forceinline__ T log_det() const {
    return T(0.5) * (Num<T>::log(prod) + T(esum) * T(0.6931471805599453094));
  }

  static __device__ __forceinline__ void load_record(Rec& rec, const T* __restrict__ pd,
                                                     const T* __restrict__ ps,
                                                     const T* __restrict__ pr, int s) {
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int j = 0; j <= i; ++j) rec.S[i * D + j] = pd[s * L::SS_M + (i * D + j) * L::ES_M];
#pragma unroll
    for (int i = 0; i < DD; ++i) rec.A[i] = ps[s * L::SS_M + i * L::ES_M];
    if (RHS) {
#pragma unroll
      for (int i = 0; i < D; ++i) rec.r[i] = pr[s * L::SS_V + i * L::ES_V];
    }
  }

  // Factorise the step held in `rec` (destroyed); emit outputs for tile-local step s.
  __device__ __forceinline__ void step(Rec& rec, T* __restrict__ od, T* __restrict__ os,
                                       T* __restrict__ ox, int s, int64_t k, int64_t Tn,
                                       bool want_logdet) {
    T* od_s = od + s * L::SS_M;
    T* os_s = os + s * L::SS_M;
    T* ox_s = ox + s * L::SS_V;
    T* S = rec.S;
    T* A = rec.A;
    T* r = rec.r;
    syrk_sub_lower<T, D>(S, Lp);  // Schur update (Lp == 0 at k == 0)
    if (RHS) gemv_sub<T, D>(r, Lp, xp);
    T pstep;
    if constexpr (FAST) {
      T lu[DD], rs[D], sq[D];
      const bool ok = ldl_minors<T, D>(S, lu, rs, sq, pstep);
      if (!ok && fail == 0) fail = (int32_t)(k + 1);
      // Ld = L_u diag(sqrt d);  column 0 of L_u diag(sqrt d_0) is S[:,0] / sqrt(d_0)
#pragma unroll
      for (int j = 0; j < D; ++j)
#pragma unroll
        for (int i = 0; i < D; ++i) {
          T v = T(0);
          if (i == j) v = sq[j];
          if (i > j) v = (j == 0 ? S[i * D + 0] * rs[0] : lu[i * D + j] * sq[j]);
          od_s[(i * D + j) * L::ES_M] = v;
        }
      if (RHS) {
        // u = L_u^{-1} r ;  x = u / sqrt(d)
#pragma unroll
        for (int i = 1; i < D; ++i)
#pragma unroll
          for (int q = 0; q < i; ++q) r[i] = Num<T>::fma(-lu[i * D + q], r[q], r[i]);
#pragma unroll
        for (int i = 0; i < D; ++i) {
          xp[i] = r[i] * rs[i];
          ox_s[i * L::ES_V] = xp[i];
        }
      }
      if (k + 1 < Tn) {
        // W = A L_u^{-T} (row-wise unit solve), Ls = W diag(1/sqrt d)
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
          for (int j = 1; j < D; ++j)
#pragma unroll
            for (int q = 0; q < j; ++q)
              A[i * D + j] = Num<T>::fma(-A[i * D + q], lu[j * D + q], A[i * D + j]);
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
          for (int j = 0; j < D; ++j) {
            Lp[i * D + j] = A[i * D + j] * rs[j];
            os_s[(i * D + j) * L::ES_M] = Lp[i * D + j];
          }
      }
    } else {
      T rinv[D];
      const bool ok = chol_lower<T, D>(S, rinv);
      if (!ok && fail == 0) fail = (int32_t)(k + 1);
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) od_s[(i * D + j) * L::ES_M] = (j <= i) ? S[i * D + j] : T(0);
      pstep = S[0];
#pragma unroll
      for (int j = 1; j < D; ++j) pstep *= S[j * D + j];
      pstep *= pstep;  // product of pivots d_j = L_jj^2, as in the fast path
      if (RHS) {
        trsv_lower<T, D>(S, rinv, r);
#pragma unroll
        for (int i = 0; i < D; ++i) { xp[i] = r[i]; ox_s[i * L::ES_V] = r[i]; }
      }
      if (k + 1 < Tn) {
        trsm_right_lower_t<T, D>(A, S, rinv);
#pragma unroll
        for (int i = 0; i < DD; ++i) { Lp[i] = A[i]; os_s[i * L::ES_M] = A[i]; }
      }
    }
    if (want_logdet) {
      // running product of pivots with the exponent peeled off every step (no log in the loop)
      prod *= pstep;
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
  }

  // Process `ns` consecutive steps starting at global step k0.  Records are prefetched one step
  // ahead into two ping-pong register sets (no register-to-register copies in the loop).
  __device__ __forceinline__ void tile(const T* __restrict__ pd, const T* __restrict__ ps,
                                       const T* __restrict__ pr, T* __restrict__ od,
                                       T* __restrict__ os, T* __restrict__ ox, int ns, int64_t k0,
                                       int64_t Tn, bool want_logdet) {
    Rec ra, rb;
    load_record(ra, pd, ps, pr, 0);
    int s = 0;
    for (; s + 1 < ns; s += 2) {
      load_record(rb, pd, ps, pr, s + 1);
      step(ra, od, os, ox, s, k0 + s, Tn, want_logdet);
      if (s + 2 < ns) load_record(ra, pd, ps, pr, s + 2);
      step(rb, od, os, ox, s + 1, k0 + s + 1, Tn, want_logdet);
    }
    if (s < ns) step(ra, od, os, ox, s, k0 + s, Tn, want_logdet);
  }
};

}  // namespace mf
