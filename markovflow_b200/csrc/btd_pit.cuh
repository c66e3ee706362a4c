// Block-tridiagonal Cholesky (+ forward solve) of FEW LONG chains, parallel in time and exact.
//
// The sweep  S_k = D_k - A_{k-1} S_{k-1}^{-1} A_{k-1}^T,  r_k = b_k - A_{k-1} S_{k-1}^{-1} r_{k-1}
// (reference block_tri_diag.py:436,350; Ld_k = chol S_k, Ls_k = A_k Ld_k^{-T}, x_k = Ld_k^{-1} r_k)
// costs T x (latency of a step) however few chains there are.  The map (S_{k-1}, r_{k-1}) -> (S_k, r_k)
// is linear-fractional and closed under composition as
//     S_out = P - Q (S_in + R)^{-1} Q^T,      r_out = p + Q (S_in + R)^{-1} (r_in + r)
// (the mirrored form of the naturals -> SSM element in ssm_sweep.cuh).  Extending an element by step
// k is the ordinary step started without an incoming block -- L = chol P, Ls = A L^{-T}, x = L^{-1} p,
// P' = D_k - Ls Ls^T, p' = b_k - Ls x -- plus  W = L^{-1} Q,  Q' = -Ls W,  R' = R - W^T W,
// r' = r + W^T x.  Every chain is cut into P segments:
//   1. CholPitSummaryCore : every segment (but the last) reduces its steps to (P, Q, R, p, r)
//   2. chol_pit_seed_kernel: per chain, fold the elements in order; for every segment s >= 1 park the
//                            pair (Ls_{k0-1}, x_{k0-1}) that its first step needs
//   3. CholPitCore        : every segment runs the ordinary sweep from its seed.
// No workspace: elements and seeds are parked in the output slots of each segment's LAST steps
// (P | seed Ls -> od[last], R -> od[last-1], Q -> os[last], p | seed x -> ox[last], r -> ox[last-1]),
// so the outputs must not alias the inputs (the caller falls back to the sequential sweep if they do).
#pragma once
#include "ssm_sweep.cuh"

namespace mf {

template <typename T>
struct CholPitParams {
  const T *diag, *sub, *rhs;
  T *od, *os, *ox;
  int32_t* info;
  int64_t B, Tn;
  int64_t P, L;
};

// first failing step of a chain across its virtual chains: smallest non-zero value wins
__device__ __forceinline__ void atomic_min_nonzero(int32_t* a, int32_t v) {
  int32_t old = *reinterpret_cast<volatile int32_t*>(a);
  while (old == 0 || v < old) {
    const int32_t seen = atomicCAS(a, old, v);
    if (seen == old) break;
    old = seen;
  }
}

template <typename T_, int D, bool RHS>
struct CholPitGeom {
  using T = T_;
  using Params = CholPitParams<T>;
  static constexpr int DD = D * D;
  static constexpr bool BACKWARD = false;
  static constexpr int NIN = RHS ? 3 : 2;
  static constexpr int ein(int i) { return i < 2 ? DD : D; }
  static __device__ __forceinline__ int64_t num_chains(const Params& p) { return p.B * p.P; }
  static __device__ __forceinline__ int64_t max_steps(const Params& p) { return p.L; }
  static __device__ __forceinline__ StreamGeom in_geom_seg(const Params& p, int i, int64_t v, bool live) {
    const int64_t c = v / p.P, k0 = (v % p.P) * p.L;
    const int64_t n = live ? seg_steps(p.Tn, k0, p.L) : 0;
    if (i == 1) return vgeom_outgoing<T>(p.sub, c, p.Tn, DD, k0, n);
    return vgeom_states<T>(i == 0 ? p.diag : p.rhs, c, p.Tn, ein(i), k0, n);
  }
};

// ---- pass 3 (or the whole job): the ordinary sweep of one segment from its seed ----------------
template <typename T_, int D, bool RHS>
struct CholPitCore : CholPitGeom<T_, D, RHS> {
  using T = T_;
  using Params = CholPitParams<T>;
  static constexpr int DD = D * D;
  static constexpr int NOUT = RHS ? 3 : 2;
  static constexpr int eout(int i) { return i < 2 ? DD : D; }
  static __device__ __forceinline__ StreamGeom in_geom(const Params& p, int i, int64_t v) {
    return CholPitGeom<T_, D, RHS>::in_geom_seg(p, i, v, true);
  }
  static __device__ __forceinline__ StreamGeom out_geom(const Params& p, int i, int64_t v) {
    const int64_t c = v / p.P, k0 = (v % p.P) * p.L;
    const int64_t n = seg_steps(p.Tn, k0, p.L);
    if (i == 1) return vgeom_outgoing<T>(p.os, c, p.Tn, DD, k0, n);
    return vgeom_states<T>(i == 0 ? p.od : p.ox, c, p.Tn, eout(i), k0, n);
  }
  T Ls[DD], x[D];
  bool coupled_;
  int32_t fail;
  int64_t Tn_, k0_, n_;
  __device__ __forceinline__ void init(const Params& p, int64_t v) {
    const int64_t c = v / p.P;
    Tn_ = p.Tn;
    k0_ = (v % p.P) * p.L;
    n_ = seg_steps(p.Tn, k0_, p.L);
    fail = 0;
    coupled_ = false;
#pragma unroll
    for (int i = 0; i < D; ++i) x[i] = T(0);
    if (k0_ > 0 && n_ > 0) {  // seed (Ls_{k0-1}, x_{k0-1}) parked in this segment's last slots
      const int64_t kl = k0_ + n_ - 1;
      load_vec_rw<T, DD>(Ls, p.od + (c * p.Tn + kl) * DD);
      if (RHS) load_vec_rw<T, D>(x, p.ox + (c * p.Tn + kl) * D);
      coupled_ = true;
    }
  }
  __device__ __forceinline__ void tile(const Params&, const T* const* in, T* const* out, int64_t j0,
                                       int ns) {
    if (n_ - j0 < ns) ns = (int)(n_ - j0);
    for (int j = 0; j < ns; ++j) {
      const int64_t k = k0_ + j0 + j;
      T S[DD], rinv[D], r[D];
      ld_s<T, DD>(S, in[0] + j * DD);
      if (RHS) ld_s<T, D>(r, in[RHS ? 2 : 0] + j * D);
      if (coupled_) {
        syrk_sub_lower<T, D>(S, Ls);
        if (RHS) gemv_sub<T, D>(r, Ls, x);
      }
      const bool ok = chol_lower<T, D>(S, rinv);
      if (!ok && fail == 0) fail = (int32_t)(k + 1);
      zero_upper<T, D>(S);
      st_s<T, DD>(out[0] + j * DD, S);
      if (RHS) {
        trsv_lower<T, D>(S, rinv, r);
#pragma unroll
        for (int i = 0; i < D; ++i) x[i] = r[i];
        st_s<T, D>(out[RHS ? 2 : 0] + j * D, x);
      }
      if (k + 1 < Tn_) {
        ld_s<T, DD>(Ls, in[1] + j * DD);
        trsm_right_lower_t<T, D>(Ls, S, rinv);
        st_s<T, DD>(out[1] + j * DD, Ls);
        coupled_ = true;
      }
    }
  }
  __device__ __forceinline__ void finish(const Params& p, int64_t v, bool valid) {
    if (!valid || !p.info) return;
    if (p.P == 1) p.info[v] = fail;
    else if (fail) atomic_min_nonzero(p.info + v / p.P, fail);
  }
};

// ---- pass 1: element (P, Q, R, p, r) of every segment but the last of each chain ---------------
template <typename T_, int D, bool RHS>
struct CholPitSummaryCore : CholPitGeom<T_, D, RHS> {
  using T = T_;
  using Params = CholPitParams<T>;
  static constexpr int DD = D * D;
  static constexpr int NOUT = 0;
  static constexpr int eout(int) { return 1; }
  static __device__ __forceinline__ bool is_live(const Params& p, int64_t v) {
    const int64_t k0 = (v % p.P) * p.L;
    return k0 + p.L < p.Tn;  // a successor exists (and the segment is complete: L steps)
  }
  static __device__ __forceinline__ StreamGeom in_geom(const Params& p, int i, int64_t v) {
    return CholPitGeom<T_, D, RHS>::in_geom_seg(p, i, v, is_live(p, v));
  }
  static __device__ __forceinline__ StreamGeom out_geom(const Params&, int, int64_t) {
    return StreamGeom{nullptr, 0, 0};
  }
  T Ls[DD], x[D], Pm[DD], Q[DD], R[DD], pv[D], rv[D], Lf[DD], rinv[D];
  int32_t fail;
  int64_t Tn_, k0_;
  bool live_, started_;
  __device__ __forceinline__ void init(const Params& p, int64_t v) {
    const int64_t c = v / p.P;
    Tn_ = p.Tn;
    k0_ = (v % p.P) * p.L;
    live_ = is_live(p, v);
    started_ = false;
    fail = 0;
#pragma unroll
    for (int i = 0; i < DD; ++i) {
      Q[i] = T(0);
      R[i] = T(0);
    }
#pragma unroll
    for (int i = 0; i < D; ++i) {
      rv[i] = T(0);
      x[i] = T(0);
    }
    if (live_ && k0_ > 0) {  // Q = -A_{k0-1}: the block that couples the segment to its predecessor
      load_vec<T, DD>(Q, p.sub + (c * (p.Tn - 1) + k0_ - 1) * DD);
#pragma unroll
      for (int i = 0; i < DD; ++i) Q[i] = -Q[i];
    }
  }
  __device__ __forceinline__ void tile(const Params&, const T* const* in, T* const*, int64_t j0, int ns) {
    if (!live_) return;
    for (int j = 0; j < ns; ++j) {
      const int64_t k = k0_ + j0 + j;
      T S[DD], r[D];
      ld_s<T, DD>(S, in[0] + j * DD);
      if (RHS) ld_s<T, D>(r, in[RHS ? 2 : 0] + j * D);
      if (started_) {
        // extend by step k: W = L^{-1} Q, Q' = -Ls W, R' = R - W^T W, r' = r + W^T x
        T W[DD], Qn[DD];
#pragma unroll
        for (int i = 0; i < DD; ++i) W[i] = Q[i];
        trsm_left_lower<T, D>(Lf, rinv, W);
        gemm<T, D>(Qn, Ls, W);
#pragma unroll
        for (int i = 0; i < DD; ++i) Q[i] = -Qn[i];
#pragma unroll
        for (int a = 0; a < D; ++a)
#pragma unroll
          for (int b = 0; b <= a; ++b) {
            T v = R[a * D + b];
#pragma unroll
            for (int s = 0; s < D; ++s) v = Num<T>::fma(-W[s * D + a], W[s * D + b], v);
            R[a * D + b] = v;
            R[b * D + a] = v;
          }
        if (RHS) gemv_t_add<T, D>(rv, W, x);
        syrk_sub_lower<T, D>(S, Ls);
        if (RHS) gemv_sub<T, D>(r, Ls, x);
      }
      started_ = true;
#pragma unroll
      for (int i = 0; i < DD; ++i) Pm[i] = S[i];
#pragma unroll
      for (int i = 0; i < D; ++i) pv[i] = RHS ? r[i] : T(0);
#pragma unroll
      for (int i = 0; i < DD; ++i) Lf[i] = S[i];
      const bool ok = chol_lower<T, D>(Lf, rinv);
      if (!ok && fail == 0) fail = (int32_t)(k + 1);
      if (RHS) {
        trsv_lower<T, D>(Lf, rinv, r);
#pragma unroll
        for (int i = 0; i < D; ++i) x[i] = r[i];
      }
      if (k + 1 < Tn_) {
        ld_s<T, DD>(Ls, in[1] + j * DD);
        trsm_right_lower_t<T, D>(Ls, Lf, rinv);
      }
    }
  }
  __device__ __forceinline__ void finish(const Params& p, int64_t v, bool valid) {
    if (!valid || !live_) return;
    const int64_t c = v / p.P;
    const int64_t kl = k0_ + p.L - 1;
    mirror_lower<T, D>(Pm);
    store_vec<T, DD>(p.od + (c * p.Tn + kl) * DD, Pm);
    store_vec<T, DD>(p.od + (c * p.Tn + kl - 1) * DD, R);
    store_vec<T, DD>(p.os + (c * (p.Tn - 1) + kl) * DD, Q);
    if (RHS) {
      store_vec<T, D>(p.ox + (c * p.Tn + kl) * D, pv);
      store_vec<T, D>(p.ox + (c * p.Tn + kl - 1) * D, rv);
    }
    if (fail && p.info) atomic_min_nonzero(p.info + c, fail);
  }
};

// ---- pass 2: fold of the elements -------------------------------------------------------------
// element of segment [k0, k0 + L) parked by CholPitSummaryCore (kl = its last step)
template <typename T, int D, bool RHS>
__device__ __forceinline__ void chol_elem_load(LftElem<T, D, RHS>& e, const CholPitParams<T>& prm,
                                               int64_t c, int64_t kl) {
  constexpr int DD = D * D;
  e.empty = 0;
  load_vec_rw<T, DD>(e.P, prm.od + (c * prm.Tn + kl) * DD);
  load_vec_rw<T, DD>(e.R, prm.od + (c * prm.Tn + kl - 1) * DD);
  load_vec_rw<T, DD>(e.Q, prm.os + (c * (prm.Tn - 1) + kl) * DD);
#pragma unroll
  for (int i = 0; i < D; ++i) e.p[i] = e.r[i] = T(0);
  if (RHS) {
    load_vec_rw<T, D>(e.p, prm.ox + (c * prm.Tn + kl) * D);
    load_vec_rw<T, D>(e.r, prm.ox + (c * prm.Tn + kl - 1) * D);
  }
}

// WARP = false: one thread per chain walks the segments in order.  WARP = true (many segments): one
// warp per chain -- lane l owns segments [l*m, (l+1)*m), combines their elements, the warp scans the
// 32 composites, and every lane then walks its segments from its exclusive prefix.  For every
// segment s >= 1 the pair (Ls_{k0-1}, x_{k0-1}) that its first step needs is parked in its last
// output slots (which held the segment's own element: loaded before it is overwritten).
template <typename T, int D, bool RHS, bool WARP>
__global__ void __launch_bounds__(128)
chol_pit_seed_kernel(const CholPitParams<T> p) {
  constexpr int DD = D * D;
  using Elem = LftElem<T, D, RHS>;
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t c = WARP ? tid / 32 : tid;
  const int lane = WARP ? (int)(tid & 31) : 0;
  if (c >= p.B) return;
  // live segments (complete, with a successor): 0 .. nlive-1
  const int64_t nlive = (p.Tn - 1) / p.L;  // k0 + L < Tn  <=>  seg < (Tn - 1) / L  (integer division)
  const int64_t m = WARP ? (p.P + 31) / 32 : p.P;
  const int64_t s0 = lane * m;
  int64_t s1 = s0 + m;
  if (s1 > p.P) s1 = p.P;
  int32_t fail = 0;
  Elem X;
  X.clear();
  if (WARP) {
    Elem e, other;
    for (int64_t seg = s0; seg < s1 && seg < nlive; ++seg) {
      chol_elem_load<T, D, RHS>(e, p, c, seg * p.L + p.L - 1);
      if (!X.then(e) && fail == 0) fail = (int32_t)(seg * p.L + 1);
    }
#pragma unroll 1
    for (int delta = 1; delta < 32; delta <<= 1) {
      other.shfl_up_from(X, delta);
      if (lane >= delta) {
        if (!other.then(X) && fail == 0) fail = (int32_t)(s0 * p.L + 1);
        X = other;
      }
    }
    other.shfl_up_from(X, 1);  // exclusive prefix: the inclusive composite of the previous lane
    X = other;
    if (lane == 0) X.clear();
  }
  for (int64_t seg = s0; seg < s1; ++seg) {
    const int64_t k0 = seg * p.L;
    const int64_t n = seg_steps(p.Tn, k0, p.L);
    if (n <= 0) break;
    const int64_t kl = k0 + n - 1;
    const bool live = seg < nlive;
    Elem e;
    e.clear();
    if (live) chol_elem_load<T, D, RHS>(e, p, c, kl);  // before its slots receive the seed
    if (seg > 0) {
      // state after segments 0..seg-1 = (P, p) of their combined element
      T Lf[DD], rinv[D], A[DD], xs[D];
#pragma unroll
      for (int i = 0; i < DD; ++i) Lf[i] = X.P[i];
      const bool ok = chol_lower<T, D>(Lf, rinv);
      if (!ok && fail == 0) fail = (int32_t)k0;
      load_vec<T, DD>(A, p.sub + (c * (p.Tn - 1) + k0 - 1) * DD);
      trsm_right_lower_t<T, D>(A, Lf, rinv);  // Ls_{k0-1} = A_{k0-1} Ld^{-T}
      store_vec<T, DD>(p.od + (c * p.Tn + kl) * DD, A);
      if (RHS) {
#pragma unroll
        for (int i = 0; i < D; ++i) xs[i] = X.p[i];
        trsv_lower<T, D>(Lf, rinv, xs);  // x_{k0-1} = Ld^{-1} r
        store_vec<T, D>(p.ox + (c * p.Tn + kl) * D, xs);
      }
    }
    if (!live) break;
    if (!X.then(e) && fail == 0) fail = (int32_t)(k0 + 1);
  }
  if (fail && p.info) atomic_min_nonzero(p.info + c, fail);
}

}  // namespace mf
