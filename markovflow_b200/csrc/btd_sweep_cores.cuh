// The remaining sequential block-tridiagonal recurrences on the TMA chain sweep (sweep.cuh): the
// arithmetic of the thread-per-chain kernels of btd_direct.cuh with every per-step record streamed
// through the shared-memory ring instead of being read from global memory on the critical path.
//
//   BtdSolveCore      forward / backward : LowerTriangularBlockTriDiagonal.solve
//                                          (reference block_tri_diag.py:339-351, solve_triang_mat :350)
//   BtdInvSubsetCore  backward           : block_diagonal_of_inverse + sub-diagonal blocks
//                                          (:318-337, inverse_from_cholesky_band :331)
//   BtdUduCore        backward           : upper_diagonal_lower (:438-545)
#pragma once
#include "ssm_sweep.cuh"

namespace mf {

// ---------------------------------------------------------------------------------------------
// Triangular solve.  The recursion is affine in x, so few long chains are evaluated parallel in time
// (exactly): every segment composes its steps into (Phi, c) with x_out = Phi x_in + c (SUMMARY pass:
// the recursion is run on c and on the D columns of Phi), a per-chain fold gives the x entering every
// segment, and the segments then run the ordinary sweep from their seeds.  No workspace: the D+1
// vectors of an element and the seed are parked in the output slots the owning segment writes last.
template <typename T>
struct BtdSolveParams {
  const T *ld, *ls, *rhs;
  T* out;
  int64_t n, Bm, Tn;
  int64_t P, L;  // segments per chain, steps per segment (P == 1: L == Tn)
};

// output step that holds parked vector i (0: c | seed, 1..D: columns of Phi) of segment [k0, k0+n)
template <bool TRANSPOSE>
__device__ __forceinline__ int64_t solve_slot(int64_t k0, int64_t n, int i) {
  return TRANSPOSE ? k0 + i : k0 + n - 1 - i;
}

// forward : x_k = Ld_k^{-1} (b_k - Ls_{k-1} x_{k-1});  backward: x_k = Ld_k^{-T} (b_k - Ls_k^T x_{k+1}).
// UNIT: identity diagonal blocks (ld == NULL);  SUB: a sub-diagonal exists.
template <typename T_, int D, bool TRANSPOSE, bool UNIT, bool SUB, bool SUMMARY = false>
struct BtdSolveCore {
  using T = T_;
  using Params = BtdSolveParams<T>;
  static constexpr int DD = D * D;
  static constexpr int NIN = 1 + (UNIT ? 0 : 1) + (SUB ? 1 : 0), NOUT = SUMMARY ? 0 : 1;
  static constexpr bool BACKWARD = TRANSPOSE;
  static constexpr int I_LD = UNIT ? -1 : 1, I_LS = SUB ? (UNIT ? 1 : 2) : -1;
  static constexpr int ein(int i) { return i == 0 ? D : DD; }
  static constexpr int eout(int) { return D; }
  static __device__ __forceinline__ int64_t num_chains(const Params& p) { return p.n * p.P; }
  static __device__ __forceinline__ int64_t max_steps(const Params& p) { return p.L; }
  // a segment is summarised iff another segment consumes its result
  static __device__ __forceinline__ bool is_live(const Params& p, int64_t v) {
    const int64_t seg = v % p.P, k0 = seg * p.L;
    return TRANSPOSE ? (seg > 0 && k0 < p.Tn) : (k0 + p.L < p.Tn);
  }
  static __device__ __forceinline__ StreamGeom in_geom(const Params& p, int i, int64_t v) {
    const int64_t c = v / p.P, k0 = (v % p.P) * p.L;
    int64_t n = seg_steps(p.Tn, k0, p.L);
    if (SUMMARY && !is_live(p, v)) n = 0;
    if (i == 0) return vgeom_states<T>(p.rhs, c, p.Tn, D, k0, n);
    if (i == I_LD) return vgeom_states<T>(p.ld, c % p.Bm, p.Tn, DD, k0, n);
    // forward needs Ls_{k-1} at step k (incoming), backward needs Ls_k at step k (outgoing)
    return TRANSPOSE ? vgeom_outgoing<T>(p.ls, c % p.Bm, p.Tn, DD, k0, n)
                     : vgeom_incoming<T>(p.ls, c % p.Bm, p.Tn, DD, k0, n);
  }
  static __device__ __forceinline__ StreamGeom out_geom(const Params& p, int, int64_t v) {
    if (SUMMARY) return StreamGeom{nullptr, 0, 0};
    const int64_t c = v / p.P, k0 = (v % p.P) * p.L;
    return vgeom_states<T>(p.out, c, p.Tn, D, k0, seg_steps(p.Tn, k0, p.L));
  }
  T x[D];
  T Phi[SUMMARY ? DD : 1];  // column q at Phi[q * D ..]
  int64_t Tn_, k0_, n_;
  bool live_;
  __device__ __forceinline__ void init(const Params& p, int64_t v) {
    const int64_t c = v / p.P;
    Tn_ = p.Tn;
    k0_ = (v % p.P) * p.L;
    n_ = seg_steps(p.Tn, k0_, p.L);
    live_ = !SUMMARY || is_live(p, v);
#pragma unroll
    for (int i = 0; i < D; ++i) x[i] = T(0);
    if (SUMMARY) {
#pragma unroll
      for (int i = 0; i < DD; ++i) Phi[SUMMARY ? i : 0] = (i / D == i % D) ? T(1) : T(0);
    } else if (SUB && n_ > 0 && (TRANSPOSE ? k0_ + n_ < p.Tn : k0_ > 0)) {
      load_vec_rw<T, D>(x, p.out + (c * p.Tn + solve_slot<TRANSPOSE>(k0_, n_, 0)) * D);  // seed
    }
  }
  // r <- Ld^{-1} (r - Ls xin)  (or the transposed form); blocks of local step j
  __device__ __forceinline__ void apply(T* __restrict__ r, const T* __restrict__ xin, const T* const* in,
                                        int j, bool coupled) {
    if (SUB && coupled) {
      T A[DD];
      ld_s<T, DD>(A, in[I_LS < 0 ? 0 : I_LS] + j * DD);
      if (TRANSPOSE) gemv_t_sub<T, D>(r, A, xin);
      else gemv_sub<T, D>(r, A, xin);
    }
    if (!UNIT) {
      T L[DD], rinv[D];
      ld_s<T, DD>(L, in[I_LD < 0 ? 0 : I_LD] + j * DD);
#pragma unroll
      for (int q = 0; q < D; ++q) rinv[q] = Num<T>::rcp(L[q * D + q]);
      if (TRANSPOSE) trsv_lower_t<T, D>(L, rinv, r);
      else trsv_lower<T, D>(L, rinv, r);
    }
  }
  __device__ __forceinline__ void step(const T* const* in, T* const* out, int j, int64_t k) {
    const bool coupled = TRANSPOSE ? (k + 1 < Tn_) : (k > 0);
    T r[D];
    ld_s<T, D>(r, in[0] + j * D);
    apply(r, x, in, j, coupled);
    if (SUMMARY) {
#pragma unroll
      for (int q = 0; q < D; ++q) {
        T h[D];
#pragma unroll
        for (int i = 0; i < D; ++i) h[i] = T(0);
        apply(h, Phi + (SUMMARY ? q * D : 0), in, j, coupled);
#pragma unroll
        for (int i = 0; i < D; ++i) Phi[SUMMARY ? q * D + i : 0] = coupled ? h[i] : T(0);
      }
    }
#pragma unroll
    for (int i = 0; i < D; ++i) x[i] = r[i];
    if (!SUMMARY) st_s<T, D>(out[0] + j * D, x);
  }
  __device__ __forceinline__ void tile(const Params&, const T* const* in, T* const* out, int64_t j0,
                                       int ns) {
    if (!live_) return;
    if (n_ - j0 < ns) ns = (int)(n_ - j0);
    if (TRANSPOSE) {
      for (int j = ns - 1; j >= 0; --j) step(in, out, j, k0_ + j0 + j);
    } else {
      for (int j = 0; j < ns; ++j) step(in, out, j, k0_ + j0 + j);
    }
  }
  __device__ __forceinline__ void finish(const Params& p, int64_t v, bool valid) {
    if (!SUMMARY || !valid || !live_) return;
    const int64_t c = v / p.P;
    store_vec<T, D>(p.out + (c * p.Tn + solve_slot<TRANSPOSE>(k0_, n_, 0)) * D, x);
#pragma unroll
    for (int q = 0; q < D; ++q)
      store_vec<T, D>(p.out + (c * p.Tn + solve_slot<TRANSPOSE>(k0_, n_, q + 1)) * D,
                      Phi + (SUMMARY ? q * D : 0));
  }
};

// fold of the segment elements of one rhs chain in sweep order (affine_fold in ssm_sweep.cuh); parks
// the x entering every segment in that segment's seed slot.
template <typename T, int D, bool TRANSPOSE, bool WARP>
__global__ void __launch_bounds__(128)
btd_solve_seed_kernel(const BtdSolveParams<T> p) {
  affine_fold<T, D, TRANSPOSE, WARP>(p.out, p.n, p.Tn, p.P, p.L, [](int64_t k0, int64_t n, int i) {
    return solve_slot<TRANSPOSE>(k0, n, i);
  });
}

// ---------------------------------------------------------------------------------------------
// Sparse inverse subset.  Sigma_kk = G_k + J_k^T Sigma_{k+1,k+1} J_k is affine in Sigma, so few long
// chains run parallel in time: elements (Phi, Gt) with Sigma_out = Phi^T Sigma_in Phi + Gt
// (Phi <- Phi J_k, Gt <- J_k^T Gt J_k + G_k going down), per-chain fold, seeded sweeps.  Elements and
// seeds are parked in the output slots of each segment's FIRST steps (Gt | seed -> od[k0],
// Phi -> od[k0+1]).
template <typename T>
struct BtdInvSubsetParams {
  const T *ld, *ls;
  T *od, *os;
  int64_t B, Tn;
  int64_t P, L;
};

// Sigma_{T-1,T-1} = (Ld Ld^T)^{-1};  J_k = Ls_k Ld_k^{-1};  Sigma_{k+1,k} = -Sigma_{k+1,k+1} J_k;
// Sigma_kk = (Ld_k Ld_k^T)^{-1} - J_k^T Sigma_{k+1,k}
template <typename T_, int D, bool WANT_SUB, bool SUMMARY = false>
struct BtdInvSubsetCore {
  using T = T_;
  using Params = BtdInvSubsetParams<T>;
  static constexpr int DD = D * D;
  static constexpr int NIN = 2, NOUT = SUMMARY ? 0 : (WANT_SUB ? 2 : 1);
  static constexpr bool BACKWARD = true;
  static constexpr int ein(int) { return DD; }
  static constexpr int eout(int) { return DD; }
  static __device__ __forceinline__ int64_t num_chains(const Params& p) { return p.B * p.P; }
  static __device__ __forceinline__ int64_t max_steps(const Params& p) { return p.L; }
  static __device__ __forceinline__ StreamGeom in_geom(const Params& p, int i, int64_t v) {
    const int64_t c = v / p.P, seg = v % p.P, k0 = seg * p.L;
    int64_t n = seg_steps(p.Tn, k0, p.L);
    if (SUMMARY && seg == 0) n = 0;  // the first segment feeds nobody
    if (i == 0) return vgeom_states<T>(p.ld, c, p.Tn, DD, k0, n);
    return vgeom_outgoing<T>(p.ls, c, p.Tn, DD, k0, n);
  }
  static __device__ __forceinline__ StreamGeom out_geom(const Params& p, int i, int64_t v) {
    if (SUMMARY) return StreamGeom{nullptr, 0, 0};
    const int64_t c = v / p.P, k0 = (v % p.P) * p.L;
    const int64_t n = seg_steps(p.Tn, k0, p.L);
    if (i == 0) return vgeom_states<T>(p.od, c, p.Tn, DD, k0, n);
    return vgeom_outgoing<T>(p.os, c, p.Tn, DD, k0, n);
  }
  T sig[DD];                 // Sigma of the step after the current one | Gt in the summary pass
  T Phi[SUMMARY ? DD : 1];
  int64_t Tn_, k0_, n_;
  bool live_;
  __device__ __forceinline__ void init(const Params& p, int64_t v) {
    const int64_t c = v / p.P;
    Tn_ = p.Tn;
    k0_ = (v % p.P) * p.L;
    n_ = seg_steps(p.Tn, k0_, p.L);
    live_ = n_ > 0 && (!SUMMARY || (v % p.P) > 0);
#pragma unroll
    for (int i = 0; i < DD; ++i) sig[i] = T(0);
    if (SUMMARY) {
#pragma unroll
      for (int i = 0; i < DD; ++i) Phi[SUMMARY ? i : 0] = (i / D == i % D) ? T(1) : T(0);
    } else if (n_ > 0 && k0_ + n_ < p.Tn) {
      load_vec_rw<T, DD>(sig, p.od + (c * p.Tn + k0_) * DD);  // seed: Sigma entering the segment
    }
  }
  __device__ __forceinline__ void tile(const Params&, const T* const* in, T* const* out, int64_t j0,
                                       int ns) {
    if (!live_) return;
    if (n_ - j0 < ns) ns = (int)(n_ - j0);
    for (int j = ns - 1; j >= 0; --j) {
      const int64_t k = k0_ + j0 + j;
      T L[DD], loc[DD], rinv[D];
      ld_s<T, DD>(L, in[0] + j * DD);
#pragma unroll
      for (int q = 0; q < D; ++q) rinv[q] = Num<T>::rcp(L[q * D + q]);
      chol_inverse<T, D>(loc, L, rinv);
      if (k + 1 < Tn_) {
        T J[DD], ssub[DD];
        ld_s<T, DD>(J, in[1] + j * DD);
        trsm_right_lower<T, D>(J, L, rinv);  // J = Ls Ld^{-1}
        gemm<T, D>(ssub, sig, J);            // Sigma_{k+1,k+1} J   (summary: Gt J)
#pragma unroll
        for (int i = 0; i < DD; ++i) ssub[i] = -ssub[i];
        if (!SUMMARY && WANT_SUB) st_s<T, DD>(out[WANT_SUB ? 1 : 0] + j * DD, ssub);
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
          for (int q = 0; q <= i; ++q) {
            T v = loc[i * D + q];
#pragma unroll
            for (int s = 0; s < D; ++s) v = Num<T>::fma(-J[s * D + i], ssub[s * D + q], v);
            loc[i * D + q] = v;
          }
        mirror_lower<T, D>(loc);
        if (SUMMARY) {
          T t[DD];
          gemm<T, D>(t, Phi, J);
#pragma unroll
          for (int i = 0; i < DD; ++i) Phi[SUMMARY ? i : 0] = t[i];
        }
      } else if (SUMMARY) {
        // k = T-1: Sigma = G, nothing enters: the element is (0, G)
#pragma unroll
        for (int i = 0; i < DD; ++i) Phi[SUMMARY ? i : 0] = T(0);
      }
      if (!SUMMARY) st_s<T, DD>(out[0] + j * DD, loc);
#pragma unroll
      for (int i = 0; i < DD; ++i) sig[i] = loc[i];
    }
  }
  __device__ __forceinline__ void finish(const Params& p, int64_t v, bool valid) {
    if (!SUMMARY || !valid || !live_) return;
    const int64_t c = v / p.P;
    store_vec<T, DD>(p.od + (c * p.Tn + k0_) * DD, sig);
    if (n_ >= 2) store_vec<T, DD>(p.od + (c * p.Tn + k0_ + 1) * DD, Phi);
  }
};

// Element of a range of steps of the inverse-subset recursion: Sigma_out = Phi^T Sigma_in Phi + Gt.
// (e1 then e2): Phi = Phi1 Phi2, Gt = Phi2^T Gt1 Phi2 + Gt2.
template <typename T, int D>
struct CongElem {
  T Phi[D * D], Gt[D * D];
  int empty;
  __device__ __forceinline__ void clear() {
    empty = 1;
#pragma unroll
    for (int i = 0; i < D * D; ++i) {
      Phi[i] = (i / D == i % D) ? T(1) : T(0);
      Gt[i] = T(0);
    }
  }
  __device__ __forceinline__ void then(const CongElem& e2) {
    if (e2.empty) return;
    if (empty) {
      *this = e2;
      return;
    }
    T t[D * D], GP[D * D];
    gemm<T, D>(t, Phi, e2.Phi);
#pragma unroll
    for (int i = 0; i < D * D; ++i) Phi[i] = t[i];
    gemm<T, D>(GP, Gt, e2.Phi);  // Gt1 Phi2
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
      for (int q = 0; q <= r; ++q) {
        T v = e2.Gt[r * D + q];
#pragma unroll
        for (int s = 0; s < D; ++s) v = Num<T>::fma(e2.Phi[s * D + r], GP[s * D + q], v);
        Gt[r * D + q] = v;
        Gt[q * D + r] = v;
      }
  }
  __device__ __forceinline__ void shfl_up_from(const CongElem& src, int delta) {
#pragma unroll
    for (int i = 0; i < D * D; ++i) {
      Phi[i] = __shfl_up_sync(0xffffffffu, src.Phi[i], delta);
      Gt[i] = __shfl_up_sync(0xffffffffu, src.Gt[i], delta);
    }
    empty = __shfl_up_sync(0xffffffffu, src.empty, delta);
  }
};

// fold from the last segment down (sweep position it <-> segment P-1-it); parks the Sigma entering
// every segment s < P-1 in od[k0].  The last segment has nothing entering (its Phi is 0), so the
// state after a prefix of the sweep is simply Gt of the combined element.  WARP: one warp per
// chain, lane l owns sweep positions [l*m, (l+1)*m).
template <typename T, int D, bool WARP>
__global__ void __launch_bounds__(128)
btd_inv_subset_seed_kernel(const BtdInvSubsetParams<T> p) {
  constexpr int DD = D * D;
  using Elem = CongElem<T, D>;
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t c = WARP ? tid / 32 : tid;
  const int lane = WARP ? (int)(tid & 31) : 0;
  if (c >= p.B) return;
  const int64_t m = WARP ? (p.P + 31) / 32 : p.P;
  const int64_t i0 = lane * m;
  int64_t i1 = i0 + m;
  if (i1 > p.P) i1 = p.P;
  auto load = [&](Elem& e, int64_t it) {
    const int64_t k0 = (p.P - 1 - it) * p.L;
    e.clear();
    e.empty = 0;
    load_vec_rw<T, DD>(e.Gt, p.od + (c * p.Tn + k0) * DD);
    if (it > 0) load_vec_rw<T, DD>(e.Phi, p.od + (c * p.Tn + k0 + 1) * DD);
    else {
#pragma unroll
      for (int i = 0; i < DD; ++i) e.Phi[i] = T(0);
    }
  };
  Elem X;
  X.clear();
  if (WARP) {
    Elem e, other;
    for (int64_t it = i0; it < i1 && it < p.P - 1; ++it) {  // segment 0 feeds nobody
      load(e, it);
      X.then(e);
    }
#pragma unroll 1
    for (int delta = 1; delta < 32; delta <<= 1) {
      other.shfl_up_from(X, delta);
      if (lane >= delta) {
        other.then(X);
        X = other;
      }
    }
    other.shfl_up_from(X, 1);
    X = other;
    if (lane == 0) X.clear();
  }
  for (int64_t it = i0; it < i1; ++it) {
    const int64_t k0 = (p.P - 1 - it) * p.L;
    Elem e;
    e.clear();
    const bool live = it + 1 < p.P;
    if (live) load(e, it);  // before its slot receives the seed
    if (it > 0) store_vec<T, DD>(p.od + (c * p.Tn + k0) * DD, X.Gt);
    if (!live) break;
    X.then(e);
  }
}

// ---------------------------------------------------------------------------------------------
// U D U^T.  D_k = K_kk - K_{k+1,k}^T D_{k+1}^{-1} K_{k+1,k} is the linear-fractional recursion of
// ssm_sweep.cuh (naturals -> SSM) without the vector part: elements D_out = P - Q (D_in + R)^{-1} Q^T,
// extension by a step  P' = K_kk - Ks^T P^{-1} Ks,  Q' = Ks^T P^{-1} Q,  R' = R - Q^T P^{-1} Q.
// Few long chains: summary pass -> per-chain fold -> seeded sweeps; elements and seeds are parked in
// the output slots of each segment's FIRST steps (P | seed D -> ocd[k0], R -> ocd[k0+1], Q -> ou[k0]).
template <typename T>
struct BtdUduParams {
  const T *diag, *sub;
  T *ou, *ocd;
  int32_t* info;
  int64_t B, Tn;
  int64_t P, L;
};

// cholD_{T-1} = chol(K_{T-1,T-1});  U_k^T = D_{k+1}^{-1} K_{k+1,k};  D_k = K_kk - K_{k+1,k}^T U_k^T
template <typename T_, int D, bool SUMMARY = false>
struct BtdUduCore {
  using T = T_;
  using Params = BtdUduParams<T>;
  static constexpr int DD = D * D;
  static constexpr int NIN = 2, NOUT = SUMMARY ? 0 : 2;
  static constexpr bool BACKWARD = true;
  static constexpr int ein(int) { return DD; }
  static constexpr int eout(int) { return DD; }
  static __device__ __forceinline__ int64_t num_chains(const Params& p) { return p.B * p.P; }
  static __device__ __forceinline__ int64_t max_steps(const Params& p) { return p.L; }
  static __device__ __forceinline__ StreamGeom in_geom(const Params& p, int i, int64_t v) {
    const int64_t c = v / p.P, seg = v % p.P, k0 = seg * p.L;
    int64_t n = seg_steps(p.Tn, k0, p.L);
    if (SUMMARY && seg == 0) n = 0;  // the first segment feeds nobody
    if (i == 0) return vgeom_states<T>(p.diag, c, p.Tn, DD, k0, n);
    return vgeom_outgoing<T>(p.sub, c, p.Tn, DD, k0, n);
  }
  static __device__ __forceinline__ StreamGeom out_geom(const Params& p, int i, int64_t v) {
    if (SUMMARY) return StreamGeom{nullptr, 0, 0};
    const int64_t c = v / p.P, k0 = (v % p.P) * p.L;
    const int64_t n = seg_steps(p.Tn, k0, p.L);
    if (i == 0) return vgeom_outgoing<T>(p.ou, c, p.Tn, DD, k0, n);
    return vgeom_states<T>(p.ocd, c, p.Tn, DD, k0, n);
  }
  T C[DD], rinv[D];
  T Pm[SUMMARY ? DD : 1], Q[SUMMARY ? DD : 1], R[SUMMARY ? DD : 1];
  int32_t fail;
  int64_t Tn_, k0_, n_;
  bool live_, have_;  // have_: C holds the factor of the block after the current step
  __device__ __forceinline__ void init(const Params& p, int64_t v) {
    const int64_t c = v / p.P;
    Tn_ = p.Tn;
    fail = 0;
    k0_ = (v % p.P) * p.L;
    n_ = seg_steps(p.Tn, k0_, p.L);
    live_ = n_ > 0 && (!SUMMARY || (v % p.P) > 0);
    have_ = false;
    if (!SUMMARY && n_ > 0 && k0_ + n_ < p.Tn) {  // seed: the D entering this segment
      load_vec_rw<T, DD>(C, p.ocd + (c * p.Tn + k0_) * DD);
      const bool ok = chol_lower<T, D>(C, rinv);
      if (!ok) fail = (int32_t)(k0_ + n_ + 1);
      have_ = true;
    }
  }
  __device__ __forceinline__ void tile(const Params&, const T* const* in, T* const* out, int64_t j0,
                                       int ns) {
    if (!live_) return;
    if (n_ - j0 < ns) ns = (int)(n_ - j0);
    for (int j = ns - 1; j >= 0; --j) {
      const int64_t k = k0_ + j0 + j;
      T Dk[DD];
      ld_s<T, DD>(Dk, in[0] + j * DD);
      if (k + 1 < Tn_) {
        T K[DD];
        ld_s<T, DD>(K, in[1] + j * DD);
        if (have_) {
          T X[DD];
#pragma unroll
          for (int i = 0; i < DD; ++i) X[i] = K[i];
          trsm_left_lower<T, D>(C, rinv, X);
          trsm_left_lower_t<T, D>(C, rinv, X);  // X = D_{k+1}^{-1} K_{k+1,k}
          if (!SUMMARY) st_s<T, DD>(out[0] + j * DD, X);
          if (SUMMARY) {
            // Q' = K^T W, R' = R - Q^T W with W = D_{k+1}^{-1} Q
            T W[DD], Qn[DD];
#pragma unroll
            for (int i = 0; i < DD; ++i) W[i] = Q[SUMMARY ? i : 0];
            trsm_left_lower<T, D>(C, rinv, W);
            trsm_left_lower_t<T, D>(C, rinv, W);
#pragma unroll
            for (int a = 0; a < D; ++a)
#pragma unroll
              for (int b = 0; b < D; ++b) {
                T vr = R[SUMMARY ? a * D + b : 0], vq = T(0);
#pragma unroll
                for (int s = 0; s < D; ++s) {
                  vr = Num<T>::fma(-Q[SUMMARY ? s * D + a : 0], W[s * D + b], vr);
                  vq = Num<T>::fma(K[s * D + a], W[s * D + b], vq);
                }
                R[SUMMARY ? a * D + b : 0] = vr;
                Qn[a * D + b] = vq;
              }
#pragma unroll
            for (int i = 0; i < DD; ++i) Q[SUMMARY ? i : 0] = Qn[i];
          }
#pragma unroll
          for (int i = 0; i < D; ++i)
#pragma unroll
            for (int q = 0; q <= i; ++q) {
              T v = Dk[i * D + q];
#pragma unroll
              for (int s = 0; s < D; ++s) v = Num<T>::fma(-K[s * D + i], X[s * D + q], v);
              Dk[i * D + q] = v;
            }
        } else if (SUMMARY) {
          // far end of the segment: P = K_kk, Q = K_{k+1,k}^T, R = 0
#pragma unroll
          for (int a = 0; a < D; ++a)
#pragma unroll
            for (int b = 0; b < D; ++b) {
              Q[SUMMARY ? a * D + b : 0] = K[b * D + a];
              R[SUMMARY ? a * D + b : 0] = T(0);
            }
        }
      }
      if (SUMMARY) {
#pragma unroll
        for (int i = 0; i < DD; ++i) Pm[SUMMARY ? i : 0] = Dk[i];
      }
#pragma unroll
      for (int i = 0; i < DD; ++i) C[i] = Dk[i];
      const bool ok = chol_lower<T, D>(C, rinv);
      if (!ok && fail == 0) fail = (int32_t)(k + 1);
      have_ = true;
      if (!SUMMARY) {
        zero_upper<T, D>(C);
        st_s<T, DD>(out[1] + j * DD, C);
      }
    }
  }
  __device__ __forceinline__ void finish(const Params& p, int64_t v, bool valid) {
    if (!valid) return;
    const int64_t c = v / p.P;
    if (SUMMARY) {
      if (!live_) return;
      mirror_lower<T, D>(Pm);
      store_vec<T, DD>(p.ocd + (c * p.Tn + k0_) * DD, Pm);
      if (n_ >= 2) store_vec<T, DD>(p.ocd + (c * p.Tn + k0_ + 1) * DD, R);
      if (k0_ < p.Tn - 1) store_vec<T, DD>(p.ou + (c * (p.Tn - 1) + k0_) * DD, Q);
      if (fail && p.info) atomicMax(p.info + c, fail);
      return;
    }
    if (!p.info) return;
    if (p.P == 1) p.info[v] = fail;
    else if (fail) atomicMax(p.info + c, fail);
  }
};

// fold of the elements from the last segment down (lft_backward_fold in ssm_sweep.cuh); parks the D
// entering every segment s < P-1 in ocd[k0]
struct UduFoldPolicy {
  template <typename T, int D>
  static __device__ __forceinline__ void load(LftElem<T, D, false>& e, const BtdUduParams<T>& p,
                                              int64_t c, int64_t k0, int64_t n, bool last) {
    constexpr int DD = D * D;
    e.clear();
    e.empty = 0;
    load_vec_rw<T, DD>(e.P, p.ocd + (c * p.Tn + k0) * DD);
    if (!last) {
      load_vec_rw<T, DD>(e.R, p.ocd + (c * p.Tn + k0 + 1) * DD);
      load_vec_rw<T, DD>(e.Q, p.ou + (c * (p.Tn - 1) + k0) * DD);
    }
  }
  template <typename T, int D>
  static __device__ __forceinline__ void seed(const BtdUduParams<T>& p, int64_t c, int64_t k0,
                                              const LftElem<T, D, false>& X) {
    store_vec<T, D * D>(p.ocd + (c * p.Tn + k0) * D * D, X.P);
  }
};

template <typename T, int D, bool WARP>
__global__ void __launch_bounds__(128)
btd_udu_seed_kernel(const BtdUduParams<T> p) {
  lft_backward_fold<T, D, false, UduFoldPolicy, BtdUduParams<T>, WARP>(p, p.info);
}

}  // namespace mf
