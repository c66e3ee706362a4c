// The sequential StateSpaceModel / natural-parameter recurrences on the TMA chain sweep
// (sweep.cuh): same arithmetic as the thread-per-chain kernels of ssm_kernels.cuh / nat_kernels.cuh,
// with every per-step record streamed through the shared-memory ring instead of being read from
// global memory on the critical path.
//
//   SsmMomentsCore   forward : marginal means / covariances / lag-one blocks  (mf_ssm_marginals)
//                              or the expectation parameters                  (mf_ssm_to_expectations)
//   SsmAffineCore    forward : x_k = A x_{k-1} + b (+ chol_q eps)             (mf_ssm_affine_scan)
//   SsmKlCore        forward : KL(q || p), chain-rule form                    (mf_ssm_kl_divergence)
//   NatToSsmCore     backward: naturals -> SSM parameters, U D U^T sweep      (mf_nat_to_ssm)
#pragma once
#include <type_traits>

#include "dispatch.cuh"
#include "nat_kernels.cuh"
#include "smallrng.cuh"
#include "sweep_tmi.cuh"

namespace mf {

template <typename T>
__device__ __forceinline__ char* byte_ptr(const T* p) {
  return reinterpret_cast<char*>(const_cast<T*>(p));
}

// entry j of a [B,T,...] stream of chain c
template <typename T>
__device__ __forceinline__ StreamGeom geom_states(const T* base, int64_t c, int64_t Tn, int E) {
  StreamGeom g;
  g.step0 = base ? byte_ptr(base) + c * Tn * (int64_t)(E * sizeof(T)) : nullptr;
  g.first = 0;
  g.end = Tn;
  return g;
}
// transition LEAVING step j (entry j of a [B,T-1,...] stream): valid for j < T-1
template <typename T>
__device__ __forceinline__ StreamGeom geom_outgoing(const T* base, int64_t c, int64_t Tn, int E) {
  StreamGeom g;
  g.step0 = base ? byte_ptr(base) + c * (Tn - 1) * (int64_t)(E * sizeof(T)) : nullptr;
  g.first = 0;
  g.end = Tn - 1;
  return g;
}
// transition LEADING INTO step j (entry j-1 of a [B,T-1,...] stream): valid for 1 <= j < T
template <typename T>
__device__ __forceinline__ StreamGeom geom_incoming(const T* base, int64_t c, int64_t Tn, int E) {
  StreamGeom g;
  g.step0 = base ? byte_ptr(base) + (c * (Tn - 1) - 1) * (int64_t)(E * sizeof(T)) : nullptr;
  g.first = 1;
  g.end = Tn;
  return g;
}

template <typename T, int N>
__device__ __forceinline__ void ld_s(T* __restrict__ r, const T* __restrict__ s) {
#pragma unroll
  for (int i = 0; i < N; ++i) r[i] = s[i];
}
template <typename T, int N>
__device__ __forceinline__ void st_s(T* __restrict__ s, const T* __restrict__ r) {
#pragma unroll
  for (int i = 0; i < N; ++i) s[i] = r[i];
}

// ---------------------------------------------------------------------------------------------
// Few long chains are cut into P segments of L steps ("virtual chains", as in kalman_sweep.cuh) and
// evaluated parallel in time -- the moment recursion is affine, so this is exact:
//   1. SsmMomSummaryCore : every segment composes its transitions into one element (Phi, c, Qt):
//                          mu_end = Phi mu + c,  Sigma_end = Phi Sigma Phi^T + Qt
//   2. ssm_moments_seed_kernel : per chain, fold the elements in order -> state before every segment
//   3. SsmMomentsCore    : every segment runs the ordinary recursion from its seed.
// No workspace: elements and seeds live in the output slots of each segment's LAST step, which the
// owning virtual chain overwrites with the final values after it has read its seed
// (Phi | seed Sigma -> o_diag[last], c | seed mu -> o_vec[last], Qt -> o_diag[last-1]; L >= 2).
template <typename T>
struct SsmMomentsParams {
  const T *mu0, *chol_p0, *a, *b, *chol_q;
  T *o_vec, *o_diag, *o_sub;  // means|eta_lin [B,T,D], covs|eta_diag [B,T,D,D], lag-one [B,T-1,D,D]
  int64_t B, Tn;
  int64_t P, L;  // segments per chain, steps per segment (P == 1: L == Tn)
  // where elements / seeds are parked: the outputs themselves (compact == 0: entry c*Tn + kl - which)
  // or a compact workspace (compact == 1: entry (c*P + seg)*2 + which) when there are no outputs
  T *pk_vec, *pk_diag;
  int compact;
};

template <typename T>
__device__ __forceinline__ int64_t mom_park(const SsmMomentsParams<T>& p, int64_t c, int64_t seg,
                                            int64_t kl, int which) {
  return p.compact ? (c * p.P + seg) * 2 + which : c * p.Tn + kl - which;
}

// entry of local step j of segment (c, k0): states / transitions leading into the step
template <typename T>
__device__ __forceinline__ StreamGeom vgeom_states(const T* base, int64_t c, int64_t Tn, int E,
                                                   int64_t k0, int64_t steps) {
  StreamGeom g;
  g.step0 = base ? byte_ptr(base) + (c * Tn + k0) * (int64_t)(E * sizeof(T)) : nullptr;
  g.first = 0;
  g.end = steps;
  return g;
}
template <typename T>
__device__ __forceinline__ StreamGeom vgeom_incoming(const T* base, int64_t c, int64_t Tn, int E,
                                                     int64_t k0, int64_t steps) {
  StreamGeom g;
  g.step0 = base ? byte_ptr(base) + (c * (Tn - 1) + k0 - 1) * (int64_t)(E * sizeof(T)) : nullptr;
  g.first = k0 == 0 ? 1 : 0;
  g.end = steps;
  return g;
}
__device__ __forceinline__ int64_t seg_steps(int64_t Tn, int64_t k0, int64_t L) {
  int64_t n = Tn - k0;
  if (n > L) n = L;
  return n < 0 ? 0 : n;
}

// EXPECT = false: (mu_k, Sigma_kk, A_k Sigma_kk);  EXPECT = true: (mu_k, Sigma_kk + mu mu^T,
// A_k Sigma_kk + mu_{k+1} mu_k^T).  The lag-one block k is produced at step k+1 (incoming form).
template <typename T_, int D, bool EXPECT>
struct SsmMomentsCore {
  using T = T_;
  using Params = SsmMomentsParams<T>;
  static constexpr int DD = D * D;
  static constexpr int NIN = 3, NOUT = 3;
  static constexpr bool BACKWARD = false;
  static constexpr int ein(int i) { return i == 1 ? D : DD; }
  static constexpr int eout(int i) { return i == 0 ? D : DD; }
  static __device__ __forceinline__ int64_t num_chains(const Params& p) { return p.B * p.P; }
  static __device__ __forceinline__ int64_t max_steps(const Params& p) { return p.L; }
  static __device__ __forceinline__ StreamGeom in_geom(const Params& p, int i, int64_t v) {
    const int64_t c = v / p.P, k0 = (v % p.P) * p.L;
    return vgeom_incoming<T>(i == 0 ? p.a : (i == 1 ? p.b : p.chol_q), c, p.Tn, ein(i), k0,
                             seg_steps(p.Tn, k0, p.L));
  }
  static __device__ __forceinline__ StreamGeom out_geom(const Params& p, int i, int64_t v) {
    const int64_t c = v / p.P, k0 = (v % p.P) * p.L;
    const int64_t n = seg_steps(p.Tn, k0, p.L);
    if (i == 2) return vgeom_incoming<T>(p.o_sub, c, p.Tn, DD, k0, n);
    return vgeom_states<T>(i == 0 ? p.o_vec : p.o_diag, c, p.Tn, eout(i), k0, n);
  }
  // in-place stages (sweep_tmi.cuh): a -> lag-one block, b -> mean, chol_q -> covariance of the same step
  static constexpr int tm_alias(int i) { return i == 0 ? 2 : (i == 1 ? 0 : 1); }
  // host-side description of the same streams for the tensor-map engine (sweep_tm.cuh)
  static void tm_describe(const Params& p, TmStream* in, TmStream* out) {
    in[0] = TmStream{p.a, p.Tn - 1, -1};
    in[1] = TmStream{p.b, p.Tn - 1, -1};
    in[2] = TmStream{p.chol_q, p.Tn - 1, -1};
    out[0] = TmStream{p.o_vec, p.Tn, 0};
    out[1] = TmStream{p.o_diag, p.Tn, 0};
    out[2] = TmStream{p.o_sub, p.Tn - 1, -1};
  }
  static int64_t tm_segments(const Params& p) { return p.P; }
  static int64_t tm_seg_len(const Params& p) { return p.L; }
  static int64_t tm_chains(const Params& p) { return p.B; }
  T mu[D], P[DD];
  int64_t k0_, n_;
  __device__ __forceinline__ void init(const Params& p, int64_t v) {
    const int64_t c = v / p.P;
    k0_ = (v % p.P) * p.L;
    n_ = seg_steps(p.Tn, k0_, p.L);
    if (k0_ == 0) {
      T L[DD];
      load_vec<T, D>(mu, p.mu0 + c * D);
      load_vec<T, DD>(L, p.chol_p0 + c * DD);
      llt<T, D>(P, L);
    } else {  // seed = state at step k0 - 1, parked in this segment's last output slots
#pragma unroll
      for (int i = 0; i < D; ++i) mu[i] = T(0);
      const int64_t kl = k0_ + seg_steps(p.Tn, k0_, p.L) - 1;
      if (kl >= k0_) {
        const int64_t e = mom_park(p, c, v % p.P, kl, 0);
        if (p.pk_vec) load_vec_rw<T, D>(mu, p.pk_vec + e * D);
        load_vec_rw<T, DD>(P, p.pk_diag + e * DD);
      }
    }
  }
  __device__ __forceinline__ void tile(const Params& p, const T* const* in, T* const* out,
                                       int64_t j0, int ns) {
    if (n_ - j0 < ns) ns = (int)(n_ - j0);  // ragged last segment
    for (int j = 0; j < ns; ++j) {
      if (k0_ + j0 + j > 0) {
        T A[DD], off[D], L[DD], AP[DD], E[DD];
        ld_s<T, DD>(A, in[0] + j * DD);
        ld_s<T, D>(off, in[1] + j * D);
        ld_s<T, DD>(L, in[2] + j * DD);
        gemm<T, D>(AP, A, P);
        gemv_add<T, D>(off, A, mu);  // mu_{k+1}
        if (EXPECT) {
#pragma unroll
          for (int r = 0; r < D; ++r)
#pragma unroll
            for (int q = 0; q < D; ++q) E[r * D + q] = Num<T>::fma(off[r], mu[q], AP[r * D + q]);
          st_s<T, DD>(out[2] + j * DD, E);
        } else {
          st_s<T, DD>(out[2] + j * DD, AP);
        }
        llt<T, D>(P, L);
#pragma unroll
        for (int r = 0; r < D; ++r)
#pragma unroll
          for (int q = 0; q <= r; ++q) {
            T v = P[r * D + q];
#pragma unroll
            for (int s = 0; s < D; ++s) v = Num<T>::fma(AP[r * D + s], A[q * D + s], v);
            P[r * D + q] = v;
            P[q * D + r] = v;
          }
#pragma unroll
        for (int r = 0; r < D; ++r) mu[r] = off[r];
      }
      st_s<T, D>(out[0] + j * D, mu);
      if (EXPECT) {
        T E[DD];
#pragma unroll
        for (int r = 0; r < D; ++r)
#pragma unroll
          for (int q = 0; q < D; ++q) E[r * D + q] = Num<T>::fma(mu[r], mu[q], P[r * D + q]);
        st_s<T, DD>(out[1] + j * DD, E);
      } else {
        st_s<T, DD>(out[1] + j * DD, P);
      }
    }
  }
  __device__ __forceinline__ void finish(const Params&, int64_t, bool) {}
};

// pass 1: the composed transition of every segment but the last of each chain
template <typename T_, int D>
struct SsmMomSummaryCore {
  using T = T_;
  using Params = SsmMomentsParams<T>;
  static constexpr int DD = D * D;
  static constexpr int NIN = 3, NOUT = 0;
  static constexpr bool BACKWARD = false;
  static constexpr int ein(int i) { return i == 1 ? D : DD; }
  static constexpr int eout(int) { return 1; }
  static __device__ __forceinline__ int64_t num_chains(const Params& p) { return p.B * p.P; }
  static __device__ __forceinline__ int64_t max_steps(const Params& p) { return p.L; }
  static __device__ __forceinline__ StreamGeom in_geom(const Params& p, int i, int64_t v) {
    const int64_t c = v / p.P, seg = v % p.P, k0 = seg * p.L;
    // the last segment has no successor: nothing to summarise
    const int64_t n = seg + 1 < p.P ? seg_steps(p.Tn, k0, p.L) : 0;
    return vgeom_incoming<T>(i == 0 ? p.a : (i == 1 ? p.b : p.chol_q), c, p.Tn, ein(i), k0, n);
  }
  static __device__ __forceinline__ StreamGeom out_geom(const Params&, int, int64_t) {
    return StreamGeom{nullptr, 0, 0};
  }
  static void tm_describe(const Params& p, TmStream* in, TmStream*) {
    in[0] = TmStream{p.a, p.Tn - 1, -1};
    in[1] = TmStream{p.b, p.Tn - 1, -1};
    in[2] = TmStream{p.chol_q, p.Tn - 1, -1};
  }
  static int64_t tm_segments(const Params& p) { return p.P; }
  static int64_t tm_seg_len(const Params& p) { return p.L; }
  static int64_t tm_chains(const Params& p) { return p.B; }
  T Phi[DD], cv[D], Qt[DD];
  int64_t k0_;
  bool live_;
  __device__ __forceinline__ void init(const Params& p, int64_t v) {
    k0_ = (v % p.P) * p.L;
    live_ = (v % p.P) + 1 < p.P;
#pragma unroll
    for (int i = 0; i < DD; ++i) {
      Phi[i] = (i / D == i % D) ? T(1) : T(0);
      Qt[i] = T(0);
    }
#pragma unroll
    for (int i = 0; i < D; ++i) cv[i] = T(0);
  }
  __device__ __forceinline__ void tile(const Params&, const T* const* in, T* const*, int64_t j0, int ns) {
    if (!live_) return;
    for (int j = 0; j < ns; ++j) {
      if (k0_ + j0 + j == 0) continue;
      T A[DD], off[D], L[DD], AP[DD];
      ld_s<T, DD>(A, in[0] + j * DD);
      ld_s<T, D>(off, in[1] + j * D);
      ld_s<T, DD>(L, in[2] + j * DD);
      gemv_add<T, D>(off, A, cv);
#pragma unroll
      for (int r = 0; r < D; ++r) cv[r] = off[r];
      gemm<T, D>(AP, A, Phi);
#pragma unroll
      for (int i = 0; i < DD; ++i) Phi[i] = AP[i];
      gemm<T, D>(AP, A, Qt);
      llt<T, D>(Qt, L);
#pragma unroll
      for (int r = 0; r < D; ++r)
#pragma unroll
        for (int q = 0; q <= r; ++q) {
          T v = Qt[r * D + q];
#pragma unroll
          for (int s = 0; s < D; ++s) v = Num<T>::fma(AP[r * D + s], A[q * D + s], v);
          Qt[r * D + q] = v;
          Qt[q * D + r] = v;
        }
    }
  }
  __device__ __forceinline__ void finish(const Params& p, int64_t v, bool valid) {
    if (!valid || !live_) return;
    const int64_t c = v / p.P;
    const int64_t kl = k0_ + p.L - 1;  // a live segment is complete: L steps
    const int64_t e0 = mom_park(p, c, v % p.P, kl, 0), e1 = mom_park(p, c, v % p.P, kl, 1);
    store_vec<T, DD>(p.pk_diag + e0 * DD, Phi);
    store_vec<T, DD>(p.pk_diag + e1 * DD, Qt);
    if (p.pk_vec) store_vec<T, D>(p.pk_vec + e0 * D, cv);
  }
};

// ---- pass 2: state before every segment -------------------------------------------------------
// element of a range of steps: mu_end = Phi mu + c,  Sigma_end = Phi Sigma Phi^T + Qt
template <typename T, int D>
struct MomElem {
  T Phi[D * D], c[D], Qt[D * D];
  __device__ __forceinline__ void identity() {
#pragma unroll
    for (int i = 0; i < D * D; ++i) {
      Phi[i] = (i / D == i % D) ? T(1) : T(0);
      Qt[i] = T(0);
    }
#pragma unroll
    for (int i = 0; i < D; ++i) c[i] = T(0);
  }
  // Sigma <- M Sigma M^T + Q (symmetric result)
  static __device__ __forceinline__ void congruence(T* __restrict__ sig, const T* __restrict__ M,
                                                    const T* __restrict__ Q) {
    T MS[D * D];
    gemm<T, D>(MS, M, sig);
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
      for (int q = 0; q <= r; ++q) {
        T v = Q[r * D + q];
#pragma unroll
        for (int s = 0; s < D; ++s) v = Num<T>::fma(MS[r * D + s], M[q * D + s], v);
        sig[r * D + q] = v;
        sig[q * D + r] = v;
      }
  }
  // this <- later o this  (apply `this` first, then `later`)
  __device__ __forceinline__ void then(const MomElem& later) {
    T t[D * D], v[D];
    gemm<T, D>(t, later.Phi, Phi);
#pragma unroll
    for (int i = 0; i < D * D; ++i) Phi[i] = t[i];
#pragma unroll
    for (int i = 0; i < D; ++i) v[i] = later.c[i];
    gemv_add<T, D>(v, later.Phi, c);
#pragma unroll
    for (int i = 0; i < D; ++i) c[i] = v[i];
    congruence(Qt, later.Phi, later.Qt);
  }
  __device__ __forceinline__ void apply(T* __restrict__ mu, T* __restrict__ sig) const {
    T v[D];
#pragma unroll
    for (int i = 0; i < D; ++i) v[i] = c[i];
    gemv_add<T, D>(v, Phi, mu);
#pragma unroll
    for (int i = 0; i < D; ++i) mu[i] = v[i];
    congruence(sig, Phi, Qt);
  }
  __device__ __forceinline__ void load(const SsmMomentsParams<T>& p, int64_t c_, int64_t seg, int64_t kl) {
    const int64_t e0 = mom_park(p, c_, seg, kl, 0), e1 = mom_park(p, c_, seg, kl, 1);
    load_vec_rw<T, D * D>(Phi, p.pk_diag + e0 * D * D);
    load_vec_rw<T, D * D>(Qt, p.pk_diag + e1 * D * D);
    if (p.pk_vec) load_vec_rw<T, D>(c, p.pk_vec + e0 * D);
    else {
#pragma unroll
      for (int i = 0; i < D; ++i) c[i] = T(0);
    }
  }
  __device__ __forceinline__ void shfl_up_from(const MomElem& src, int delta) {
#pragma unroll
    for (int i = 0; i < D * D; ++i) {
      Phi[i] = __shfl_up_sync(0xffffffffu, src.Phi[i], delta);
      Qt[i] = __shfl_up_sync(0xffffffffu, src.Qt[i], delta);
    }
#pragma unroll
    for (int i = 0; i < D; ++i) c[i] = __shfl_up_sync(0xffffffffu, src.c[i], delta);
  }
};

// One WARP per chain: lane l owns segments [l*m, (l+1)*m); it composes their elements, the warp
// scans the 32 composites, and every lane then walks its segments from its prefix state, parking
// the state before every segment s >= 1 in that segment's last output slots (which held the
// segment's own element: loaded before it is overwritten).  One thread per chain when P is small.
template <typename T, int D, bool WARP>
__global__ void __launch_bounds__(128)
ssm_moments_seed_kernel(const SsmMomentsParams<T> p) {
  constexpr int DD = D * D;
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t c = WARP ? tid / 32 : tid;
  const int lane = WARP ? (int)(tid & 31) : 0;
  if (c >= p.B) return;
  const int64_t nlive = p.P - 1;  // segments that have a successor
  const int64_t m = WARP ? (nlive + 31) / 32 : p.P;
  const int64_t s0 = lane * m;
  int64_t s1 = s0 + m;
  if (s1 > p.P) s1 = p.P;
  T mu[D], sig[DD], L[DD];
  load_vec<T, D>(mu, p.mu0 + c * D);
  load_vec<T, DD>(L, p.chol_p0 + c * DD);
  llt<T, D>(sig, L);
  if (WARP) {
    MomElem<T, D> mine, e, other;
    mine.identity();
    for (int64_t seg = s0; seg < s1 && seg < nlive; ++seg) {
      e.load(p, c, seg, seg * p.L + p.L - 1);
      mine.then(e);
    }
#pragma unroll 1
    for (int delta = 1; delta < 32; delta <<= 1) {
      other.shfl_up_from(mine, delta);
      if (lane >= delta) {
        other.then(mine);
        mine = other;
      }
    }
    // exclusive prefix: the inclusive composite of the previous lane
    other.shfl_up_from(mine, 1);
    if (lane > 0) other.apply(mu, sig);
    if (lane == 31) s1 = p.P;  // the last lane also parks the seed of the final segment
  }
  for (int64_t seg = s0; seg < s1; ++seg) {
    const int64_t k0 = seg * p.L;
    const int64_t n = seg_steps(p.Tn, k0, p.L);
    if (n <= 0) break;
    const int64_t kl = k0 + n - 1;
    MomElem<T, D> e;
    const bool live = seg < nlive;
    if (live) e.load(p, c, seg, kl);
    if (seg > 0) {
      const int64_t e0 = mom_park(p, c, seg, kl, 0);
      if (p.pk_vec) store_vec<T, D>(p.pk_vec + e0 * D, mu);
      store_vec<T, DD>(p.pk_diag + e0 * DD, sig);
    }
    if (live) e.apply(mu, sig);
  }
}

// ---------------------------------------------------------------------------------------------
// Element of a range of steps of a VECTOR-affine recursion x_out = Phi x_in + c (triangular solves,
// means, samples).  Phi is stored column by column (column q at Phi[q*D ..]), as the summary passes
// produce it.  (e1 then e2): Phi = Phi2 Phi1, c = Phi2 c1 + c2.
template <typename T, int D>
struct AffElem {
  T Phi[D * D], c[D];
  int empty;
  __device__ __forceinline__ void clear() {
    empty = 1;
#pragma unroll
    for (int i = 0; i < D * D; ++i) Phi[i] = (i / D == i % D) ? T(1) : T(0);
#pragma unroll
    for (int i = 0; i < D; ++i) c[i] = T(0);
  }
  __device__ __forceinline__ void then(const AffElem& e2) {
    if (e2.empty) return;
    if (empty) {
      *this = e2;
      return;
    }
    T Pn[D * D], cn[D];
#pragma unroll
    for (int q = 0; q < D; ++q)  // column q of Phi2 Phi1 = Phi2 (column q of Phi1)
#pragma unroll
      for (int i = 0; i < D; ++i) {
        T v = T(0);
#pragma unroll
        for (int s2 = 0; s2 < D; ++s2) v = Num<T>::fma(e2.Phi[s2 * D + i], Phi[q * D + s2], v);
        Pn[q * D + i] = v;
      }
#pragma unroll
    for (int i = 0; i < D; ++i) {
      T v = e2.c[i];
#pragma unroll
      for (int s2 = 0; s2 < D; ++s2) v = Num<T>::fma(e2.Phi[s2 * D + i], c[s2], v);
      cn[i] = v;
    }
#pragma unroll
    for (int i = 0; i < D * D; ++i) Phi[i] = Pn[i];
#pragma unroll
    for (int i = 0; i < D; ++i) c[i] = cn[i];
  }
  // x <- Phi x + c
  __device__ __forceinline__ void apply(T* __restrict__ x) const {
    T y[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
      T v = c[i];
#pragma unroll
      for (int q = 0; q < D; ++q) v = Num<T>::fma(Phi[q * D + i], x[q], v);
      y[i] = v;
    }
#pragma unroll
    for (int i = 0; i < D; ++i) x[i] = y[i];
  }
  __device__ __forceinline__ void shfl_up_from(const AffElem& src, int delta) {
#pragma unroll
    for (int i = 0; i < D * D; ++i) Phi[i] = __shfl_up_sync(0xffffffffu, src.Phi[i], delta);
#pragma unroll
    for (int i = 0; i < D; ++i) c[i] = __shfl_up_sync(0xffffffffu, src.c[i], delta);
    empty = __shfl_up_sync(0xffffffffu, src.empty, delta);
  }
};

// Fold of the elements of a vector-affine sweep over `out` [n,T,D] in sweep order it = 0..P-1
// (segment it, or P-1-it when BACKWARD).  slot(k0, n, i) is the output step that holds parked vector
// i (0: c | seed, 1..D: columns of Phi) of segment [k0, k0+n).  The first segment of the sweep starts
// from a known state (x0; its parked c already is the state at its end), so its Phi is ignored.
// WARP: one warp per chain, lane l owns sweep positions [l*m, (l+1)*m).
template <typename T, int D, bool BACKWARD, bool WARP, class Slot>
__device__ __forceinline__ void affine_fold(T* out, int64_t nchains, int64_t Tn, int64_t P, int64_t L,
                                            Slot slot) {
  using Elem = AffElem<T, D>;
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t c = WARP ? tid / 32 : tid;
  const int lane = WARP ? (int)(tid & 31) : 0;
  if (c >= nchains) return;
  const int64_t m = WARP ? (P + 31) / 32 : P;
  const int64_t i0 = lane * m;
  int64_t i1 = i0 + m;
  if (i1 > P) i1 = P;
  T* o = out + c * Tn * D;
  auto load = [&](Elem& e, int64_t it) {
    const int64_t seg = BACKWARD ? P - 1 - it : it;
    const int64_t k0 = seg * L, n = seg_steps(Tn, k0, L);
    e.empty = 0;
    load_vec_rw<T, D>(e.c, o + slot(k0, n, 0) * D);
#pragma unroll
    for (int q = 0; q < D; ++q) load_vec_rw<T, D>(e.Phi + q * D, o + slot(k0, n, q + 1) * D);
    if (it == 0) {  // starts from the known initial state: x_out = c whatever enters
#pragma unroll
      for (int i = 0; i < D * D; ++i) e.Phi[i] = T(0);
    }
  };
  Elem X;
  X.clear();
  if (WARP) {
    Elem e, other;
    for (int64_t it = i0; it < i1 && it < P - 1; ++it) {  // the last segment of the sweep feeds nobody
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
    const int64_t seg = BACKWARD ? P - 1 - it : it;
    const int64_t k0 = seg * L, n = seg_steps(Tn, k0, L);
    Elem e;
    e.clear();
    const bool live = it + 1 < P;
    if (live) load(e, it);  // before its slot receives the seed
    if (it > 0) store_vec<T, D>(o + slot(k0, n, 0) * D, X.c);  // state entering = c of the prefix (Phi x0 term is 0)
    if (!live) break;
    X.then(e);
  }
}

// ---------------------------------------------------------------------------------------------
// x_k = A_k x_{k-1} + b_k (+ chol_q eps): affine in x, so few long chains run parallel in time: the
// SUMMARY pass composes every segment into (Phi, c) by running the recursion on c and on the columns
// of Phi, ssm_affine_seed_kernel folds them per chain, and the segments restart from their seeds.
// The D+1 vectors of an element / the seed are parked in the segment's LAST output steps.
template <typename T>
struct SsmAffineParams {
  const T *mu0, *chol_p0, *a, *b, *chol_q, *eps;
  T* out;
  int64_t n, Bm, Tn;
  int64_t P, L;
  unsigned long long seed;  // RNG cores: the standard normals are drawn in the kernel (smallrng.cuh)
};

// NOISE: x_k += chol_q eps_k; RNG: eps is drawn in the kernel instead of being streamed in
template <typename T_, int D, bool NOISE, bool SUMMARY = false, bool RNG = false>
struct SsmAffineCore {
  using T = T_;
  using Params = SsmAffineParams<T>;
  static constexpr int DD = D * D;
  static constexpr int NIN = NOISE ? (RNG ? 3 : 4) : 2, NOUT = SUMMARY ? 0 : 1;
  static constexpr bool BACKWARD = false;
  static constexpr int ein(int i) { return (i == 0 || i == 2) ? DD : D; }
  static constexpr int eout(int) { return D; }
  static __device__ __forceinline__ int64_t num_chains(const Params& p) { return p.n * p.P; }
  static __device__ __forceinline__ int64_t max_steps(const Params& p) { return p.L; }
  static __device__ __forceinline__ bool is_live(const Params& p, int64_t v) {
    return ((v % p.P) + 1) * p.L < p.Tn;  // complete segment with a successor
  }
  static __device__ __forceinline__ StreamGeom in_geom(const Params& p, int i, int64_t v) {
    const int64_t c = v / p.P, k0 = (v % p.P) * p.L;
    int64_t n = seg_steps(p.Tn, k0, p.L);
    if (SUMMARY && !is_live(p, v)) n = 0;
    if (i == 3) return vgeom_states<T>(p.eps, c, p.Tn, D, k0, n);
    return vgeom_incoming<T>(i == 0 ? p.a : (i == 1 ? p.b : p.chol_q), c % p.Bm, p.Tn, ein(i), k0, n);
  }
  static __device__ __forceinline__ StreamGeom out_geom(const Params& p, int, int64_t v) {
    if (SUMMARY) return StreamGeom{nullptr, 0, 0};
    const int64_t c = v / p.P, k0 = (v % p.P) * p.L;
    return vgeom_states<T>(p.out, c, p.Tn, D, k0, seg_steps(p.Tn, k0, p.L));
  }
  T x[D];
  T Phi[SUMMARY ? DD : 1];  // column q at Phi[q * D ..]
  int64_t cm, k0_, n_;
  bool live_;
  ChainRng rng;
  __device__ __forceinline__ void init(const Params& p, int64_t v) {
    const int64_t c = v / p.P;
    cm = c % p.Bm;
    k0_ = (v % p.P) * p.L;
    n_ = seg_steps(p.Tn, k0_, p.L);
    live_ = !SUMMARY || is_live(p, v);
    if (RNG) rng.init(p.seed, c, k0_, D);
    if (SUMMARY) {
#pragma unroll
      for (int i = 0; i < DD; ++i) Phi[SUMMARY ? i : 0] = (i / D == i % D) ? T(1) : T(0);
    }
    if (k0_ == 0) {
      load_vec<T, D>(x, p.mu0 + cm * D);
    } else if (SUMMARY) {
#pragma unroll
      for (int i = 0; i < D; ++i) x[i] = T(0);
    } else if (n_ > 0) {
      load_vec_rw<T, D>(x, p.out + (c * p.Tn + k0_ + n_ - 1) * D);  // seed: x_{k0-1}
    }
  }
  __device__ __forceinline__ void tile(const Params& p, const T* const* in, T* const* out,
                                       int64_t j0, int ns) {
    if (!live_) return;
    if (n_ - j0 < ns) ns = (int)(n_ - j0);
    for (int j = 0; j < ns; ++j) {
      if (k0_ + j0 + j > 0) {
        T A[DD], off[D];
        ld_s<T, DD>(A, in[0] + j * DD);
        ld_s<T, D>(off, in[1] + j * D);
        if (NOISE) {
          T L[DD], e[D];
          ld_s<T, DD>(L, in[2] + j * DD);
          zero_upper<T, D>(L);
          if (RNG) rng.template draw<T, D>(e);
          else ld_s<T, D>(e, in[RNG ? 0 : 3] + j * D);
          gemv_add<T, D>(off, L, e);
        }
        gemv_add<T, D>(off, A, x);
#pragma unroll
        for (int r = 0; r < D; ++r) x[r] = off[r];
        if (SUMMARY) {
          T t[DD];
          // columns of Phi <- A (columns of Phi): Phi is stored column by column, i.e. as Phi^T
#pragma unroll
          for (int q = 0; q < D; ++q)
#pragma unroll
            for (int r = 0; r < D; ++r) {
              T v = T(0);
#pragma unroll
              for (int s2 = 0; s2 < D; ++s2) v = Num<T>::fma(A[r * D + s2], Phi[SUMMARY ? q * D + s2 : 0], v);
              t[q * D + r] = v;
            }
#pragma unroll
          for (int i = 0; i < DD; ++i) Phi[SUMMARY ? i : 0] = t[i];
        }
      } else {
        if (NOISE) {
          T L[DD], e[D];
          load_vec<T, DD>(L, p.chol_p0 + cm * DD);
          zero_upper<T, D>(L);
          if (RNG) rng.template draw<T, D>(e);
          else ld_s<T, D>(e, in[RNG ? 0 : 3] + j * D);
          gemv_add<T, D>(x, L, e);
        }
      }
      if (!SUMMARY) st_s<T, D>(out[0] + j * D, x);
    }
  }
  __device__ __forceinline__ void finish(const Params& p, int64_t v, bool valid) {
    if (!SUMMARY || !valid || !live_) return;
    const int64_t c = v / p.P;
    const int64_t kl = k0_ + n_ - 1;
    store_vec<T, D>(p.out + (c * p.Tn + kl) * D, x);
#pragma unroll
    for (int q = 0; q < D; ++q)
      store_vec<T, D>(p.out + (c * p.Tn + kl - 1 - q) * D, Phi + (SUMMARY ? q * D : 0));
  }
};

// fold of the segment elements of one chain (affine_fold): parks x_{k0-1} in the seed slot (the last
// output step) of every segment s >= 1; the vectors of an element sit in the last D+1 steps.
template <typename T, int D, bool WARP>
__global__ void __launch_bounds__(128)
ssm_affine_seed_kernel(const SsmAffineParams<T> p) {
  affine_fold<T, D, false, WARP>(p.out, p.n, p.Tn, p.P, p.L,
                                 [](int64_t k0, int64_t n, int i) { return k0 + n - 1 - i; });
}

// ---------------------------------------------------------------------------------------------
// KL(q || p), chain-rule form.  q's marginal (mu_k, Sigma_k) is the only state carried along the
// chain, so few long chains are evaluated parallel in time with the moment elements above: summary
// pass over q's transitions -> seeds -> every segment accumulates its share of the sum
// (SsmKlParams::partial), added in segment order by ssm_kl_reduce_kernel.  KL has no large outputs
// to park elements in: they live in a compact workspace (seed_vec / seed_diag).
template <typename T>
struct SsmKlParams {
  const T *q_mu0, *q_chol_p0, *q_a, *q_b, *q_chol_q;
  const T *p_mu0, *p_chol_p0, *p_a, *p_b, *p_chol_q;
  T* out;
  int64_t B, Tn;
  int64_t P, L;
  const T *seed_vec, *seed_diag;  // [B*P*2, D], [B*P*2, D*D]: entry (c*P + seg)*2 = state before seg
  T* partial;                     // [B*P]
};

template <typename T_, int D>
struct SsmKlCore {
  using T = T_;
  using Params = SsmKlParams<T>;
  static constexpr int DD = D * D;
  static constexpr int NIN = 6, NOUT = 0;
  static constexpr bool BACKWARD = false;
  static constexpr int ein(int i) { return (i % 3 == 1) ? D : DD; }
  static constexpr int eout(int) { return 1; }
  static __device__ __forceinline__ int64_t num_chains(const Params& p) { return p.B * p.P; }
  static __device__ __forceinline__ int64_t max_steps(const Params& p) { return p.L; }
  static __device__ __forceinline__ StreamGeom in_geom(const Params& p, int i, int64_t v) {
    const T* base = i == 0 ? p.q_a : i == 1 ? p.q_b : i == 2 ? p.q_chol_q
                  : i == 3 ? p.p_a : i == 4 ? p.p_b : p.p_chol_q;
    const int64_t c = v / p.P, k0 = (v % p.P) * p.L;
    return vgeom_incoming<T>(base, c, p.Tn, ein(i), k0, seg_steps(p.Tn, k0, p.L));
  }
  static __device__ __forceinline__ StreamGeom out_geom(const Params&, int, int64_t) {
    return StreamGeom{nullptr, 0, 0};
  }
  T mu[D], P[DD], kl;
  LogProd<T> ratio;
  int64_t k0_, n_;
  __device__ __forceinline__ void init(const Params& p, int64_t v) {
    const int64_t c = v / p.P;
    k0_ = (v % p.P) * p.L;
    n_ = seg_steps(p.Tn, k0_, p.L);
    ratio.init();
    kl = T(0);
    if (k0_ == 0) {
      T Lq[DD], Lp[DD], dm[D];
      load_vec<T, D>(mu, p.q_mu0 + c * D);
      load_vec<T, DD>(Lq, p.q_chol_p0 + c * DD);
      load_vec<T, DD>(Lp, p.p_chol_p0 + c * DD);
      load_vec<T, D>(dm, p.p_mu0 + c * D);
#pragma unroll
      for (int i = 0; i < D; ++i) dm[i] = mu[i] - dm[i];
      kl = kl_gauss_term<T, D>(Lp, Lq, nullptr, dm, nullptr, ratio);
      llt<T, D>(P, Lq);
    } else {
      load_vec_rw<T, D>(mu, p.seed_vec + v * 2 * D);
      load_vec_rw<T, DD>(P, p.seed_diag + v * 2 * DD);
    }
  }
  __device__ __forceinline__ void tile(const Params&, const T* const* in, T* const*, int64_t j0,
                                       int ns) {
    if (n_ - j0 < ns) ns = (int)(n_ - j0);
    for (int j = 0; j < ns; ++j) {
      if (k0_ + j0 + j == 0) continue;
      T Aq[DD], Ap[DD], bq[D], dm[D], Lq[DD], Lp[DD], AP[DD];
      ld_s<T, DD>(Aq, in[0] + j * DD);
      ld_s<T, D>(bq, in[1] + j * D);
      ld_s<T, DD>(Lq, in[2] + j * DD);
      ld_s<T, DD>(Ap, in[3] + j * DD);
      ld_s<T, D>(dm, in[4] + j * D);
      ld_s<T, DD>(Lp, in[5] + j * DD);
#pragma unroll
      for (int i = 0; i < DD; ++i) Ap[i] = Aq[i] - Ap[i];  // dA
#pragma unroll
      for (int i = 0; i < D; ++i) dm[i] = bq[i] - dm[i];
      gemv_add<T, D>(dm, Ap, mu);
      kl += kl_gauss_term<T, D>(Lp, Lq, Ap, dm, P, ratio);
      gemm<T, D>(AP, Aq, P);
      llt<T, D>(P, Lq);
#pragma unroll
      for (int r = 0; r < D; ++r)
#pragma unroll
        for (int q = 0; q <= r; ++q) {
          T v = P[r * D + q];
#pragma unroll
          for (int s = 0; s < D; ++s) v = Num<T>::fma(AP[r * D + s], Aq[q * D + s], v);
          P[r * D + q] = v;
          P[q * D + r] = v;
        }
      gemv_add<T, D>(bq, Aq, mu);
#pragma unroll
      for (int i = 0; i < D; ++i) mu[i] = bq[i];
    }
  }
  __device__ __forceinline__ void finish(const Params& p, int64_t v, bool valid) {
    if (!valid) return;
    const T val = kl + ratio.log_abs();
    if (p.P == 1) p.out[v] = val;
    else p.partial[v] = val;
  }
};

template <typename T>
__global__ void __launch_bounds__(128)
ssm_kl_reduce_kernel(const T* __restrict__ partial, T* __restrict__ out, int64_t B, int64_t P) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= B) return;
  T s = T(0);
  for (int64_t seg = 0; seg < P; ++seg) s += partial[c * P + seg];
  out[c] = s;
}

// ---------------------------------------------------------------------------------------------
// Element of a range of steps of a linear-fractional recursion (block Cholesky forward, U D U^T /
// naturals -> SSM backward):  S_out = P - Q (S_in + R)^{-1} Q^T,  r_out = p + Q (S_in + R)^{-1} (r_in + r).
// A chain's first range (in sweep order) has no incoming block, which behaves like
// S_in = infinity: (S_in + R)^{-1} = 0, so the state after ranges 0..s-1 is simply (P, p) of their
// COMBINED element -- the fold needs the combination only, never an application to a state:
//   (e1 then e2):  S = P1 + R2 = C C^T,  Y = S^{-1} Q1,  u = S^{-1} (p1 + r2)
//                  P = P2 - Q2 S^{-1} Q2^T,  Q = Q2 Y,  R = R1 - Q1^T Y,  p = p2 + Q2 u,  r = r1 + Q1^T u
template <typename T, int D, bool RHS>
struct LftElem {
  static constexpr int DD = D * D;
  T P[DD], Q[DD], R[DD], p[D], r[D];
  int empty;  // 1: identity (nothing combined yet)
  __device__ __forceinline__ void clear() {
    empty = 1;
#pragma unroll
    for (int i = 0; i < DD; ++i) P[i] = Q[i] = R[i] = T(0);
#pragma unroll
    for (int i = 0; i < D; ++i) p[i] = r[i] = T(0);
  }
  // this <- (this then later); returns false if S = P1 + R2 is not positive definite
  __device__ __forceinline__ bool then(const LftElem& e2) {
    if (e2.empty) return true;
    if (empty) {
      *this = e2;
      return true;
    }
    T S[DD], rinv[D], Y[DD], Z[DD], u[D];
#pragma unroll
    for (int i = 0; i < DD; ++i) S[i] = P[i] + e2.R[i];
    const bool ok = chol_lower<T, D>(S, rinv);
#pragma unroll
    for (int i = 0; i < DD; ++i) Y[i] = Q[i];
    trsm_left_lower<T, D>(S, rinv, Y);
    trsm_left_lower_t<T, D>(S, rinv, Y);  // Y = S^{-1} Q1
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
      for (int b = 0; b < D; ++b) Z[a * D + b] = e2.Q[b * D + a];
    trsm_left_lower<T, D>(S, rinv, Z);
    trsm_left_lower_t<T, D>(S, rinv, Z);  // Z = S^{-1} Q2^T
#pragma unroll
    for (int i = 0; i < D; ++i) u[i] = p[i] + e2.r[i];
    trsv_lower<T, D>(S, rinv, u);
    trsv_lower_t<T, D>(S, rinv, u);  // u = S^{-1} (p1 + r2)
    T Pn[DD], Qn[DD], Rn[DD], pn[D], rn[D];
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
      for (int b = 0; b < D; ++b) {
        T vp = e2.P[a * D + b], vq = T(0), vr = R[a * D + b];
#pragma unroll
        for (int s = 0; s < D; ++s) {
          vp = Num<T>::fma(-e2.Q[a * D + s], Z[s * D + b], vp);
          vq = Num<T>::fma(e2.Q[a * D + s], Y[s * D + b], vq);
          vr = Num<T>::fma(-Q[s * D + a], Y[s * D + b], vr);
        }
        Pn[a * D + b] = vp;
        Qn[a * D + b] = vq;
        Rn[a * D + b] = vr;
      }
#pragma unroll
    for (int a = 0; a < D; ++a) {
      T vp = e2.p[a], vr = r[a];
#pragma unroll
      for (int s = 0; s < D; ++s) {
        vp = Num<T>::fma(e2.Q[a * D + s], u[s], vp);
        vr = Num<T>::fma(Q[s * D + a], u[s], vr);
      }
      pn[a] = vp;
      rn[a] = vr;
    }
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
      for (int b = 0; b <= a; ++b) {  // P and R are symmetric: keep them exactly so
        P[a * D + b] = P[b * D + a] = Pn[a * D + b];
        R[a * D + b] = R[b * D + a] = Rn[a * D + b];
      }
#pragma unroll
    for (int i = 0; i < DD; ++i) Q[i] = Qn[i];
#pragma unroll
    for (int i = 0; i < D; ++i) {
      p[i] = pn[i];
      r[i] = rn[i];
    }
    return ok;
  }
  __device__ __forceinline__ void shfl_up_from(const LftElem& src, int delta) {
#pragma unroll
    for (int i = 0; i < DD; ++i) {
      P[i] = __shfl_up_sync(0xffffffffu, src.P[i], delta);
      Q[i] = __shfl_up_sync(0xffffffffu, src.Q[i], delta);
      R[i] = __shfl_up_sync(0xffffffffu, src.R[i], delta);
    }
#pragma unroll
    for (int i = 0; i < D; ++i) {
      p[i] = RHS ? __shfl_up_sync(0xffffffffu, src.p[i], delta) : T(0);
      r[i] = RHS ? __shfl_up_sync(0xffffffffu, src.r[i], delta) : T(0);
    }
    empty = __shfl_up_sync(0xffffffffu, src.empty, delta);
  }
};

// ---------------------------------------------------------------------------------------------
// naturals -> SSM parameters: backward U D U^T sweep (nat_to_ssm_kernel in nat_kernels.cuh has the
// algebra)
//     D_k = Th_k - Ths_k^T D_{k+1}^{-1} Ths_k,   z_k = th_k + Ths_k^T D_{k+1}^{-1} z_{k+1}
// (Th = -2 theta_diag, Ths = theta_sub).  The map (D_{k+1}, z_{k+1}) -> (D_k, z_k) is linear-
// fractional and closed under composition in the form
//     D_out = P - Q (D_in + R)^{-1} Q^T,      z_out = p + Q (D_in + R)^{-1} (z_in + r),
// with the extension by one step  P' = Th - Ths^T P^{-1} Ths  (the recursion itself, started without
// an incoming block),  Q' = Ths^T P^{-1} Q,  R' = R - Q^T P^{-1} Q,  p' = th + Ths^T P^{-1} p,
// r' = r + Q^T P^{-1} p.  Few long chains are therefore evaluated parallel in time, exactly:
//   1. NatSummaryCore : every segment (but the first) reduces its steps to one element (P,Q,R,p,r)
//   2. nat_seed_kernel: per chain, fold the elements from the last segment down -> (D, z) entering
//                       every segment
//   3. NatToSsmCore   : every segment runs the ordinary sweep from its seed.
// No workspace: elements and seeds are parked in the output slots of each segment's FIRST two steps
// (the last ones a backward sweep writes):  P | seed D -> out_chol[k0], R -> out_chol[k0+1],
// Q -> out_a[k0], p | seed z -> out_off[k0], r -> out_off[k0+1];  L >= 2.
template <typename T>
struct NatToSsmParams {
  const T *th_lin, *th_diag, *th_sub;
  T *out_a, *out_off, *out_chol;
  int32_t* info;
  int64_t B, Tn;
  int64_t P, L;  // segments per chain, steps per segment (P == 1: L == Tn)
};

// transition LEAVING local step j of segment (c, k0): entry k0 + j of a [B,T-1,...] stream
template <typename T>
__device__ __forceinline__ StreamGeom vgeom_outgoing(const T* base, int64_t c, int64_t Tn, int E,
                                                     int64_t k0, int64_t steps) {
  StreamGeom g;
  g.step0 = base ? byte_ptr(base) + (c * (Tn - 1) + k0) * (int64_t)(E * sizeof(T)) : nullptr;
  g.first = 0;
  int64_t n = Tn - 1 - k0;
  g.end = n < steps ? (n < 0 ? 0 : n) : steps;
  return g;
}

template <typename T_, int D>
struct NatGeomBase {
  using T = T_;
  using Params = NatToSsmParams<T>;
  static constexpr int DD = D * D;
  static constexpr bool BACKWARD = true;
  static __device__ __forceinline__ int64_t num_chains(const Params& p) { return p.B * p.P; }
  static __device__ __forceinline__ int64_t max_steps(const Params& p) { return p.L; }
  static constexpr int ein(int i) { return i == 0 ? D : DD; }
  static __device__ __forceinline__ StreamGeom in_geom(const Params& p, int i, int64_t v) {
    const int64_t c = v / p.P, k0 = (v % p.P) * p.L;
    const int64_t n = seg_steps(p.Tn, k0, p.L);
    if (i == 2) return vgeom_outgoing<T>(p.th_sub, c, p.Tn, DD, k0, n);
    return vgeom_states<T>(i == 0 ? p.th_lin : p.th_diag, c, p.Tn, ein(i), k0, n);
  }
  // host-side description for the tensor-map engine: inputs here, outputs (NatToSsmCore only) below
  static void tm_describe(const Params& p, TmStream* in, TmStream* out) {
    in[0] = TmStream{p.th_lin, p.Tn, 0};
    in[1] = TmStream{p.th_diag, p.Tn, 0};
    in[2] = TmStream{p.th_sub, p.Tn - 1, 0};
    if (out) {
      out[0] = TmStream{p.out_a, p.Tn - 1, 0};
      out[1] = TmStream{p.out_off, p.Tn, 0};
      out[2] = TmStream{p.out_chol, p.Tn, 0};
    }
  }
  static int64_t tm_segments(const Params& p) { return p.P; }
  static int64_t tm_seg_len(const Params& p) { return p.L; }
  static int64_t tm_chains(const Params& p) { return p.B; }
};

// Backward U D U^T sweep of one segment, seeded with the state entering it.
template <typename T_, int D>
struct NatToSsmCore : NatGeomBase<T_, D> {
  using T = T_;
  using Params = NatToSsmParams<T>;
  static constexpr int DD = D * D;
  static constexpr int NIN = 3, NOUT = 3;
  static constexpr int eout(int i) { return i == 1 ? D : DD; }
  // in-place stages (sweep_tmi.cuh): theta_lin -> offset, theta_diag -> chol Q, theta_sub -> A of the same step
  static constexpr int tm_alias(int i) { return i == 0 ? 1 : (i == 1 ? 2 : 0); }
  static __device__ __forceinline__ StreamGeom out_geom(const Params& p, int i, int64_t v) {
    const int64_t c = v / p.P, k0 = (v % p.P) * p.L;
    const int64_t n = seg_steps(p.Tn, k0, p.L);
    if (i == 0) return vgeom_outgoing<T>(p.out_a, c, p.Tn, DD, k0, n);
    return vgeom_states<T>(i == 1 ? p.out_off : p.out_chol, c, p.Tn, eout(i), k0, n);
  }
  T S[DD], rinv[D], z[D];
  T S2_[DD], r2_[D], z2_[D];
  int32_t fail;
  int64_t Tn_, k0_, n_;
  __device__ __forceinline__ void init(const Params& p, int64_t v) {
    fail = 0;
    Tn_ = p.Tn;
    const int64_t c = v / p.P;
    k0_ = (v % p.P) * p.L;
    n_ = seg_steps(p.Tn, k0_, p.L);
#pragma unroll
    for (int i = 0; i < D; ++i) z[i] = T(0);
    if (n_ > 0 && k0_ + n_ < p.Tn) {  // not the last segment: factor the seed D entering it
      load_vec_rw<T, DD>(S, p.out_chol + (c * p.Tn + k0_) * DD);
      load_vec_rw<T, D>(z, p.out_off + (c * p.Tn + k0_) * D);
      bool ok;
      if constexpr (D == 2) {
        ok = prep2(S[0], S[2], S[3], S, rinv);
      } else {
        ok = chol_lower<T, D>(S, rinv);
      }
      if (!ok) fail = (int32_t)(k0_ + n_ + 1);
    }
  }
  // ---- D = 2: closed forms instead of factor + substitutions -----------------------------------------------
  // For a 2 x 2 block D = [[d00, d10], [d10, d11]] everything a step needs follows from TWO INDEPENDENT
  // reciprocal square roots, ra = rsqrt(d11) and rb = rsqrt(det D):
  //     D^-1 = rb^2 adj(D),      chol(D^-1) = [[d11 ra rb, 0], [-d10 ra rb, ra]]
  // (check: c00^2 = d11 / det, c10 c00 = -d10 / det, c10^2 + c11^2 = d00 / det), against two dependent rsqrt's
  // for chol(D), two triangular solves per right-hand side, an explicit inverse and two more dependent rsqrt's
  // for its factor in the general path.  det = d00 d11 - d10^2 carries the same cancellation as the second
  // Cholesky pivot det / d00.  State: s = (d00, d10, d11, rb^2), r = (ra, rb).
  static __device__ __forceinline__ bool prep2(T d00, T d10, T d11, T* __restrict__ s, T* __restrict__ r) {
    const T det = Num<T>::fma(d00, d11, -(d10 * d10));
    r[0] = Num<T>::rsqrt(d11);
    r[1] = Num<T>::rsqrt(det);
    s[0] = d00;
    s[1] = d10;
    s[2] = d11;
    s[3] = r[1] * r[1];
    return (d00 > T(0)) && (det > T(0));
  }
  __device__ __forceinline__ void emit2(T* const* out, int j) {
    const T d00 = S[0], d10 = S[1], d11 = S[2], idet = S[3], ra = rinv[0], rb = rinv[1];
    T off[2], Qc[4];
    off[0] = idet * Num<T>::fma(d11, z[0], -(d10 * z[1]));
    off[1] = idet * Num<T>::fma(d00, z[1], -(d10 * z[0]));
    st_s<T, 2>(out[1] + j * 2, off);
    const T rab = ra * rb;
    Qc[0] = d11 * rab;
    Qc[1] = T(0);
    Qc[2] = -(d10 * rab);
    Qc[3] = ra;
    st_s<T, 4>(out[2] + j * 4, Qc);
  }
  __device__ __forceinline__ void advance2(const T* const* in, T* const* out, int j, int64_t k) {
    T th[2], Dk[4];
    ld_s<T, 2>(th, in[0] + j * 2);
    ld_s<T, 4>(Dk, in[1] + j * 4);
    T d00 = T(-2) * Dk[0], d10 = T(-2) * Dk[2], d11 = T(-2) * Dk[3];
    if (k + 1 < Tn_) {
      T Th[4], A[4];
      ld_s<T, 4>(Th, in[2] + j * 4);
      const T n00 = S[0], n10 = S[1], n11 = S[2], idet = S[3];
      // A_k = D_{k+1}^-1 theta_sub_k = rb^2 adj(D_{k+1}) theta_sub_k
      A[0] = idet * Num<T>::fma(n11, Th[0], -(n10 * Th[2]));
      A[1] = idet * Num<T>::fma(n11, Th[1], -(n10 * Th[3]));
      A[2] = idet * Num<T>::fma(n00, Th[2], -(n10 * Th[0]));
      A[3] = idet * Num<T>::fma(n00, Th[3], -(n10 * Th[1]));
      st_s<T, 4>(out[0] + j * 4, A);
      // D_k = -2 theta_diag_k - theta_sub_k^T A_k  (lower triangle)
      d00 = Num<T>::fma(-Th[2], A[2], Num<T>::fma(-Th[0], A[0], d00));
      d10 = Num<T>::fma(-Th[3], A[2], Num<T>::fma(-Th[1], A[0], d10));
      d11 = Num<T>::fma(-Th[3], A[3], Num<T>::fma(-Th[1], A[1], d11));
      // z_k = theta_lin_k + A_k^T z_{k+1}
      th[0] = Num<T>::fma(A[2], z[1], Num<T>::fma(A[0], z[0], th[0]));
      th[1] = Num<T>::fma(A[3], z[1], Num<T>::fma(A[1], z[0], th[1]));
    }
    const bool ok = prep2(d00, d10, d11, S2_, r2_);
    if (!ok && fail == 0) fail = (int32_t)(k + 1);
    z2_[0] = th[0];
    z2_[1] = th[1];
  }
  // Outputs of a step that are NOT on the recursion's dependent path (offsets, chol of the inverse):
  // they only need the step's own factor S, so they are evaluated one iteration late, in the same
  // basic block as the next step's dependent chain -- the two interleave instead of queueing up
  // behind each other in the in-order pipeline.
  __device__ __forceinline__ void emit(T* const* out, int j) {
    if constexpr (D == 2) {
      emit2(out, j);
      return;
    }
    T off[D], Qc[DD], r2[D];
#pragma unroll
    for (int i = 0; i < D; ++i) off[i] = z[i];
    trsv_lower<T, D>(S, rinv, off);
    trsv_lower_t<T, D>(S, rinv, off);
    st_s<T, D>(out[1] + j * D, off);
    chol_inverse<T, D>(Qc, S, rinv);
    chol_lower<T, D>(Qc, r2);
    zero_upper<T, D>(Qc);
    st_s<T, DD>(out[2] + j * DD, Qc);
  }
  // dependent chain of step k: D_k, its factor and z_k from the factor of step k+1
  __device__ __forceinline__ void advance(const T* const* in, T* const* out, int j, int64_t k) {
    if constexpr (D == 2) {
      advance2(in, out, j, k);
      return;
    }
    T Dk[DD], th[D], S2[DD], rinv2[D];
    ld_s<T, D>(th, in[0] + j * D);
    ld_s<T, DD>(Dk, in[1] + j * DD);
#pragma unroll
    for (int i = 0; i < DD; ++i) Dk[i] = T(-2) * Dk[i];
    if (k + 1 < Tn_) {
      T A[DD], Th[DD];
      ld_s<T, DD>(A, in[2] + j * DD);
#pragma unroll
      for (int i = 0; i < DD; ++i) Th[i] = A[i];
      trsm_left_lower<T, D>(S, rinv, A);
      trsm_left_lower_t<T, D>(S, rinv, A);  // A_k = D_{k+1}^{-1} theta_sub_k
      st_s<T, DD>(out[0] + j * DD, A);
#pragma unroll
      for (int r = 0; r < D; ++r)
#pragma unroll
        for (int q = 0; q <= r; ++q) {
          T v = Dk[r * D + q];
#pragma unroll
          for (int s = 0; s < D; ++s) v = Num<T>::fma(-Th[s * D + r], A[s * D + q], v);
          Dk[r * D + q] = v;
        }
      gemv_t_add<T, D>(th, A, z);
    }
#pragma unroll
    for (int i = 0; i < DD; ++i) S2[i] = Dk[i];
    const bool ok = chol_lower<T, D>(S2, rinv2);
    if (!ok && fail == 0) fail = (int32_t)(k + 1);
    // commit the new state only later: emit() of the previous step still reads the old one
#pragma unroll
    for (int i = 0; i < D; ++i) z2_[i] = th[i];
#pragma unroll
    for (int i = 0; i < DD; ++i) S2_[i] = S2[i];
#pragma unroll
    for (int i = 0; i < D; ++i) r2_[i] = rinv2[i];
  }
  __device__ __forceinline__ void commit() {
#pragma unroll
    for (int i = 0; i < DD; ++i) S[i] = S2_[i];
#pragma unroll
    for (int i = 0; i < D; ++i) {
      rinv[i] = r2_[i];
      z[i] = z2_[i];
    }
  }
  __device__ __forceinline__ void tile(const Params&, const T* const* in, T* const* out, int64_t j0,
                                       int ns) {
    if (n_ - j0 < ns) ns = (int)(n_ - j0);  // ragged last segment
    if (ns <= 0) return;
    advance(in, out, ns - 1, k0_ + j0 + ns - 1);
    commit();
    for (int j = ns - 2; j >= 0; --j) {
      advance(in, out, j, k0_ + j0 + j);  // reads the state of step j+1 ...
      emit(out, j + 1);                   // ... and so does this: independent, interleaved by ptxas
      commit();
    }
    emit(out, 0);  // the tile's output stage is handed over when tile() returns
  }
  __device__ __forceinline__ void finish(const Params& p, int64_t v, bool valid) {
    if (!valid || !p.info) return;
    if (p.P == 1) p.info[v] = fail;
    else if (fail) atomicMax(p.info + v / p.P, fail);
  }
};

// pass 1: element (P, Q, R, p, r) of every segment but the first
template <typename T_, int D>
struct NatSummaryCore : NatGeomBase<T_, D> {
  using T = T_;
  using Params = NatToSsmParams<T>;
  static constexpr int DD = D * D;
  static constexpr int NIN = 3, NOUT = 0;
  static constexpr int eout(int) { return 1; }
  static __device__ __forceinline__ StreamGeom in_geom(const Params& p, int i, int64_t v) {
    StreamGeom g = NatGeomBase<T_, D>::in_geom(p, i, v);
    if (v % p.P == 0) g.end = 0;  // the first segment feeds nobody: nothing to load
    return g;
  }
  static void tm_describe(const Params& p, TmStream* in, TmStream*) {
    NatGeomBase<T_, D>::tm_describe(p, in, nullptr);
  }
  static __device__ __forceinline__ StreamGeom out_geom(const Params&, int, int64_t) {
    return StreamGeom{nullptr, 0, 0};
  }
  T S[DD], rinv[D], Pm[DD], Q[DD], R[DD], pv[D], rv[D];
  int32_t fail;
  int64_t Tn_, k0_, n_;
  bool live_, started_;
  __device__ __forceinline__ void init(const Params& p, int64_t v) {
    Tn_ = p.Tn;
    k0_ = (v % p.P) * p.L;
    n_ = seg_steps(p.Tn, k0_, p.L);
    live_ = (v % p.P) > 0 && n_ > 0;
    started_ = false;
    fail = 0;
  }
  __device__ __forceinline__ void tile(const Params&, const T* const* in, T* const*, int64_t j0, int ns) {
    if (!live_) return;
    if (n_ - j0 < ns) ns = (int)(n_ - j0);
    for (int j = ns - 1; j >= 0; --j) {
      const int64_t k = k0_ + j0 + j;
      T Dk[DD], th[D], Ths[DD];
      ld_s<T, D>(th, in[0] + j * D);
      ld_s<T, DD>(Dk, in[1] + j * DD);
#pragma unroll
      for (int i = 0; i < DD; ++i) Dk[i] = T(-2) * Dk[i];
      const bool coupled = k + 1 < Tn_;
      if (coupled) ld_s<T, DD>(Ths, in[2] + j * DD);
      if (!started_) {
        // far end of the segment: P = Th, Q = Ths^T, R = 0, p = th, r = 0 (the last segment of a
        // chain has no incoming block: only P and p matter)
#pragma unroll
        for (int a = 0; a < D; ++a)
#pragma unroll
          for (int b = 0; b < D; ++b) {
            Q[a * D + b] = coupled ? Ths[b * D + a] : T(0);
            R[a * D + b] = T(0);
          }
#pragma unroll
        for (int i = 0; i < D; ++i) {
          pv[i] = th[i];
          rv[i] = T(0);
        }
        started_ = true;
      } else {
        // S = chol(P) of the previous step: W = P^{-1} Q, A = P^{-1} Ths
        T W[DD], A[DD], u[D];
#pragma unroll
        for (int i = 0; i < DD; ++i) {
          W[i] = Q[i];
          A[i] = Ths[i];
        }
        if constexpr (D == 2) {
          // S holds P^{-1} = adj(P) / det(P) of the previous step (i00, i10, i11)
          const T q0 = W[0], q1 = W[1], q2 = W[2], q3 = W[3];
          W[0] = Num<T>::fma(S[0], q0, S[1] * q2);
          W[1] = Num<T>::fma(S[0], q1, S[1] * q3);
          W[2] = Num<T>::fma(S[1], q0, S[2] * q2);
          W[3] = Num<T>::fma(S[1], q1, S[2] * q3);
          const T a0 = A[0], a1 = A[1], a2 = A[2], a3 = A[3];
          A[0] = Num<T>::fma(S[0], a0, S[1] * a2);
          A[1] = Num<T>::fma(S[0], a1, S[1] * a3);
          A[2] = Num<T>::fma(S[1], a0, S[2] * a2);
          A[3] = Num<T>::fma(S[1], a1, S[2] * a3);
        } else {
          trsm_left_lower<T, D>(S, rinv, W);
          trsm_left_lower_t<T, D>(S, rinv, W);
          trsm_left_lower<T, D>(S, rinv, A);
          trsm_left_lower_t<T, D>(S, rinv, A);
        }
        // r' = r + W^T p,  R' = R - Q^T W,  Q' = Ths^T W,  p' = th + A^T p,  P' = Th - Ths^T A
        gemv_t_add<T, D>(rv, W, pv);
        T Qn[DD];
#pragma unroll
        for (int a = 0; a < D; ++a)
#pragma unroll
          for (int b = 0; b < D; ++b) {
            T vr = R[a * D + b], vq = T(0);
#pragma unroll
            for (int s = 0; s < D; ++s) {
              vr = Num<T>::fma(-Q[s * D + a], W[s * D + b], vr);
              vq = Num<T>::fma(Ths[s * D + a], W[s * D + b], vq);
            }
            R[a * D + b] = vr;
            Qn[a * D + b] = vq;
          }
#pragma unroll
        for (int i = 0; i < DD; ++i) Q[i] = Qn[i];
#pragma unroll
        for (int i = 0; i < D; ++i) u[i] = th[i];
        gemv_t_add<T, D>(u, A, pv);
#pragma unroll
        for (int i = 0; i < D; ++i) pv[i] = u[i];
#pragma unroll
        for (int r = 0; r < D; ++r)
#pragma unroll
          for (int q = 0; q <= r; ++q) {
            T v = Dk[r * D + q];
#pragma unroll
            for (int s = 0; s < D; ++s) v = Num<T>::fma(-Ths[s * D + r], A[s * D + q], v);
            Dk[r * D + q] = v;
          }
      }
#pragma unroll
      for (int i = 0; i < DD; ++i) Pm[i] = Dk[i];
      bool ok;
      if constexpr (D == 2) {
        const T det = Num<T>::fma(-Dk[2], Dk[2], Dk[0] * Dk[3]);
        const T rd = Num<T>::rcp(det);
        S[0] = Dk[3] * rd;
        S[1] = -Dk[2] * rd;
        S[2] = Dk[0] * rd;
        ok = Dk[0] > T(0) && det > T(0);
      } else {
#pragma unroll
        for (int i = 0; i < DD; ++i) S[i] = Dk[i];
        ok = chol_lower<T, D>(S, rinv);
      }
      if (!ok && fail == 0) fail = (int32_t)(k + 1);
    }
  }
  __device__ __forceinline__ void finish(const Params& p, int64_t v, bool valid) {
    if (!valid || !live_) return;
    const int64_t c = v / p.P;
    mirror_lower<T, D>(Pm);
    store_vec<T, DD>(p.out_chol + (c * p.Tn + k0_) * DD, Pm);
    store_vec<T, D>(p.out_off + (c * p.Tn + k0_) * D, pv);
    if (n_ >= 2) {
      store_vec<T, DD>(p.out_chol + (c * p.Tn + k0_ + 1) * DD, R);
      store_vec<T, D>(p.out_off + (c * p.Tn + k0_ + 1) * D, rv);
    }
    if (k0_ < p.Tn - 1) store_vec<T, DD>(p.out_a + (c * (p.Tn - 1) + k0_) * DD, Q);
    if (fail && p.info) atomicMax(p.info + c, fail);
  }
};

// Fold of the elements of a BACKWARD linear-fractional sweep (naturals -> SSM, U D U^T), in sweep
// order it = 0..P-1 <-> segment P-1-it.  The state entering segment s < P-1 is (P, p) of the
// combined element of the segments above it (LftElem).  WARP = false: one thread per chain;
// WARP = true (many segments): one warp per chain, lane l owns sweep positions [l*m, (l+1)*m),
// combines them, the warp scans the composites, every lane walks its positions from its prefix.
// Policy: static void load(Elem&, params, c, k0, n, last)  -- element parked by the summary pass
//         static void seed(params, c, k0, const Elem&)      -- park the state entering the segment
template <typename T, int D, bool VEC, class Policy, class Params, bool WARP>
__device__ __forceinline__ void lft_backward_fold(const Params& p, int32_t* info) {
  using Elem = LftElem<T, D, VEC>;
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t c = WARP ? tid / 32 : tid;
  const int lane = WARP ? (int)(tid & 31) : 0;
  if (c >= p.B) return;
  const int64_t m = WARP ? (p.P + 31) / 32 : p.P;
  const int64_t i0 = lane * m;
  int64_t i1 = i0 + m;
  if (i1 > p.P) i1 = p.P;
  int32_t fail = 0;
  Elem X;
  X.clear();
  auto geom = [&](int64_t it, int64_t& k0, int64_t& n, bool& last) {
    const int64_t seg = p.P - 1 - it;
    k0 = seg * p.L;
    n = seg_steps(p.Tn, k0, p.L);
    last = k0 + n >= p.Tn;
    return seg;
  };
  if (WARP) {
    Elem e, other;
    for (int64_t it = i0; it < i1 && it < p.P - 1; ++it) {  // the first segment feeds nobody
      int64_t k0, n;
      bool last;
      geom(it, k0, n, last);
      Policy::load(e, p, c, k0, n, last);
      if (!X.then(e) && fail == 0) fail = (int32_t)(k0 + n + 1);
    }
#pragma unroll 1
    for (int delta = 1; delta < 32; delta <<= 1) {
      other.shfl_up_from(X, delta);
      if (lane >= delta) {
        if (!other.then(X) && fail == 0) fail = (int32_t)(p.Tn - i0 * p.L);
        X = other;
      }
    }
    other.shfl_up_from(X, 1);
    X = other;
    if (lane == 0) X.clear();
  }
  for (int64_t it = i0; it < i1; ++it) {
    int64_t k0, n;
    bool last;
    const int64_t seg = geom(it, k0, n, last);
    Elem e;
    e.clear();
    if (seg > 0) Policy::load(e, p, c, k0, n, last);  // before its slots receive the seed
    if (it > 0) Policy::seed(p, c, k0, X);
    if (seg == 0) break;
    if (!X.then(e) && fail == 0) fail = (int32_t)(k0 + n + 1);
  }
  if (fail && info) atomicMax(info + c, fail);
}

struct NatFoldPolicy {
  template <typename T, int D>
  static __device__ __forceinline__ void load(LftElem<T, D, true>& e, const NatToSsmParams<T>& p,
                                              int64_t c, int64_t k0, int64_t n, bool last) {
    constexpr int DD = D * D;
    e.clear();
    e.empty = 0;
    load_vec_rw<T, DD>(e.P, p.out_chol + (c * p.Tn + k0) * DD);
    load_vec_rw<T, D>(e.p, p.out_off + (c * p.Tn + k0) * D);
    if (!last) {
      load_vec_rw<T, DD>(e.R, p.out_chol + (c * p.Tn + k0 + 1) * DD);
      load_vec_rw<T, D>(e.r, p.out_off + (c * p.Tn + k0 + 1) * D);
      load_vec_rw<T, DD>(e.Q, p.out_a + (c * (p.Tn - 1) + k0) * DD);
    }
  }
  template <typename T, int D>
  static __device__ __forceinline__ void seed(const NatToSsmParams<T>& p, int64_t c, int64_t k0,
                                              const LftElem<T, D, true>& X) {
    store_vec<T, D * D>(p.out_chol + (c * p.Tn + k0) * D * D, X.P);
    store_vec<T, D>(p.out_off + (c * p.Tn + k0) * D, X.p);
  }
};

template <typename T, int D, bool WARP>
__global__ void __launch_bounds__(128)
nat_seed_kernel(const NatToSsmParams<T> p) {
  lft_backward_fold<T, D, true, NatFoldPolicy, NatToSsmParams<T>, WARP>(p, p.info);
}

// ---------------------------------------------------------------------------------------------
// 4-step tiles (half the ring, two resident CTAs per SM, twice the bulk copies): measured on config 5
// in float64 they pay only for the arithmetic-heaviest sweep, NatToSsmCore 561 -> 516 us; the moment
// sweep and both summary passes lose 15-55 % (TMA issue).  Knob 11 = 1: off.
template <class Core> struct PrefersShortTiles { static constexpr bool value = false; };
template <> struct PrefersShortTiles<NatToSsmCore<double, 2>> { static constexpr bool value = true; };

// Cores that describe their streams for the tensor-map engine (sweep_tm.cuh) try it first; it declines
// (cudaErrorNotSupported) geometries it cannot map and the 1-D engine below takes over.  Records of up to
// 4 elements (D <= 2) only: those are the sweeps bound by bulk-copy issue.  Knob 13 = 1: off;
// knob 14: tile geometry (0 default: K=4 in float64, K=8 in float32; 1: K=8, 2: K=4 with one output stage,
// 3: K=2, 4: K=4).
template <class Core, class = void>
struct HasTm : std::false_type {};
template <class Core>
struct HasTm<Core, std::void_t<decltype(&Core::tm_describe)>> : std::true_type {};

template <class Core, class = void>
struct HasAlias : std::false_type {};
template <class Core>
struct HasAlias<Core, std::void_t<decltype(&Core::tm_alias)>> : std::true_type {};

template <class Core>
struct TmAuto {
  static constexpr int max_e() {
    int m = 0;
    for (int i = 0; i < Core::NIN; ++i) m = Core::ein(i) > m ? Core::ein(i) : m;
    for (int i = 0; i < Core::NOUT; ++i) m = Core::eout(i) > m ? Core::eout(i) : m;
    return m;
  }
  static constexpr bool small = max_e() <= 4;
  template <int K, int NSI, int NSO>
  static constexpr bool fits() { return SweepTmCfg<Core, 64, K, NSI, NSO, 4>::FITS; }
  static constexpr bool ok = small && fits<4, 2, 2>();
  // (D = 3 records, 72 bytes: chains of T - 1 entries are 8 bytes off a 16-byte stride for even T, so the
  //  sub-diagonal streams cannot be mapped; the block-tridiagonal solve / inverse-subset / U D U^T cores, two
  //  or three 32-byte streams at D = 2, were measured on this engine and are 20-30 % slower than on 8-step 1-D
  //  tiles (tools/btd_tm_ab.py) -- they stay on the 1-D engine)
  static cudaError_t launch(const typename Core::Params& prm, cudaStream_t s) {
    if constexpr (ok) {
      const int g = tuning(14);
      // cores that pair their streams run on in-place stages (one ring of 4, the loader two tiles ahead);
      // knob 14 = 8: separate rings as below
      if constexpr (HasAlias<Core>::value) {
        constexpr int KI = sizeof(typename Core::T) == 4 ? 8 : 4;  // the same bytes per stage in either dtype
        // (measured on config 5: three stages and three CTAs per SM with one chain per CTA, 0.64 / 0.55 ms
        //  against 0.60 / 0.51 ms for four stages and two CTAs)
        // (two stages of twice the steps -- 256-byte pieces per row instead of 128: 0.635 / 0.500 ms against
        //  0.642 / 0.513 ms in the same run: within the run-to-run spread, not adopted)
        if constexpr (SweepTmiCfg<Core, 64, KI, 4, 4>::FITS) {
          if (g == 0) {
            const cudaError_t e = launch_chain_sweep_tmi<Core, 64, KI, 4, 4>(prm, s);
            if (e != cudaErrorNotSupported) return e;
          }
        }
      }
      if constexpr (fits<8, 2, 2>()) {
        if (g == 1) return launch_chain_sweep_tm<Core, 64, 8, 2, 2, 4>(prm, s);
      }
      if constexpr (fits<4, 2, 1>()) {
        if (g == 2) return launch_chain_sweep_tm<Core, 64, 4, 2, 1, 4>(prm, s);
      }
      if constexpr (fits<2, 2, 2>()) {
        if (g == 3) return launch_chain_sweep_tm<Core, 64, 2, 2, 2, 4>(prm, s);
      }
      if (g == 4) return launch_chain_sweep_tm<Core, 64, 4, 2, 2, 4>(prm, s);
      // (measured on config 5, float64: 32-row CTAs with 3 input stages 0.81 / 0.56 ms, with 8-step tiles
      //  0.82 / 0.53 ms, with the default stages 0.68 / 0.52 ms against 0.66 / 0.53 ms for this default)
      // float32 rows are half as wide: 8-step tiles keep two CTAs per SM and halve the tile overheads
      // (config 5, naturals -> SSM: 0.51 -> 0.41 ms); float64 loses a resident CTA to them (0.67 -> 0.85 ms)
      if constexpr (sizeof(typename Core::T) == 4 && fits<8, 2, 2>()) {
        if (g == 0) return launch_chain_sweep_tm<Core, 64, 8, 2, 2, 4>(prm, s);
      }
      return launch_chain_sweep_tm<Core, 64, 4, 2, 2, 4>(prm, s);
    } else {
      return cudaErrorNotSupported;
    }
  }
};

// Compile-time choice of a ring geometry that fits (chains per CTA, steps per tile, stages).
template <class Core>
struct SweepAuto {
  template <int C, int K, int NSI, int NSO>
  static constexpr bool fits() { return SweepCfg<Core, C, K, NSI, NSO>::FITS; }
  static constexpr int choice = fits<64, 8, 2, 2>() ? 0 : (fits<32, 8, 2, 2>() ? 1 : (fits<32, 4, 2, 2>() ? 2 : -1));
  static constexpr bool ok = choice >= 0;
  // small records (float32, D <= 2) of a pass WITHOUT outputs: tiles of 16 steps fit beside 64 chains --
  // half the bulk copies per step at the same number of compute warps (the D <= 2 sweeps are bound by
  // TMA issue).  Measured on config 5 in f32: summary passes 190 -> 168 us and 138 -> 116 us;
  // passes with outputs got slower (282 -> 342 us, 265 -> 280 us) and keep K = 8.  Knob 11 = 1: off.
  // float64 summary passes lose a resident CTA per SM to the longer tiles (config 5: 290 -> 426 us).
  static constexpr bool long_tiles = fits<64, 16, 2, 2>() && Core::NOUT == 0 && sizeof(typename Core::T) == 4;
  static cudaError_t launch(const typename Core::Params& prm, int64_t nchains, cudaStream_t s) {
    if constexpr (HasTm<Core>::value) {
      if (tuning(13) != 1) {
        const cudaError_t e = TmAuto<Core>::launch(prm, s);
        if (e != cudaErrorNotSupported) return e;
      }
    }
    if constexpr (long_tiles) {
      if (nchains > (int64_t)148 * 48 && tuning(11) != 1) return launch_chain_sweep<Core, 64, 16, 2, 2>(prm, nchains, s, true);
    }
    if constexpr (PrefersShortTiles<Core>::value && fits<64, 4, 2, 2>()) {
      if (tuning(11) != 1 && nchains > (int64_t)148 * 48) return launch_chain_sweep<Core, 64, 4, 2, 2>(prm, nchains, s, true);
    }
    if constexpr (choice == 0) {
      // few chains: one compute warp per CTA so that more SMs get a CTA
      if (nchains <= (int64_t)148 * 48) return launch_chain_sweep<Core, 32, 8, 2, 2>(prm, nchains, s, true);
      // (measured: <32, 16, 2, 2> -- half the bulk copies per step but one compute warp per CTA --
      // is 0-45 % slower on the config-5 transforms: resident compute warps matter more)
      // (measured: DIRECT stores, <64, 8, 2, 0> -- no output stages / storer threads, twice the CTAs
      // per SM -- are 15-40 % slower than bulk stores from the output ring on the same transforms)
      return launch_chain_sweep<Core, 64, 8, 2, 2>(prm, nchains, s, true);
    } else if constexpr (choice == 1) {
      return launch_chain_sweep<Core, 32, 8, 2, 2>(prm, nchains, s, true);
    } else if constexpr (choice == 2) {
      return launch_chain_sweep<Core, 32, 4, 2, 2>(prm, nchains, s, true);
    } else {
      return cudaErrorInvalidConfiguration;
    }
  }
};

}  // namespace mf
