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
#include "nat_kernels.cuh"
#include "sweep.cuh"

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
template <typename T>
struct SsmMomentsParams {
  const T *mu0, *chol_p0, *a, *b, *chol_q;
  T *o_vec, *o_diag, *o_sub;  // means|eta_lin [B,T,D], covs|eta_diag [B,T,D,D], lag-one [B,T-1,D,D]
  int64_t B, Tn;
};

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
  static __device__ __forceinline__ int64_t num_chains(const Params& p) { return p.B; }
  static __device__ __forceinline__ int64_t max_steps(const Params& p) { return p.Tn; }
  static __device__ __forceinline__ StreamGeom in_geom(const Params& p, int i, int64_t c) {
    return geom_incoming<T>(i == 0 ? p.a : (i == 1 ? p.b : p.chol_q), c, p.Tn, ein(i));
  }
  static __device__ __forceinline__ StreamGeom out_geom(const Params& p, int i, int64_t c) {
    if (i == 2) return geom_incoming<T>(p.o_sub, c, p.Tn, DD);
    return geom_states<T>(i == 0 ? p.o_vec : p.o_diag, c, p.Tn, eout(i));
  }
  T mu[D], P[DD];
  __device__ __forceinline__ void init(const Params& p, int64_t c) {
    T L[DD];
    load_vec<T, D>(mu, p.mu0 + c * D);
    load_vec<T, DD>(L, p.chol_p0 + c * DD);
    llt<T, D>(P, L);
  }
  __device__ __forceinline__ void tile(const Params& p, const T* const* in, T* const* out,
                                       int64_t j0, int ns) {
    for (int j = 0; j < ns; ++j) {
      if (j0 + j > 0) {
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

// ---------------------------------------------------------------------------------------------
template <typename T>
struct SsmAffineParams {
  const T *mu0, *chol_p0, *a, *b, *chol_q, *eps;
  T* out;
  int64_t n, Bm, Tn;
};

template <typename T_, int D, bool NOISE>
struct SsmAffineCore {
  using T = T_;
  using Params = SsmAffineParams<T>;
  static constexpr int DD = D * D;
  static constexpr int NIN = NOISE ? 4 : 2, NOUT = 1;
  static constexpr bool BACKWARD = false;
  static constexpr int ein(int i) { return (i == 0 || i == 2) ? DD : D; }
  static constexpr int eout(int) { return D; }
  static __device__ __forceinline__ int64_t num_chains(const Params& p) { return p.n; }
  static __device__ __forceinline__ int64_t max_steps(const Params& p) { return p.Tn; }
  static __device__ __forceinline__ StreamGeom in_geom(const Params& p, int i, int64_t c) {
    if (i == 3) return geom_states<T>(p.eps, c, p.Tn, D);
    return geom_incoming<T>(i == 0 ? p.a : (i == 1 ? p.b : p.chol_q), c % p.Bm, p.Tn, ein(i));
  }
  static __device__ __forceinline__ StreamGeom out_geom(const Params& p, int, int64_t c) {
    return geom_states<T>(p.out, c, p.Tn, D);
  }
  T x[D];
  int64_t cm;
  __device__ __forceinline__ void init(const Params& p, int64_t c) {
    cm = c % p.Bm;
    load_vec<T, D>(x, p.mu0 + cm * D);
  }
  __device__ __forceinline__ void tile(const Params& p, const T* const* in, T* const* out,
                                       int64_t j0, int ns) {
    for (int j = 0; j < ns; ++j) {
      if (j0 + j > 0) {
        T A[DD], off[D];
        ld_s<T, DD>(A, in[0] + j * DD);
        ld_s<T, D>(off, in[1] + j * D);
        if (NOISE) {
          T L[DD], e[D];
          ld_s<T, DD>(L, in[2] + j * DD);
          zero_upper<T, D>(L);
          ld_s<T, D>(e, in[3] + j * D);
          gemv_add<T, D>(off, L, e);
        }
        gemv_add<T, D>(off, A, x);
#pragma unroll
        for (int r = 0; r < D; ++r) x[r] = off[r];
      } else if (NOISE) {
        T L[DD], e[D];
        load_vec<T, DD>(L, p.chol_p0 + cm * DD);
        zero_upper<T, D>(L);
        ld_s<T, D>(e, in[3] + j * D);
        gemv_add<T, D>(x, L, e);
      }
      st_s<T, D>(out[0] + j * D, x);
    }
  }
  __device__ __forceinline__ void finish(const Params&, int64_t, bool) {}
};

// ---------------------------------------------------------------------------------------------
template <typename T>
struct SsmKlParams {
  const T *q_mu0, *q_chol_p0, *q_a, *q_b, *q_chol_q;
  const T *p_mu0, *p_chol_p0, *p_a, *p_b, *p_chol_q;
  T* out;
  int64_t B, Tn;
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
  static __device__ __forceinline__ int64_t num_chains(const Params& p) { return p.B; }
  static __device__ __forceinline__ int64_t max_steps(const Params& p) { return p.Tn; }
  static __device__ __forceinline__ StreamGeom in_geom(const Params& p, int i, int64_t c) {
    const T* base = i == 0 ? p.q_a : i == 1 ? p.q_b : i == 2 ? p.q_chol_q
                  : i == 3 ? p.p_a : i == 4 ? p.p_b : p.p_chol_q;
    return geom_incoming<T>(base, c, p.Tn, ein(i));
  }
  static __device__ __forceinline__ StreamGeom out_geom(const Params&, int, int64_t) {
    return StreamGeom{nullptr, 0, 0};
  }
  T mu[D], P[DD], kl;
  LogProd<T> ratio;
  __device__ __forceinline__ void init(const Params& p, int64_t c) {
    T Lq[DD], Lp[DD], dm[D];
    ratio.init();
    load_vec<T, D>(mu, p.q_mu0 + c * D);
    load_vec<T, DD>(Lq, p.q_chol_p0 + c * DD);
    load_vec<T, DD>(Lp, p.p_chol_p0 + c * DD);
    load_vec<T, D>(dm, p.p_mu0 + c * D);
#pragma unroll
    for (int i = 0; i < D; ++i) dm[i] = mu[i] - dm[i];
    kl = kl_gauss_term<T, D>(Lp, Lq, nullptr, dm, nullptr, ratio);
    llt<T, D>(P, Lq);
  }
  __device__ __forceinline__ void tile(const Params&, const T* const* in, T* const*, int64_t j0,
                                       int ns) {
    for (int j = 0; j < ns; ++j) {
      if (j0 + j == 0) continue;
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
  __device__ __forceinline__ void finish(const Params& p, int64_t c, bool valid) {
    if (valid) p.out[c] = kl + ratio.log_abs();
  }
};

// ---------------------------------------------------------------------------------------------
template <typename T>
struct NatToSsmParams {
  const T *th_lin, *th_diag, *th_sub;
  T *out_a, *out_off, *out_chol;
  int32_t* info;
  int64_t B, Tn;
};

// Backward U D U^T sweep (see nat_to_ssm_kernel in nat_kernels.cuh for the algebra).
template <typename T_, int D>
struct NatToSsmCore {
  using T = T_;
  using Params = NatToSsmParams<T>;
  static constexpr int DD = D * D;
  static constexpr int NIN = 3, NOUT = 3;
  static constexpr bool BACKWARD = true;
  static constexpr int ein(int i) { return i == 0 ? D : DD; }
  static constexpr int eout(int i) { return i == 1 ? D : DD; }
  static __device__ __forceinline__ int64_t num_chains(const Params& p) { return p.B; }
  static __device__ __forceinline__ int64_t max_steps(const Params& p) { return p.Tn; }
  static __device__ __forceinline__ StreamGeom in_geom(const Params& p, int i, int64_t c) {
    if (i == 2) return geom_outgoing<T>(p.th_sub, c, p.Tn, DD);
    return geom_states<T>(i == 0 ? p.th_lin : p.th_diag, c, p.Tn, ein(i));
  }
  static __device__ __forceinline__ StreamGeom out_geom(const Params& p, int i, int64_t c) {
    if (i == 0) return geom_outgoing<T>(p.out_a, c, p.Tn, DD);
    return geom_states<T>(i == 1 ? p.out_off : p.out_chol, c, p.Tn, eout(i));
  }
  T S[DD], rinv[D], z[D];
  int32_t fail;
  int64_t Tn_;
  __device__ __forceinline__ void init(const Params& p, int64_t) {
    fail = 0;
    Tn_ = p.Tn;
#pragma unroll
    for (int i = 0; i < D; ++i) z[i] = T(0);
  }
  __device__ __forceinline__ void tile(const Params&, const T* const* in, T* const* out, int64_t j0,
                                       int ns) {
    for (int j = ns - 1; j >= 0; --j) {
      const int64_t k = j0 + j;
      T Dk[DD], th[D], off[D], Qc[DD], r2[D];
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
      for (int i = 0; i < D; ++i) z[i] = th[i];
#pragma unroll
      for (int i = 0; i < DD; ++i) S[i] = Dk[i];
      const bool ok = chol_lower<T, D>(S, rinv);
      if (!ok && fail == 0) fail = (int32_t)(k + 1);
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
  }
  __device__ __forceinline__ void finish(const Params& p, int64_t c, bool valid) {
    if (valid && p.info) p.info[c] = fail;
  }
};

// ---------------------------------------------------------------------------------------------
// Compile-time choice of a ring geometry that fits (chains per CTA, steps per tile, stages).
template <class Core>
struct SweepAuto {
  template <int C, int K, int NSI, int NSO>
  static constexpr bool fits() { return SweepCfg<Core, C, K, NSI, NSO>::FITS; }
  static constexpr int choice = fits<64, 8, 2, 2>() ? 0 : (fits<32, 8, 2, 2>() ? 1 : (fits<32, 4, 2, 2>() ? 2 : -1));
  static constexpr bool ok = choice >= 0;
  static cudaError_t launch(const typename Core::Params& prm, int64_t nchains, cudaStream_t s) {
    if constexpr (choice == 0) {
      // few chains: one compute warp per CTA so that more SMs get a CTA
      if (nchains <= (int64_t)148 * 48) return launch_chain_sweep<Core, 32, 8, 2, 2>(prm, nchains, s);
      return launch_chain_sweep<Core, 64, 8, 2, 2>(prm, nchains, s);
    } else if constexpr (choice == 1) {
      return launch_chain_sweep<Core, 32, 8, 2, 2>(prm, nchains, s);
    } else if constexpr (choice == 2) {
      return launch_chain_sweep<Core, 32, 4, 2, 2>(prm, nchains, s);
    } else {
      return cudaErrorInvalidConfiguration;
    }
  }
};

}  // namespace mf
