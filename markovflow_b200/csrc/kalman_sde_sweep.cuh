// Kalman log-likelihood of a stationary Matern SDE prior with the state-space model built INSIDE
// the sweep from the time deltas.
//
// Reference path replaced: SDEKernel.state_space_model (kernels/sde_kernel.py:153-171) ->
// StationaryKernel.transition_statistics (:421-446, Q_k = Pinf - A_k Pinf A_k^T + jitter I) with the
// closed-form Matern state transitions (kernels/matern.py:80-86 Matern12, :299-324 Matern32,
// :434-460 Matern52), generate_emission_model (sde_kernel.py:173-211, H = [1,0,..]) and
// KalmanFilter.log_likelihood (kalman_filter.py:184-255).  The reference materialises A [T-1,D,D],
// chol Q [T-1,D,D], b, H for the filter to read back ((2D^2+2D+1) values per step); here a step
// reads TWO values (dt_k, y_k) and A_k, Q_k live in registers only.
//
// Streams per step: dt (1), y (1).  A "chain" of the sweep is a virtual chain (segment p of series
// c) exactly as in kalman_sweep.cuh; the per-warp join and the ordered reduction are shared with it.
#pragma once
#include "kalman_sweep.cuh"

#include <type_traits>

namespace mf {

template <typename T>
struct KalmanSdeParams {
  const T* ls;      // [B] lengthscales
  const T* var;     // [B] variances
  const T* dt;      // [B,NT]  NT = Tn - first_is_initial; dt[k - fi] leads INTO local step k
  const T* obs;     // [B,Tn]
  const T* chol_r;  // [1]
  T jitter;
  int64_t B, Tn, P, L;
  int first_is_initial;
  int reduce_warp;  // summaries: one element per 32 consecutive segments (out [B,P/32,N])
  T* out;           // summaries [B,P,N] (or [B,P/32,N]) | log-likelihoods [B]  (P == 1)
};

template <typename T> struct SdeNum;
template <> struct SdeNum<double> {
  static __device__ __forceinline__ double exp(double x) { return ::exp(x); }
  static __device__ __forceinline__ double sqrt(double x) { return ::sqrt(x); }
};
template <> struct SdeNum<float> {
  static __device__ __forceinline__ float exp(float x) { return ::expf(x); }
  static __device__ __forceinline__ float sqrt(float x) { return ::sqrtf(x); }
};

// Closed-form Matern-(D - 1/2) transition statistics of one chain.
//   A(dt) = exp(-lam dt) (I + N dt + N^2 dt^2 / 2),  N = F + lam I (nilpotent of order D)
//   Q(dt) = Pinf - A Pinf A^T + jitter I
template <typename T, int D>
struct MaternStats {
  static constexpr int DD = D * D;
  T lam, jit;
  T n1[DD];   // N
  T n2[DD];   // N^2 / 2  (D == 3 only)
  T pinf[DD];
  __device__ __forceinline__ void init(T lengthscale, T variance, T jitter) {
    jit = jitter;
#pragma unroll
    for (int i = 0; i < DD; ++i) { n1[i] = T(0); n2[i] = T(0); pinf[i] = T(0); }
    if (D == 1) {
      lam = T(1) / lengthscale;
      pinf[0] = variance;
    } else if (D == 2) {
      lam = SdeNum<T>::sqrt(T(3)) / lengthscale;
      n1[0] = lam; n1[1] = T(1); n1[D] = -lam * lam; n1[D + 1] = -lam;
      pinf[0] = variance; pinf[D + 1] = variance * lam * lam;
    } else {
      lam = SdeNum<T>::sqrt(T(5)) / lengthscale;
      const T l2 = lam * lam, l3 = l2 * lam;
      // N = [[lam,1,0],[0,lam,1],[-lam^3,-3lam^2,-2lam]]
      n1[0] = lam; n1[1] = T(1);
      n1[D + 1] = lam; n1[D + 2] = T(1);
      n1[2 * D] = -l3; n1[2 * D + 1] = T(-3) * l2; n1[2 * D + 2] = T(-2) * lam;
      // N^2/2 = [[lam^2,2lam,1],[-lam^3,-2lam^2,-lam],[lam^4,2lam^3,lam^2]] / 2
      n2[0] = T(0.5) * l2; n2[1] = lam; n2[2] = T(0.5);
      n2[D] = T(-0.5) * l3; n2[D + 1] = -l2; n2[D + 2] = T(-0.5) * lam;
      n2[2 * D] = T(0.5) * l2 * l2; n2[2 * D + 1] = l3; n2[2 * D + 2] = T(0.5) * l2;
      const T c = variance * l2 / T(3);
      pinf[0] = variance; pinf[2] = -c; pinf[D + 1] = c; pinf[2 * D] = -c;
      pinf[2 * D + 2] = variance * l2 * l2;
    }
  }
  // F = A(dt) (full).  Q is never formed: with Ct = C - Pinf the prediction C' = F C F^T + Q is
  // Ct' = F Ct F^T + jitter I (Q = Pinf - F Pinf F^T + jitter I), see cov_predict_centred.
  __device__ __forceinline__ void make(T dt, T* __restrict__ F) const {
    const T e = SdeNum<T>::exp(-lam * dt);
    if (D == 1) {
      F[0] = e;
      return;
    }
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int j = 0; j < D; ++j) {
        T v = n1[i * D + j];
        if (D >= 3) v = Num<T>::fma(dt, n2[i * D + j], v);
        v = Num<T>::fma(dt, v, (i == j) ? T(1) : T(0));
        F[i * D + j] = e * v;
      }
  }
};

// Ct <- F Ct F^T + jit I  (Ct = covariance minus the stationary covariance, full symmetric)
template <typename T, int D>
__device__ __forceinline__ void cov_predict_centred(T* __restrict__ P, const T* __restrict__ F, T jit) {
  T FP[D * D];
  gemm<T, D>(FP, F, P);
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      T v = (i == j) ? jit : T(0);
#pragma unroll
      for (int q = 0; q < D; ++q) v = Num<T>::fma(FP[i * D + q], F[j * D + q], v);
      P[i * D + j] = v;
      P[j * D + i] = v;
    }
}

// v <- F v
template <typename T, int D>
__device__ __forceinline__ void mean_predict0(T* __restrict__ v, const T* __restrict__ F) {
  T t[D];
#pragma unroll
  for (int i = 0; i < D; ++i) {
    T a = F[i * D] * v[0];
#pragma unroll
    for (int q = 1; q < D; ++q) a = Num<T>::fma(F[i * D + q], v[q], a);
    t[i] = a;
  }
#pragma unroll
  for (int i = 0; i < D; ++i) v[i] = t[i];
}

// Sinks for a zero-offset stationary transition F and an observation of the FIRST state component
// with whitener w (h' = w e_0, y' = w y): the arithmetic of filter_absorb / elem_absorb
// (kalman_core.cuh) with the structural zeros of h removed.  While a sink walks, its covariance
// (st.P / e.C) holds the CENTRED value C - Pinf; p0 = first column of Pinf.
template <typename T, int D>
struct SdeFilterSink {
  FilterState<T, D> st;
  T quad;
  LogProd<T> det;
  __device__ __forceinline__ void init(const T*) { quad = T(0); det.init(); }
  __device__ __forceinline__ void start_prior(T jit) {  // C = Pinf + jit I
#pragma unroll
    for (int i = 0; i < D; ++i) {
      st.m[i] = T(0);
#pragma unroll
      for (int j = 0; j < D; ++j) st.P[i * D + j] = (i == j) ? jit : T(0);
    }
  }
  __device__ __forceinline__ void transition(const T* F, T jit) {
    mean_predict0<T, D>(st.m, F);
    cov_predict_centred<T, D>(st.P, F, jit);
  }
  __device__ __forceinline__ void absorb(T w, T yw, const T* p0) {
    T g[D];
#pragma unroll
    for (int i = 0; i < D; ++i) g[i] = (st.P[i * D] + p0[i]) * w;
    const T s = Num<T>::fma(w, g[0], T(1));
    const T v = Num<T>::fma(-w, st.m[0], yw);
    const T rs = Num<T>::rcp(s);
    const T vs = v * rs;
    quad = Num<T>::fma(v, vs, quad);
    det.mul_lazy(s);
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
  __device__ __forceinline__ T loglik(T logw, int64_t nobs) const {
    return T(-0.5) * (quad + det.log_abs()) + logw - T(0.5 * 1.8378770664093454836) * T(nobs);
  }
};

template <typename T, int D>
struct SdeElemSink {
  ScanElem<T, D> e;
  T quad;
  LogProd<T> det;
  // identity element (C = 0), centred
  __device__ __forceinline__ void init(const T* pinf) {
    elem_identity<T, D>(e);
#pragma unroll
    for (int i = 0; i < D * D; ++i) e.C[i] = -pinf[i];
    quad = T(0);
    det.init();
  }
  __device__ __forceinline__ void start_prior(T jit) {  // C = Pinf + jit I, no dependence on the past
    e.ell = T(0);
#pragma unroll
    for (int i = 0; i < D; ++i) {
      e.b[i] = T(0);
      e.eta[i] = T(0);
#pragma unroll
      for (int j = 0; j < D; ++j) {
        e.A[i * D + j] = T(0);
        e.J[i * D + j] = T(0);
        e.C[i * D + j] = (i == j) ? jit : T(0);
      }
    }
  }
  __device__ __forceinline__ void transition(const T* F, T jit) {
    T FA[D * D];
    gemm<T, D>(FA, F, e.A);
#pragma unroll
    for (int i = 0; i < D * D; ++i) e.A[i] = FA[i];
    mean_predict0<T, D>(e.b, F);
    cov_predict_centred<T, D>(e.C, F, jit);
  }
  __device__ __forceinline__ void absorb(T w, T yw, const T* p0) {
    T g[D], a[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
      g[i] = (e.C[i * D] + p0[i]) * w;  // C h
      a[i] = e.A[i] * w;      // A^T h  (row 0 of A)
    }
    const T s = Num<T>::fma(w, g[0], T(1));
    const T v = Num<T>::fma(-w, e.b[0], yw);
    const T rs = Num<T>::rcp(s);
    const T vs = v * rs;
    quad = Num<T>::fma(v, vs, quad);
    det.mul_lazy(s);
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const T ki = g[i] * rs;
      const T wi = a[i] * rs;
      e.b[i] = Num<T>::fma(g[i], vs, e.b[i]);
      e.eta[i] = Num<T>::fma(a[i], vs, e.eta[i]);
#pragma unroll
      for (int j = 0; j < D; ++j) e.A[i * D + j] = Num<T>::fma(-ki, a[j], e.A[i * D + j]);
#pragma unroll
      for (int j = 0; j <= i; ++j) {
        const T c = Num<T>::fma(-ki, g[j], e.C[i * D + j]);
        e.C[i * D + j] = c;
        e.C[j * D + i] = c;
        const T jj = Num<T>::fma(wi, a[j], e.J[i * D + j]);
        e.J[i * D + j] = jj;
        e.J[j * D + i] = jj;
      }
    }
  }
  // back to the uncentred covariance; fold the scalar terms of nobs observations into ell
  __device__ __forceinline__ void finalize(const T* pinf, T logw, int64_t nobs) {
#pragma unroll
    for (int i = 0; i < D * D; ++i) e.C[i] += pinf[i];
    e.ell += T(-0.5) * (quad + det.log_abs()) + logw - T(0.5 * 1.8378770664093454836) * T(nobs);
  }
};

template <typename T_, int D>
struct KalmanSdeCoreBase {
  using T = T_;
  using Params = KalmanSdeParams<T>;
  static constexpr int DD = D * D;
  static constexpr int NIN = 2, NOUT = 0;
  static constexpr bool BACKWARD = false;
  static constexpr int ein(int) { return 1; }
  static constexpr int eout(int) { return 1; }
  static __device__ __forceinline__ int64_t num_chains(const Params& p) { return p.B * p.P; }
  static __device__ __forceinline__ int64_t max_steps(const Params& p) { return p.L; }
  static __device__ __forceinline__ StreamGeom out_geom(const Params&, int, int64_t) {
    return StreamGeom{nullptr, 0, 0};
  }
  static __device__ __forceinline__ StreamGeom in_geom(const Params& p, int i, int64_t chain) {
    const int64_t c = chain / p.P, seg = chain % p.P;
    const int64_t k0 = seg * p.L;
    int64_t steps = p.Tn - k0;
    if (steps > p.L) steps = p.L;
    if (steps < 0) steps = 0;
    const int fi = p.first_is_initial;
    constexpr int ES = (int)sizeof(T);
    StreamGeom sg;
    sg.end = steps;
    sg.first = 0;
    if (i == 0) {  // the delta leading INTO step k is entry k - fi
      sg.step0 = (char*)const_cast<T*>(p.dt) + (c * (p.Tn - fi) + k0 - fi) * (int64_t)ES;
      sg.first = (fi && k0 == 0) ? 1 : 0;
    } else {
      sg.step0 = (char*)const_cast<T*>(p.obs) + (c * p.Tn + k0) * (int64_t)ES;
    }
    return sg;
  }

  MaternStats<T, D> ms_;
  T p0_[D];  // first column of Pinf
  int64_t steps_;
  T w1_;  // 1 / chol_r
  int nobs_;
  bool prior_start_;

  __device__ __forceinline__ T log_whiteners() const {
    return T(nobs_) * Num<T>::log(Num<T>::abs(w1_));
  }

  __device__ __forceinline__ void init_base(const Params& p, int64_t chain) {
    const int64_t c = chain / p.P, seg = chain % p.P;
    const int64_t k0 = seg * p.L;
    steps_ = p.Tn - k0;
    if (steps_ > p.L) steps_ = p.L;
    if (steps_ < 0) steps_ = 0;  // padding slot past the end of the series: identity element
    w1_ = Num<T>::rcp(p.chol_r[0]);
    nobs_ = 0;
    prior_start_ = p.first_is_initial && k0 == 0;
    ms_.init(p.ls[c], p.var[c], p.jitter);
#pragma unroll
    for (int i = 0; i < D; ++i) p0_[i] = ms_.pinf[i * D];
  }

  template <class Sink>
  __device__ __forceinline__ void walk_tile(const T* const* in, int64_t j0, int ns, Sink& sink) {
    int n = ns;
    if (j0 + n > steps_) n = (int)(steps_ - j0);
    if (n <= 0) return;
    // F of step j+1 is built while step j runs down its dependent chain: two ping-pong sets
    struct Rec {
      T F[DD], yw;
    };
    auto fetch = [&](Rec& rec, int j) {
      ms_.make(in[0][j], rec.F);
      rec.yw = in[1][j] * w1_;
    };
    auto step = [&](Rec& rec) {
      sink.transition(rec.F, ms_.jit);
      sink.absorb(w1_, rec.yw, p0_);
      ++nobs_;
    };
    Rec ra, rb;
    int j = 0;
    if (j0 == 0 && prior_start_) {  // the series' very first step starts from the stationary prior
      sink.start_prior(ms_.jit);
      sink.absorb(w1_, in[1][0] * w1_, p0_);
      ++nobs_;
      sink.det.peel();
      j = 1;
      if (n <= 1) return;
    }
    fetch(ra, j);
    for (; j + 1 < n; j += 2) {
      fetch(rb, j + 1);
      step(ra);
      if (j + 2 < n) fetch(ra, j + 2);
      step(rb);
      sink.det.peel();  // two pivots multiplied per renormalisation
    }
    if (j < n) {
      step(ra);
      sink.det.peel();
    }
  }

  // The same walk straight from global memory (no staging): dtp[j] is the delta leading into local
  // step j, yp[j] its observation.  Raw values travel R steps ahead of their use in registers; the
  // F of the R steps of a group do not depend on the filter state, so the compiler overlaps
  // them with the dependent chain of the previous steps.
  template <class Sink>
  __device__ __forceinline__ void walk_direct(const T* __restrict__ dtp, const T* __restrict__ yp,
                                              Sink& sink) {
    const int64_t n = steps_;
    if (n <= 0) return;
    int64_t j = 0;
    if (prior_start_) {
      sink.start_prior(ms_.jit);
      sink.absorb(w1_, __ldg(yp) * w1_, p0_);
      ++nobs_;
      sink.det.peel();
      j = 1;
    }
    constexpr int R = 4;
    T dtr[R], yr[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const bool ok = j + r < n;
      dtr[r] = ok ? __ldg(dtp + j + r) : T(0);
      yr[r] = ok ? __ldg(yp + j + r) : T(0);
    }
    for (; j + R <= n; j += R) {
      T dtn[R], yn[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const bool ok = j + R + r < n;
        dtn[r] = ok ? __ldg(dtp + j + R + r) : T(0);
        yn[r] = ok ? __ldg(yp + j + R + r) : T(0);
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        T F[DD];
        ms_.make(dtr[r], F);
        sink.transition(F, ms_.jit);
        sink.absorb(w1_, yr[r] * w1_, p0_);
        if (r & 1) sink.det.peel();
      }
      nobs_ += R;
#pragma unroll
      for (int r = 0; r < R; ++r) { dtr[r] = dtn[r]; yr[r] = yn[r]; }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (j + r < n) {
        T F[DD];
        ms_.make(dtr[r], F);
        sink.transition(F, ms_.jit);
        sink.absorb(w1_, yr[r] * w1_, p0_);
        ++nobs_;
        sink.det.peel();
      }
    }
  }
};

// Range element of every virtual chain (then a per-warp join, as KalmanSummaryCore).
template <typename T_, int D>
struct KalmanSdeSummaryCore : KalmanSdeCoreBase<T_, D> {
  using Base = KalmanSdeCoreBase<T_, D>;
  using T = T_;
  using Params = typename Base::Params;
  SdeElemSink<T, D> sink;
  __device__ __forceinline__ void init(const Params& p, int64_t chain) {
    this->init_base(p, chain);
    sink.init(this->ms_.pinf);
  }
  __device__ __forceinline__ void tile(const Params&, const T* const* in, T* const*, int64_t j0, int ns) {
    this->walk_tile(in, j0, ns, sink);
  }
  __device__ __forceinline__ void finish(const Params& p, int64_t chain, bool valid) {
    if (valid) sink.finalize(this->ms_.pinf, this->log_whiteners(), this->nobs_); else elem_identity<T, D>(sink.e);
    if (!p.reduce_warp) {
      if (valid) elem_store<T, D>(p.out + chain * ScanElem<T, D>::N, sink.e);
      return;
    }
    const int lane = threadIdx.x & 31;
    ScanElem<T, D> other, tmp;
#pragma unroll 1
    for (int delta = 1; delta < 32; delta <<= 1) {
      elem_shfl_up<T, D>(other, sink.e, delta);
      if (lane >= delta) {
        if (D <= 2) elem_combine_inl<T, D>(tmp, other, sink.e); else elem_combine<T, D>(tmp, other, sink.e);
        sink.e = tmp;
      }
    }
    if (lane == 31 && valid) elem_store<T, D>(p.out + (chain / 32) * ScanElem<T, D>::N, sink.e);
  }
};

// One virtual chain per series (many series): the plain sequential filter.
template <typename T_, int D>
struct KalmanSdeFilterCore : KalmanSdeCoreBase<T_, D> {
  using Base = KalmanSdeCoreBase<T_, D>;
  using T = T_;
  using Params = typename Base::Params;
  SdeFilterSink<T, D> sink;
  __device__ __forceinline__ void init(const Params& p, int64_t chain) {
    this->init_base(p, chain);
    sink.init(this->ms_.pinf);
  }
  __device__ __forceinline__ void tile(const Params&, const T* const* in, T* const*, int64_t j0, int ns) {
    this->walk_tile(in, j0, ns, sink);
  }
  __device__ __forceinline__ void finish(const Params& p, int64_t chain, bool valid) {
    if (valid) p.out[chain] = sink.loglik(this->log_whiteners(), this->nobs_);
  }
};

// ---- direct (unstaged) kernels: every warp computes ----------------------------------------------
// The sweep above spends two thirds of a CTA's registers on producer threads; with 16 bytes per step
// there is nothing to stage, so thread = virtual chain, all warps compute, loads run R steps ahead.
// SUMMARY: range element per virtual chain (+ per-warp join when reduce_warp); otherwise the plain
// filter with out[chain] = log-likelihood (P == 1).
template <typename T, int D, bool SUMMARY, int NT, int MINB = 1>
__global__ void __launch_bounds__(NT, MINB)
kalman_sde_direct_kernel(const KalmanSdeParams<T> p) {
  using Core = typename std::conditional<SUMMARY, KalmanSdeSummaryCore<T, D>, KalmanSdeFilterCore<T, D>>::type;
  const int64_t chain = (int64_t)blockIdx.x * NT + threadIdx.x;
  const int64_t nchains = p.B * p.P;
  const bool valid = chain < nchains;
  Core core;
  if (valid) {
    core.init(p, chain);
    const StreamGeom gd = Core::in_geom(p, 0, chain), gy = Core::in_geom(p, 1, chain);
    core.walk_direct(reinterpret_cast<const T*>(gd.step0), reinterpret_cast<const T*>(gy.step0), core.sink);
  }
  core.finish(p, chain, valid);
}

}  // namespace mf
