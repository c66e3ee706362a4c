// Kalman log-likelihood kernels (reference markovflow/kalman_filter.py:184-255).
//
//  * kalman_loglik_chain_kernel     one thread per chain, sequential filter (many chains)
//  * parallel-in-time path for few, long chains (BASELINE config 3):
//      kalman_segment_summary_kernel  thread (segment p, chain): range element of its L steps
//      kalman_summary_scan_kernel     one block per chain: exclusive prefix over the P summaries
//                                     -> filtered (mean, cov) just before every segment (+ total)
//      kalman_seeded_filter_kernel    thread (p, chain): sequential filter over its L steps from
//                                     the seed, partial log-likelihood
//      kalman_partial_sum_kernel      deterministic sum of the P partials per chain
//
// Array convention ("incoming transitions"): with first_is_initial = 1 the chain starts at the
// prior (mu0, chol_p0) and a/b/chol_q hold T-1 transitions (a[k-1] leads INTO step k).  With
// first_is_initial = 0 (a time segment of a longer series held by another GPU rank) a/b/chol_q hold
// T transitions, a[k] leading into local step k, and the state before step 0 comes from a prefix
// element.
#pragma once
#include "kalman_core.cuh"

namespace mf {

template <typename T>
struct KalmanArgs {
  const T* mu0;      // [B,D]           (first_is_initial only)
  const T* chol_p0;  // [B,D,D]
  const T* a;        // [B,NT,D,D]      NT = T-1 or T
  const T* b;        // [B,NT,D]
  const T* chol_q;   // [B,NT,D,D]
  const T* h;        // [Bh,T,m,D]      Bh = 1 or B
  const T* obs;      // [B,T,m]
  const T* chol_r;   // [Tr,m,m]        Tr = 1 or T
  int64_t B, Tn, Bh, Tr;
  int m;
  int first_is_initial;
};

// Walks the steps [k0,k1) of chain c and feeds them to a sink with
//   sink.start_prior(mu0, L0) / sink.transition(F,u,Lq) / sink.absorb(h', y') ; returns sum log W_ii.
template <typename T, int D, bool M1, class Sink>
__device__ __forceinline__ T kalman_walk(const KalmanArgs<T>& g, int64_t c, int64_t k0, int64_t k1,
                                         Sink& sink, int64_t& nobs) {
  constexpr int DD = D * D;
  const int m = M1 ? 1 : g.m;
  const int64_t nt = g.Tn - g.first_is_initial;
  const T* ap = g.a + c * nt * DD;
  const T* bp = g.b + c * nt * D;
  const T* qp = g.chol_q + c * nt * DD;
  const T* hp = g.h + (g.Bh == 1 ? 0 : c) * g.Tn * (int64_t)m * D;
  const T* yp = g.obs + c * g.Tn * (int64_t)m;
  Whitener<T> wh;
  T w1 = T(1);
  LogProd<T> wdet;
  wdet.init();
  T logw = T(0);
  if (g.Tr == 1) {
    if (M1) {
      w1 = Num<T>::rcp(g.chol_r[0]);
    } else {
      wh.set(g.chol_r, m);
    }
  }
  for (int64_t k = k0; k < k1; ++k) {
    if (k == 0 && g.first_is_initial) {
      T mu[D], L0[DD];
      load_vec<T, D>(mu, g.mu0 + c * D);
      load_vec<T, DD>(L0, g.chol_p0 + c * DD);
      sink.start_prior(mu, L0);
    } else {
      const int64_t ti = k - g.first_is_initial;
      T F[DD], u[D], Lq[DD];
      load_vec<T, DD>(F, ap + ti * DD);
      load_vec<T, D>(u, bp + ti * D);
      load_vec<T, DD>(Lq, qp + ti * DD);
      sink.transition(F, u, Lq);
    }
    if (M1) {
      if (g.Tr != 1) w1 = Num<T>::rcp(g.chol_r[k]);
      if (w1 != T(0)) {  // an infinite noise scale marks a step without observation
        wdet.mul(w1);
        T hv[D];
        load_vec<T, D>(hv, hp + k * D);
#pragma unroll
        for (int i = 0; i < D; ++i) hv[i] *= w1;
        sink.absorb(hv, __ldg(yp + k) * w1);
        ++nobs;
      }
    } else {
      if (g.Tr != 1) wh.set(g.chol_r + k * (int64_t)m * m, m);
      const T* hk = hp + k * (int64_t)m * D;
      const T* yk = yp + k * (int64_t)m;
      for (int i = 0; i < m; ++i) {
        T hv[D];
#pragma unroll
        for (int p = 0; p < D; ++p) hv[p] = T(0);
        T y = T(0);
        for (int j = 0; j <= i; ++j) {
          const T wij = wh.W[i * kMaxObsDim + j];
          y = Num<T>::fma(wij, yk[j], y);
#pragma unroll
          for (int p = 0; p < D; ++p) hv[p] = Num<T>::fma(wij, hk[j * D + p], hv[p]);
        }
        wdet.mul(wh.W[i * kMaxObsDim + i]);
        sink.absorb(hv, y);
        ++nobs;
      }
    }
    sink.det.peel();
  }
  logw = wdet.log_abs();
  return logw;
}

template <typename T, int D>
struct FilterSink {
  FilterState<T, D> st;
  T quad;
  LogProd<T> det;
  __device__ __forceinline__ void init() { quad = T(0); det.init(); }
  __device__ __forceinline__ void start_prior(const T* mu, const T* L0) { filter_init<T, D>(st, mu, L0); }
  __device__ __forceinline__ void transition(const T* F, const T* u, const T* Lq) {
    filter_predict<T, D>(st, F, u, Lq);
  }
  __device__ __forceinline__ void absorb(const T* h, T y) { filter_absorb<T, D>(st, h, y, quad, det); }
  // log-likelihood of nobs scalar observations given log|W| of their whiteners
  __device__ __forceinline__ T loglik(T logw, int64_t nobs) const {
    return T(-0.5) * (quad + det.log_abs()) + logw - T(0.5 * 1.8378770664093454836) * T(nobs);
  }
};

template <typename T, int D>
struct ElemSink {
  ScanElem<T, D> e;
  T quad;
  LogProd<T> det;
  __device__ __forceinline__ void init() { elem_identity<T, D>(e); quad = T(0); det.init(); }
  __device__ __forceinline__ void start_prior(const T* mu, const T* L0) { elem_prior<T, D>(e, mu, L0); }
  __device__ __forceinline__ void transition(const T* F, const T* u, const T* Lq) {
    elem_transition<T, D>(e, F, u, Lq);
  }
  __device__ __forceinline__ void absorb(const T* h, T y) { elem_absorb<T, D>(e, h, y, quad, det); }
  // fold the accumulated scalar terms of nobs absorbed observations into the element's ell
  __device__ __forceinline__ void finalize(T logw, int64_t nobs) {
    e.ell += T(-0.5) * (quad + det.log_abs()) + logw - T(0.5 * 1.8378770664093454836) * T(nobs);
  }
};

// ---------------------------------------------------------------------------------------------
template <typename T, int D, bool M1>
__global__ void __launch_bounds__(32)
kalman_loglik_chain_kernel(KalmanArgs<T> g, T* __restrict__ out) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= g.B) return;
  FilterSink<T, D> sink;
  sink.init();
  int64_t nobs = 0;
  const T logw = kalman_walk<T, D, M1>(g, c, 0, g.Tn, sink, nobs);
  out[c] = sink.loglik(logw, nobs);
}

// ---------------------------------------------------------------------------------------------
// Parallel-in-time path.  Segment p of chain c covers steps [p*L, min((p+1)*L, T)).
// ---------------------------------------------------------------------------------------------
template <typename T, int D, bool M1>
__global__ void __launch_bounds__(128)
kalman_segment_summary_kernel(KalmanArgs<T> g, T* __restrict__ summaries, int64_t P, int64_t L) {
  const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t c = blockIdx.y;
  if (p >= P) return;
  ElemSink<T, D> sink;
  sink.init();
  const int64_t k0 = p * L;
  const int64_t k1 = (k0 + L < g.Tn) ? k0 + L : g.Tn;
  int64_t nobs = 0;
  const T logw = kalman_walk<T, D, M1>(g, c, k0, k1, sink, nobs);
  sink.finalize(logw, nobs);
  elem_store<T, D>(summaries + (c * P + p) * ScanElem<T, D>::N, sink.e);
}

// One block per chain.  Thread t owns the run of summaries [t*R, (t+1)*R); run totals are scanned
// serially by thread 0 in shared memory (NT <= 128 combines), then every thread walks its run.
//   seeds [B,P,D+D*D]: filtered (mean, cov) at the step just before segment p (undefined for p = 0
//   when there is no incoming prefix);  total [B,N] (optional): join of all P summaries (local).
template <typename T, int D, int NT>
__global__ void __launch_bounds__(NT)
kalman_summary_scan_kernel(const T* __restrict__ summaries, const T* __restrict__ prefix_in,
                           T* __restrict__ seeds, T* __restrict__ total, int64_t P) {
  constexpr int N = ScanElem<T, D>::N, DD = D * D;
  __shared__ T run_tot[NT * N];
  const int64_t c = blockIdx.x;
  const int t = threadIdx.x;
  const int64_t R = (P + NT - 1) / NT;
  const int64_t p0 = t * R;
  const int64_t p1 = (p0 + R < P) ? p0 + R : P;
  const T* sp = summaries + c * P * N;
  ScanElem<T, D> acc, nxt, tmp;
  if (p0 < p1) {
    elem_load<T, D>(acc, sp + p0 * N);
    for (int64_t p = p0 + 1; p < p1; ++p) {
      elem_load<T, D>(nxt, sp + p * N);
      elem_combine<T, D>(tmp, acc, nxt);
      acc = tmp;
    }
    elem_store<T, D>(run_tot + t * N, acc);
  }
  __syncthreads();
  const int nruns = (int)((P + R - 1) / R);
  if (t == 0) {
    // exclusive scan of the run totals in place: slot r (r >= 1) <- join of runs < r
    elem_load<T, D>(acc, run_tot);
    for (int r = 1; r < nruns; ++r) {
      elem_load<T, D>(nxt, run_tot + r * N);
      elem_store<T, D>(run_tot + r * N, acc);
      elem_combine<T, D>(tmp, acc, nxt);
      acc = tmp;
    }
    if (total) elem_store<T, D>(total + c * N, acc);  // join of all local summaries
  }
  __syncthreads();
  if (p0 < p1) {
    bool have = t > 0;
    if (have) elem_load<T, D>(acc, run_tot + t * N);
    if (prefix_in) {
      elem_load<T, D>(nxt, prefix_in + c * N);
      if (have) {
        elem_combine<T, D>(tmp, nxt, acc);
        acc = tmp;
      } else {
        acc = nxt;
        have = true;
      }
    }
    T* seedp = seeds + c * P * (D + DD);
    for (int64_t p = p0; p < p1; ++p) {
      if (have) {
        T* s = seedp + p * (D + DD);
#pragma unroll
        for (int i = 0; i < D; ++i) s[i] = acc.b[i];
#pragma unroll
        for (int i = 0; i < DD; ++i) s[D + i] = acc.C[i];
      }
      if (p + 1 < p1) {
        elem_load<T, D>(nxt, sp + p * N);
        if (have) {
          elem_combine<T, D>(tmp, acc, nxt);
          acc = tmp;
        } else {
          acc = nxt;
          have = true;
        }
      }
    }
  }
}

template <typename T, int D, bool M1>
__global__ void __launch_bounds__(128)
kalman_seeded_filter_kernel(KalmanArgs<T> g, const T* __restrict__ seeds, T* __restrict__ partial,
                            int64_t P, int64_t L, int have_prefix) {
  constexpr int DD = D * D;
  const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t c = blockIdx.y;
  if (p >= P) return;
  FilterSink<T, D> sink;
  sink.init();
  const int64_t k0 = p * L;
  const int64_t k1 = (k0 + L < g.Tn) ? k0 + L : g.Tn;
  if (p > 0 || have_prefix) {
    const T* s = seeds + (c * P + p) * (D + DD);
#pragma unroll
    for (int i = 0; i < D; ++i) sink.st.m[i] = s[i];
#pragma unroll
    for (int i = 0; i < DD; ++i) sink.st.P[i] = s[D + i];
  }
  int64_t nobs = 0;
  const T logw = kalman_walk<T, D, M1>(g, c, k0, k1, sink, nobs);
  partial[c * P + p] = sink.loglik(logw, nobs);
}

template <typename T>
__global__ void __launch_bounds__(256)
kalman_partial_sum_kernel(const T* __restrict__ partial, T* __restrict__ out, int64_t P) {
  const int64_t c = blockIdx.x;
  T acc = T(0);
  for (int64_t i = threadIdx.x; i < P; i += blockDim.x) acc += partial[c * P + i];
  __shared__ T red[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    T s = T(0);
    for (int w = 0; w < 8; ++w) s += red[w];
    out[c] = s;
  }
}

// Join n elements per chain in order: elems [n,B,N] -> out [B,N]  (folds the all-gathered segment
// summaries of the earlier GPU ranks into this rank's incoming prefix).
template <typename T, int D>
__global__ void __launch_bounds__(32)
kalman_fold_elements_kernel(const T* __restrict__ elems, T* __restrict__ out, int64_t n, int64_t B) {
  constexpr int N = ScanElem<T, D>::N;
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= B) return;
  ScanElem<T, D> acc, nxt, tmp;
  elem_load<T, D>(acc, elems + c * N);
  for (int64_t i = 1; i < n; ++i) {
    elem_load<T, D>(nxt, elems + (i * B + c) * N);
    elem_combine<T, D>(tmp, acc, nxt);
    acc = tmp;
  }
  elem_store<T, D>(out + c * N, acc);
}

}  // namespace mf
