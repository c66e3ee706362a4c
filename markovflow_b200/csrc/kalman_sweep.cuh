// Kalman log-likelihood on the TMA chain sweep (sweep.cuh), output_dim m = 1.
//
// A "chain" of the sweep is a VIRTUAL chain: segment p of series c, i.e. steps [p*L, p*L+L) of a
// real series (P = 1, L = T: ordinary batched filtering).  Few long series are cut into many
// virtual chains and evaluated parallel-in-time:
//   1. KalmanSummaryCore : every virtual chain reduces its steps to one scan element
//   2. kalman_block_scan_kernel / kalman_top_scan_kernel : exclusive prefix over the P elements of
//      each series (warp-shuffle scans of 3D^2+2D-double elements; the identity element pads)
//   3. KalmanFilterCore<SEEDED> : every virtual chain filters its steps from the prefix state and
//      emits its share of the log-likelihood;  kalman_partial_sum_kernel adds the shares.
// Streams per step: A (D*D), b (D), chol_q (D*D), h (D), y (1) [, chol_r (1) when per-step].
#pragma once
#include "kalman_kernels.cuh"
#include "sweep.cuh"

namespace mf {

template <typename T>
struct KalmanSweepParams {
  KalmanArgs<T> g;
  int64_t P, L;            // segment slots per series (multiple of 32 when > 1), steps per segment
  int reduce_warp;         // summaries: emit one element per 32 consecutive segments (out [B,P/32,N])
  const T* local_prefix;   // [B,P,N]   exclusive prefix of the summaries inside their scan block
  const T* block_prefix;   // [B,nblk,N] exclusive prefix of the scan-block aggregates (+ incoming)
  int64_t nblk, scan_nt;
  int have_prefix;         // an incoming prefix exists (segment of a longer series)
  T* out;                  // summaries [B,P,N] | log-likelihood shares [B,P]
};

template <typename T, int D>
__device__ __forceinline__ void elem_shfl_up(ScanElem<T, D>& dst, const ScanElem<T, D>& src, int delta) {
  constexpr int DD = D * D;
#pragma unroll
  for (int i = 0; i < DD; ++i) {
    dst.A[i] = __shfl_up_sync(0xffffffffu, src.A[i], delta);
    dst.C[i] = __shfl_up_sync(0xffffffffu, src.C[i], delta);
    dst.J[i] = __shfl_up_sync(0xffffffffu, src.J[i], delta);
  }
#pragma unroll
  for (int i = 0; i < D; ++i) {
    dst.b[i] = __shfl_up_sync(0xffffffffu, src.b[i], delta);
    dst.eta[i] = __shfl_up_sync(0xffffffffu, src.eta[i], delta);
  }
  dst.ell = __shfl_up_sync(0xffffffffu, src.ell, delta);
}

template <typename T_, int D, bool TVR>
struct KalmanCoreBase {
  using T = T_;
  using Params = KalmanSweepParams<T>;
  static constexpr int DD = D * D;
  static constexpr int NIN = TVR ? 6 : 5, NOUT = 0;
  static constexpr bool BACKWARD = false;
  static constexpr int ein(int i) { return (i == 0 || i == 2) ? DD : ((i == 1 || i == 3) ? D : 1); }
  static constexpr int eout(int) { return 1; }
  static __device__ __forceinline__ int64_t num_chains(const Params& p) { return p.g.B * p.P; }
  static __device__ __forceinline__ int64_t max_steps(const Params& p) { return p.L; }
  static __device__ __forceinline__ StreamGeom out_geom(const Params&, int, int64_t) {
    return StreamGeom{nullptr, 0, 0};
  }
  static __device__ __forceinline__ StreamGeom in_geom(const Params& p, int i, int64_t chain) {
    const int64_t c = chain / p.P, seg = chain % p.P;
    const int64_t k0 = seg * p.L;
    int64_t steps = p.g.Tn - k0;
    if (steps > p.L) steps = p.L;
    if (steps < 0) steps = 0;
    const int fi = p.g.first_is_initial;
    const int64_t nt = p.g.Tn - fi;
    constexpr int ES = (int)sizeof(T);
    StreamGeom sg;
    sg.end = steps;
    sg.first = 0;
    if (i <= 2) {  // transition leading INTO step k is entry k - fi
      const T* base = (i == 0) ? p.g.a : (i == 1 ? p.g.b : p.g.chol_q);
      const int E = (i == 1) ? D : DD;
      sg.step0 = (char*)const_cast<T*>(base) + (c * nt + k0 - fi) * (int64_t)(E * ES);
      sg.first = (fi && k0 == 0) ? 1 : 0;
    } else if (i == 3) {
      sg.step0 = (char*)const_cast<T*>(p.g.h) + ((p.g.Bh == 1 ? 0 : c) * p.g.Tn + k0) * (int64_t)(D * ES);
    } else if (i == 4) {
      sg.step0 = (char*)const_cast<T*>(p.g.obs) + (c * p.g.Tn + k0) * (int64_t)ES;
    } else {
      sg.step0 = (char*)const_cast<T*>(p.g.chol_r) + k0 * (int64_t)ES;
    }
    return sg;
  }

  int64_t k0_;     // global step of local step 0
  int64_t steps_;  // steps of this virtual chain
  T w1_;           // 1 / chol_r (shared noise)
  LogProd<T> wdet_;  // product of the per-step whiteners (TVR only)
  int nobs_;
  bool prior_start_;

  // sum of log whiteners of the absorbed observations
  __device__ __forceinline__ T log_whiteners() const {
    if (TVR) return wdet_.log_abs();
    return T(nobs_) * Num<T>::log(Num<T>::abs(w1_));
  }

  __device__ __forceinline__ void init_base(const Params& p, int64_t chain) {
    const int64_t seg = chain % p.P;
    k0_ = seg * p.L;
    steps_ = p.g.Tn - k0_;
    if (steps_ > p.L) steps_ = p.L;
    if (steps_ < 0) steps_ = 0;  // padding slot past the end of the series: identity element
    w1_ = TVR ? T(1) : Num<T>::rcp(p.g.chol_r[0]);
    wdet_.init();
    nobs_ = 0;
    prior_start_ = p.g.first_is_initial && k0_ == 0;
  }

  // feeds local steps j0 .. j0+ns-1 of the tile to the sink
  template <class Sink>
  __device__ __forceinline__ void walk_tile(const Params& p, const T* const* in, int64_t j0, int ns,
                                            int64_t chain, Sink& sink) {
    int n = ns;
    if (j0 + n > steps_) n = (int)(steps_ - j0);
    if (n <= 0) return;
    // Records are prefetched one step ahead into two ping-pong register sets (shared-memory latency
    // off the dependent chain, no register-to-register copies).
    struct Rec {
      T F[DD], u[D], Lq[DD], h[D], y, r;
    };
    auto fetch = [&](Rec& rec, int j) {
#pragma unroll
      for (int i = 0; i < DD; ++i) { rec.F[i] = in[0][j * DD + i]; rec.Lq[i] = in[2][j * DD + i]; }
#pragma unroll
      for (int i = 0; i < D; ++i) { rec.u[i] = in[1][j * D + i]; rec.h[i] = in[3][j * D + i]; }
      rec.y = in[4][j];
      rec.r = TVR ? in[5][j] : T(1);
    };
    auto step = [&](Rec& rec, bool from_prior) {
      if (from_prior) {
        const int64_t c = chain / p.P;
        T mu[D], L0[DD];
        load_vec<T, D>(mu, p.g.mu0 + c * D);
        load_vec<T, DD>(L0, p.g.chol_p0 + c * DD);
        sink.start_prior(mu, L0);
      } else {
        sink.transition(rec.F, rec.u, rec.Lq);
      }
      T w = w1_;
      if (TVR) w = Num<T>::rcp(rec.r);
      if (!TVR || w != T(0)) {  // per-step noise: an infinite scale marks a step without observation
#pragma unroll
        for (int i = 0; i < D; ++i) rec.h[i] *= w;
        if (TVR) wdet_.mul(w);
        sink.absorb(rec.h, rec.y * w);
        ++nobs_;
      }
    };
    Rec ra, rb;
    fetch(ra, 0);
    int j = 0;
    if (j0 == 0 && prior_start_) {  // the chain's very first step starts from the prior
      if (n > 1) fetch(rb, 1);
      step(ra, true);
      sink.det.peel();
      j = 1;
      if (n > 1) ra = rb;
    }
    for (; j + 1 < n; j += 2) {
      fetch(rb, j + 1);
      step(ra, false);
      if (j + 2 < n) fetch(ra, j + 2);
      step(rb, false);
      sink.det.peel();  // two pivots multiplied per renormalisation
    }
    if (j < n) {
      step(ra, false);
      sink.det.peel();
    }
  }
};

// ---- pass 1: range element of every virtual chain ---------------------------------------------
template <typename T_, int D, bool TVR>
struct KalmanSummaryCore : KalmanCoreBase<T_, D, TVR> {
  using Base = KalmanCoreBase<T_, D, TVR>;
  using T = T_;
  using Params = typename Base::Params;
  ElemSink<T, D> sink;
  int64_t chain_;
  __device__ __forceinline__ void init(const Params& p, int64_t chain) {
    this->init_base(p, chain);
    chain_ = chain;
    sink.init();
  }
  __device__ __forceinline__ void tile(const Params& p, const T* const* in, T* const*, int64_t j0, int ns) {
    this->walk_tile(p, in, j0, ns, chain_, sink);
  }
  __device__ __forceinline__ void finish(const Params& p, int64_t chain, bool valid) {
    if (valid) sink.finalize(this->log_whiteners(), this->nobs_); else sink.init();
    if (!p.reduce_warp) {
      if (valid) elem_store<T, D>(p.out + chain * ScanElem<T, D>::N, sink.e);
      return;
    }
    // the 32 segments of a warp are consecutive in time (P is a multiple of 32): join them here
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

// ---- pass 3 (or the whole job when P == 1): sequential filter, optionally seeded ---------------
template <typename T_, int D, bool TVR, bool SEEDED>
struct KalmanFilterCore : KalmanCoreBase<T_, D, TVR> {
  using Base = KalmanCoreBase<T_, D, TVR>;
  using T = T_;
  using Params = typename Base::Params;
  FilterSink<T, D> sink;
  int64_t chain_;
  __device__ __forceinline__ void init(const Params& p, int64_t chain) {
    this->init_base(p, chain);
    chain_ = chain;
    sink.init();
    if (SEEDED) {
      constexpr int N = ScanElem<T, D>::N;
      const int64_t c = chain / p.P, seg = chain % p.P;
      if (seg > 0 || p.have_prefix) {
        // filtered state just before this segment: (b, C) of  block_prefix (+) local_prefix
        ScanElem<T, D> bp, lp, e;
        elem_load<T, D>(bp, p.block_prefix + (c * p.nblk + seg / p.scan_nt) * N);
        elem_load<T, D>(lp, p.local_prefix + chain * N);
        elem_combine<T, D>(e, bp, lp);
#pragma unroll
        for (int i = 0; i < D; ++i) sink.st.m[i] = e.b[i];
#pragma unroll
        for (int i = 0; i < D * D; ++i) sink.st.P[i] = e.C[i];
      }
    }
  }
  __device__ __forceinline__ void tile(const Params& p, const T* const* in, T* const*, int64_t j0, int ns) {
    this->walk_tile(p, in, j0, ns, chain_, sink);
  }
  __device__ __forceinline__ void finish(const Params& p, int64_t chain, bool valid) {
    if (valid) p.out[chain] = sink.loglik(this->log_whiteners(), this->nobs_);
  }
};

// ---- scans over range elements ----------------------------------------------------------------

// Inclusive warp scan (in time order: lower lanes are earlier).
template <typename T, int D>
__device__ __forceinline__ void elem_warp_scan(ScanElem<T, D>& e, int lane) {
  ScanElem<T, D> other, tmp;
#pragma unroll 1
  for (int delta = 1; delta < 32; delta <<= 1) {
    elem_shfl_up<T, D>(other, e, delta);
    if (lane >= delta) {
      elem_combine<T, D>(tmp, other, e);
      e = tmp;
    }
  }
}

// Exclusive scan of NT elements held one per thread (NT multiple of 32, <= 1024).
// Returns the thread's exclusive prefix in `exc`; `total` (valid in all threads) is the join of
// all NT elements.  smem: NT/32 elements of N values.
template <typename T, int D, int NT>
__device__ __forceinline__ void elem_block_exclusive_scan(ScanElem<T, D>& e, ScanElem<T, D>& exc,
                                                          ScanElem<T, D>& total, T* smem) {
  constexpr int N = ScanElem<T, D>::N, NW = NT / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  elem_warp_scan<T, D>(e, lane);  // inclusive within the warp
  if (lane == 31) elem_store<T, D>(smem + warp * N, e);
  __syncthreads();
  if (warp == 0) {
    ScanElem<T, D> w;
    if (lane < NW) elem_load<T, D>(w, smem + lane * N); else elem_identity<T, D>(w);
    elem_warp_scan<T, D>(w, lane);  // inclusive over warp totals
    if (lane < NW) elem_store<T, D>(smem + lane * N, w);
  }
  __syncthreads();
  ScanElem<T, D> up, tmp;
  elem_shfl_up<T, D>(up, e, 1);
  if (lane == 0) elem_identity<T, D>(up);  // exclusive within the warp
  if (warp > 0) {
    ScanElem<T, D> wp;
    elem_load<T, D>(wp, smem + (warp - 1) * N);
    elem_combine<T, D>(tmp, wp, up);
    exc = tmp;
  } else {
    exc = up;
  }
  elem_load<T, D>(total, smem + (NW - 1) * N);
}

// grid (nblk, B): scan block j of series c covers summaries [j*NT, (j+1)*NT).  In place: slot p of
// `elems` receives the exclusive prefix of element p INSIDE its block; block_agg [B,nblk,N] the
// block totals.
template <typename T, int D, int NT>
__global__ void __launch_bounds__(NT)
kalman_block_scan_kernel(T* __restrict__ elems, T* __restrict__ block_agg, int64_t P, int64_t nblk) {
  constexpr int N = ScanElem<T, D>::N;
  __shared__ T smem[(NT / 32) * N];
  const int64_t c = blockIdx.y, j = blockIdx.x;
  const int64_t p = j * NT + threadIdx.x;
  ScanElem<T, D> e, exc, total;
  if (p < P) elem_load<T, D>(e, elems + (c * P + p) * N); else elem_identity<T, D>(e);
  elem_block_exclusive_scan<T, D, NT>(e, exc, total, smem);
  if (p < P) elem_store<T, D>(elems + (c * P + p) * N, exc);
  if (threadIdx.x == 0) elem_store<T, D>(block_agg + (c * nblk + j) * N, total);
}

// grid (B): exclusive scan of the nblk block totals of one series, seeded with the incoming prefix
// (prefix_in [B,N] or NULL): block_prefix [B,nblk,N] (optional); total_out [B,N] (optional) = join
// of ALL block totals WITHOUT the incoming prefix (what the all-gather exchanges); ell_out [B]
// (optional) = log-normaliser of the join including the incoming prefix.
template <typename T, int D, int NT>
__global__ void __launch_bounds__(NT)
kalman_top_scan_kernel(const T* __restrict__ block_agg, const T* __restrict__ prefix_in,
                       T* __restrict__ block_prefix, T* __restrict__ total_out,
                       T* __restrict__ ell_out, int64_t nblk) {
  constexpr int N = ScanElem<T, D>::N;
  __shared__ T smem[(NT / 32) * N];
  const int64_t c = blockIdx.x;
  ScanElem<T, D> carry, carry_local, e, exc, total, tmp;
  if (prefix_in) elem_load<T, D>(carry, prefix_in + c * N); else elem_identity<T, D>(carry);
  elem_identity<T, D>(carry_local);
  for (int64_t base = 0; base < nblk; base += NT) {
    const int64_t j = base + threadIdx.x;
    if (j < nblk) elem_load<T, D>(e, block_agg + (c * nblk + j) * N); else elem_identity<T, D>(e);
    elem_block_exclusive_scan<T, D, NT>(e, exc, total, smem);
    if (j < nblk && block_prefix) {
      elem_combine<T, D>(tmp, carry, exc);
      elem_store<T, D>(block_prefix + (c * nblk + j) * N, tmp);
    }
    elem_combine<T, D>(tmp, carry, total);
    carry = tmp;
    if (total_out || ell_out) {
      elem_combine<T, D>(tmp, carry_local, total);
      carry_local = tmp;
    }
    __syncthreads();
  }
  if (total_out && threadIdx.x == 0) elem_store<T, D>(total_out + c * N, carry_local);
  // the ell of the join of a whole series (prior first) is its marginal log-likelihood
  if (ell_out && threadIdx.x == 0) ell_out[c] = prefix_in ? carry.ell : carry_local.ell;
}

// grid (B): ordered reduction of the P elements of one series; ell_out[c] = log-normaliser of the
// join = the marginal log-likelihood when the first element starts at the prior.  total_out [B,N]
// optional.  Thread t folds the run [t*r, (t+1)*r) sequentially, then warp-shuffle scans join the
// NT partial results in time order.
// Time-sharded series over several GPUs: the exchange of the per-rank elements happens INSIDE
// the reduction's tail.  Every rank owns a peer-mapped region (cudaIpc / NVLink P2P, mf_peer_*) laid out as
//   elements [2][world][B][N]  |  flags (uint64) [2][world][B]          (2 = call parity: double buffer)
// The thread that holds a chain's local total stores it into slot `rank` of EVERY rank's region, fences, raises
// the flags, then waits for the flags of all ranks in its own region and joins the `world` elements in rank
// (= time) order: no NCCL call, no extra launch, no host glue.  world == 1: plain reduction.
struct PeerExchange {
  void* region[8];        // region[r]: rank r's region as mapped into this process
  unsigned long long epoch;  // call counter, > 0, the same on every rank
  int rank, world;
  int64_t B;
};

template <typename T, int D>
__device__ __forceinline__ void peer_exchange_join(ScanElem<T, D>& acc, const PeerExchange& px, int64_t c) {
  constexpr int N = ScanElem<T, D>::N;
  const int par = (int)(px.epoch & 1ull);
  const size_t flag_off = sizeof(T) * (size_t)2 * px.world * px.B * N;
  T tmp[N];
  elem_store<T, D>(tmp, acc);
  const size_t slot = ((size_t)(par * px.world + px.rank) * px.B + c);
  for (int r = 0; r < px.world; ++r) {
    volatile T* dst = reinterpret_cast<volatile T*>(px.region[r]) + slot * N;
#pragma unroll
    for (int i = 0; i < N; ++i) dst[i] = tmp[i];
  }
  __threadfence_system();
  for (int r = 0; r < px.world; ++r) {
    volatile unsigned long long* f =
        reinterpret_cast<volatile unsigned long long*>(reinterpret_cast<char*>(px.region[r]) + flag_off) + slot;
    *f = px.epoch;
  }
  char* own = reinterpret_cast<char*>(px.region[px.rank]);
  ScanElem<T, D> tot, nxt, cmb;
  for (int r = 0; r < px.world; ++r) {
    const size_t s = ((size_t)(par * px.world + r) * px.B + c);
    volatile unsigned long long* f = reinterpret_cast<volatile unsigned long long*>(own + flag_off) + s;
    while (*f < px.epoch) {
    }
    __threadfence_system();
    const volatile T* src = reinterpret_cast<const volatile T*>(own) + s * N;
#pragma unroll
    for (int i = 0; i < N; ++i) tmp[i] = src[i];
    elem_load<T, D>(nxt, tmp);
    if (r == 0) {
      tot = nxt;
    } else {
      elem_combine<T, D>(cmb, tot, nxt);
      tot = cmb;
    }
  }
  acc = tot;
}

template <typename T, int D, int NT>
__global__ void __launch_bounds__(NT)
kalman_reduce_kernel(const T* __restrict__ elems, T* __restrict__ total_out,
                     T* __restrict__ ell_out, int64_t P, const PeerExchange px) {
  constexpr int N = ScanElem<T, D>::N, NW = NT / 32;
  __shared__ T smem[NW * N];
  const int64_t c = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (P == 1) {  // batches cut into one warp of segments per series: the warp join already is the total
    if (threadIdx.x == 0) {
      ScanElem<T, D> e;
      elem_load<T, D>(e, elems + c * N);
      if (px.world > 1) peer_exchange_join<T, D>(e, px, c);
      if (total_out) elem_store<T, D>(total_out + c * N, e);
      if (ell_out) ell_out[c] = e.ell;
    }
    return;
  }
  const int64_t r = (P + NT - 1) / NT;
  const int64_t p0 = threadIdx.x * r;
  const int64_t p1 = (p0 + r < P) ? p0 + r : P;
  const T* sp = elems + c * P * N;
  ScanElem<T, D> acc, nxt, tmp;
  elem_identity<T, D>(acc);
  if (p0 < p1) {
    elem_load<T, D>(acc, sp + p0 * N);
    for (int64_t p = p0 + 1; p < p1; ++p) {
      elem_load<T, D>(nxt, sp + p * N);
      if (D <= 2) elem_combine_inl<T, D>(tmp, acc, nxt); else elem_combine<T, D>(tmp, acc, nxt);
      acc = tmp;
    }
  }
  // inclusive warp scan (time order); lane 31 then holds the warp's join
#pragma unroll 1
  for (int delta = 1; delta < 32; delta <<= 1) {
    elem_shfl_up<T, D>(nxt, acc, delta);
    if (lane >= delta) {
      if (D <= 2) elem_combine_inl<T, D>(tmp, nxt, acc); else elem_combine<T, D>(tmp, nxt, acc);
      acc = tmp;
    }
  }
  if (lane == 31) elem_store<T, D>(smem + warp * N, acc);
  __syncthreads();
  if (warp == 0) {
    if (lane < NW) elem_load<T, D>(acc, smem + lane * N); else elem_identity<T, D>(acc);
#pragma unroll 1
    for (int delta = 1; delta < NW; delta <<= 1) {
      elem_shfl_up<T, D>(nxt, acc, delta);
      if (lane >= delta) {
        if (D <= 2) elem_combine_inl<T, D>(tmp, nxt, acc); else elem_combine<T, D>(tmp, nxt, acc);
        acc = tmp;
      }
    }
    if (lane == NW - 1) {
      if (px.world > 1) peer_exchange_join<T, D>(acc, px, c);
      if (total_out) elem_store<T, D>(total_out + c * N, acc);
      if (ell_out) ell_out[c] = acc.ell;
    }
  }
}

}  // namespace mf
