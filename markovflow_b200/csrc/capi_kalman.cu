// C-ABI entry points for the Kalman log-likelihood (include/markovflow_b200.h).
#include <cstring>

#include "dispatch.cuh"
#include "mid_api.h"
#include "kalman_kernels.cuh"
#include "kalman_sweep_api.h"

using namespace mf;

namespace {

struct Plan {
  bool pscan;
  int64_t P, L;
};

// Few long chains -> parallel-in-time (P segments of L steps per chain); many chains -> one thread
// per chain.  tuning knob 2: 0 auto, 1 sequential, 2 parallel-in-time; knob 3: L override.
Plan make_plan(int64_t B, int64_t T) {
  const int64_t target = 148 * 256;
  int64_t pmax = target / (B > 0 ? B : 1);
  if (pmax < 1) pmax = 1;
  int64_t L = (T + pmax - 1) / pmax;
  if (L < 16) L = 16;
  if (tuning(3) > 0) L = tuning(3);
  Plan p;
  p.L = L;
  p.P = (T + L - 1) / L;
  p.pscan = p.P >= 4;
  if (tuning(2) == 1) p.pscan = false;
  if (tuning(2) == 2) p.pscan = p.P >= 2;
  return p;
}

size_t align_up(size_t v) { return (v + 255) & ~size_t(255); }

struct Workspace {
  size_t summaries, seeds, partial, total;
};

// ---- TMA-sweep path (m = 1, D <= kKalmanSweepMaxD; kalman_sweep.cuh) ----------------------------
// tuning knob 4: 1 disables it (legacy direct-load kernels, kept for A/B measurements).
struct SweepPlan {
  bool use;
  int64_t P, L, nblk;
};

SweepPlan make_sweep_plan(int64_t B, int64_t T, int64_t D, int64_t m) {
  SweepPlan p;
  p.use = (m == 1 && D <= kKalmanSweepMaxD && tuning(4) != 1);
  p.P = 1; p.L = T; p.nblk = 1;
  if (!p.use) return p;
  const int64_t wave = 148 * (int64_t)kalman_sweep_chains_per_cta(D);
  int64_t ptarget = wave / (B > 0 ? B : 1);
  // segment slots come in whole warps (32), so cutting series pays only when a wave holds >= 32
  // segments per series; otherwise one chain per series
  if (tuning(2) == 1) return p;
  if (tuning(2) == 0 && ptarget < 32) {
    // A batch that leaves most SMs without a CTA when every series is one chain (1024 series = 16
    // CTAs) is still cut into 32 segments per series, as long as those keep >= 64 steps; from about
    // half a wave of series on, the cheaper plain filter with one chain per series wins.
    if (B * 2 > wave || T < 32 * 64) return p;
  }
  if (ptarget < 32) ptarget = 32;
  int64_t L = (T + ptarget - 1) / ptarget;
  if (L < 64) L = 64;
  L = (L + 7) / 8 * 8;
  if (tuning(3) > 0) L = tuning(3);
  p.L = L;
  p.P = ((T + L - 1) / L + 31) / 32 * 32;  // segment slots, padded to whole warps (empty = identity)
  const int64_t nt = kalman_sweep_scan_threads(D);
  p.nblk = (p.P + nt - 1) / nt;
  return p;
}

struct SweepWs {
  size_t elems, block_agg, block_prefix, partial, total;
};

SweepWs sweep_layout(size_t es, int64_t B, const SweepPlan& pl, int64_t D) {
  const size_t N = 3 * D * D + 2 * D + 1;
  SweepWs w;
  w.elems = 0;
  w.block_agg = align_up(w.elems + es * B * pl.P * N);
  w.block_prefix = align_up(w.block_agg + es * B * pl.nblk * N);
  w.partial = align_up(w.block_prefix + es * B * pl.nblk * N);
  w.total = align_up(w.partial + es * B * pl.P);
  return w;
}

KalmanRawArgs raw_args(int dtype, const void* mu0, const void* chol_p0, const void* a,
                       const void* b, const void* chol_q, const void* h, const void* obs,
                       const void* chol_r, int64_t B, int64_t T, int64_t D, int64_t h_batch,
                       int64_t r_steps, int first_is_initial) {
  KalmanRawArgs r;
  r.dtype = dtype; r.mu0 = mu0; r.chol_p0 = chol_p0; r.a = a; r.b = b; r.chol_q = chol_q; r.h = h;
  r.obs = obs; r.chol_r = chol_r; r.B = B; r.T = T; r.D = D; r.h_batch = h_batch;
  r.r_steps = r_steps; r.first_is_initial = first_is_initial;
  return r;
}

// summaries -> in-place local prefixes + block aggregates
int sweep_summaries(const KalmanRawArgs& r, const SweepPlan& pl, const SweepWs& w, char* ws,
                    cudaStream_t s) {
  int rc = kalman_sweep_launch(1, r, pl.P, pl.L, nullptr, nullptr, pl.nblk, 0, ws + w.elems, s);
  if (rc != MF_OK) return rc;
  return kalman_sweep_block_scan(r.dtype, r.D, ws + w.elems, ws + w.block_agg, r.B, pl.P, pl.nblk, s);
}

int sweep_seeded(const KalmanRawArgs& r, const SweepPlan& pl, const SweepWs& w, char* ws,
                 const void* prefix_elem, void* out, cudaStream_t s) {
  int rc = kalman_sweep_top_scan(r.dtype, r.D, ws + w.block_agg, prefix_elem, ws + w.block_prefix,
                                 nullptr, nullptr, r.B, pl.nblk, s);
  if (rc != MF_OK) return rc;
  rc = kalman_sweep_launch(2, r, pl.P, pl.L, ws + w.elems, ws + w.block_prefix, pl.nblk,
                           prefix_elem != nullptr, ws + w.partial, s);
  if (rc != MF_OK) return rc;
  return dispatch_dtype(r.dtype, [&](auto tt) {
    using Tp = typename decltype(tt)::type;
    kalman_partial_sum_kernel<Tp><<<(unsigned)r.B, 256, 0, s>>>((const Tp*)(ws + w.partial), (Tp*)out, pl.P);
    return check_launch();
  });
}

Workspace layout(size_t es, int64_t B, int64_t P, int64_t D) {
  const size_t N = 3 * D * D + 2 * D + 1;
  Workspace w;
  w.summaries = 0;
  w.seeds = align_up(w.summaries + es * B * P * N);
  w.partial = align_up(w.seeds + es * B * P * (D + D * D));
  w.total = align_up(w.partial + es * B * P);
  return w;
}

template <typename T>
KalmanArgs<T> make_args(const void* mu0, const void* chol_p0, const void* a, const void* b,
                        const void* chol_q, const void* h, const void* obs, const void* chol_r,
                        int64_t B, int64_t Tn, int64_t m, int64_t h_batch, int64_t r_steps,
                        int first_is_initial) {
  KalmanArgs<T> g;
  g.mu0 = (const T*)mu0; g.chol_p0 = (const T*)chol_p0; g.a = (const T*)a; g.b = (const T*)b;
  g.chol_q = (const T*)chol_q; g.h = (const T*)h; g.obs = (const T*)obs; g.chol_r = (const T*)chol_r;
  g.B = B; g.Tn = Tn; g.Bh = h_batch; g.Tr = r_steps; g.m = (int)m;
  g.first_is_initial = first_is_initial;
  return g;
}

int check_args(const void* mu0, const void* chol_p0, const void* a, const void* b,
               const void* chol_q, const void* h, const void* obs, const void* chol_r, int64_t B,
               int64_t T, int64_t D, int64_t m, int64_t h_batch, int64_t r_steps,
               int first_is_initial) {
  if (B < 0 || T < 1 || D < 1 || m < 1) return MF_ERR_BAD_ARG;
  if (m > kMaxObsDim) return MF_ERR_UNSUPPORTED;
  if (!h || !obs || !chol_r) return MF_ERR_BAD_ARG;
  if (first_is_initial && (!mu0 || !chol_p0)) return MF_ERR_BAD_ARG;
  if ((T - (first_is_initial ? 1 : 0)) > 0 && (!a || !b || !chol_q)) return MF_ERR_BAD_ARG;
  if ((h_batch != 1 && h_batch != B) || (r_steps != 1 && r_steps != T)) return MF_ERR_BAD_ARG;
  return MF_OK;
}

template <int D> struct ScanThreads { static constexpr int value = D <= 3 ? 128 : (D <= 5 ? 64 : (D <= 7 ? 32 : 16)); };

template <typename T, int D, bool M1>
int run_summaries(const KalmanArgs<T>& g, const Plan& pl, T* summaries, cudaStream_t s) {
  dim3 grid(grid_for(pl.P, 128), (unsigned)g.B);
  kalman_segment_summary_kernel<T, D, M1><<<grid, 128, 0, s>>>(g, summaries, pl.P, pl.L);
  return check_launch();
}

}  // namespace

extern "C" {

size_t mf_kalman_workspace_bytes(int dtype, int64_t B, int64_t T, int64_t D) {
  if (B < 1 || T < 1 || D < 1) return 0;
  const Plan pl = make_plan(B, T);
  const size_t es = dtype == MF_F64 ? 8 : 4;
  const size_t N = 3 * D * D + 2 * D + 1;
  size_t legacy = layout(es, B, pl.P, D).total + align_up(es * B * N);
  if (D <= kKalmanSweepMaxD) {
    const size_t sweep = sweep_layout(es, B, make_sweep_plan(B, T, D, 1), D).total;
    if (sweep > legacy) legacy = sweep;
  }
  return legacy;
}

}  // extern "C"

namespace {

int fill_peers(KalmanPeerArgs& pa, void* const* regions, int rank, int world, uint64_t epoch) {
  if (world < 1 || world > 8 || rank < 0 || rank >= world || epoch == 0 || !regions) return MF_ERR_BAD_ARG;
  pa.rank = rank; pa.world = world; pa.epoch = epoch;
  for (int r = 0; r < 8; ++r) pa.region[r] = r < world ? regions[r] : nullptr;
  for (int r = 0; r < world; ++r)
    if (!pa.region[r]) return MF_ERR_BAD_ARG;
  return MF_OK;
}

int segment_summary_impl(int dtype, const void* mu0, const void* chol_p0, const void* a,
                         const void* b, const void* chol_q, const void* h, const void* obs,
                         const void* chol_r, void* out_elem, int64_t B, int64_t T, int64_t D,
                         int64_t m, int64_t h_batch, int64_t r_steps, int first_is_initial,
                         void* workspace, size_t workspace_bytes, void* stream, const KalmanPeerArgs* peers,
                         void* out_ell);

}  // namespace

extern "C" {

int mf_kalman_segment_summary(int dtype, const void* mu0, const void* chol_p0, const void* a,
                              const void* b, const void* chol_q, const void* h, const void* obs,
                              const void* chol_r, void* out_elem, int64_t B, int64_t T, int64_t D,
                              int64_t m, int64_t h_batch, int64_t r_steps, int first_is_initial,
                              void* workspace, size_t workspace_bytes, void* stream) {
  return segment_summary_impl(dtype, mu0, chol_p0, a, b, chol_q, h, obs, chol_r, out_elem, B, T, D, m, h_batch,
                              r_steps, first_is_initial, workspace, workspace_bytes, stream, nullptr, nullptr);
}

int mf_kalman_time_sharded_log_likelihood(int dtype, const void* mu0, const void* chol_p0, const void* a,
                                          const void* b, const void* chol_q, const void* h, const void* obs,
                                          const void* chol_r, void* out, void* out_elem, int64_t B, int64_t T,
                                          int64_t D, int64_t m, int64_t h_batch, int64_t r_steps,
                                          int first_is_initial, void* const* peer_regions, int rank, int world,
                                          uint64_t epoch, void* workspace, size_t workspace_bytes, void* stream) {
  if (!out || !out_elem) return MF_ERR_BAD_ARG;
  KalmanPeerArgs pa;
  const int rc = fill_peers(pa, peer_regions, rank, world, epoch);
  if (rc != MF_OK) return rc;
  return segment_summary_impl(dtype, mu0, chol_p0, a, b, chol_q, h, obs, chol_r, out_elem, B, T, D, m, h_batch,
                              r_steps, first_is_initial, workspace, workspace_bytes, stream, &pa, out);
}

size_t mf_kalman_peer_region_bytes(int dtype, int64_t B, int64_t D, int world) {
  if (B < 1 || D < 1 || world < 1) return 0;
  const size_t es = dtype == MF_F64 ? 8 : 4;
  const size_t N = 3 * D * D + 2 * D + 1;
  return es * 2 * (size_t)world * B * N + 8 * 2 * (size_t)world * B;
}

int mf_peer_alloc(size_t bytes, void** ptr, unsigned char* handle64) {
  if (!bytes || !ptr || !handle64) return MF_ERR_BAD_ARG;
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) return check_launch();
  if (cudaMemset(p, 0, bytes) != cudaSuccess) return check_launch();
  cudaIpcMemHandle_t hd;
  if (cudaIpcGetMemHandle(&hd, p) != cudaSuccess) { cudaFree(p); return check_launch(); }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &hd, 64);
  *ptr = p;
  return MF_OK;
}

int mf_peer_open(const unsigned char* handle64, void** ptr) {
  if (!handle64 || !ptr) return MF_ERR_BAD_ARG;
  cudaIpcMemHandle_t hd;
  memcpy(&hd, handle64, 64);
  if (cudaIpcOpenMemHandle(ptr, hd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) return check_launch();
  return MF_OK;
}

int mf_peer_close(void* ptr) {
  if (!ptr) return MF_ERR_BAD_ARG;
  if (cudaIpcCloseMemHandle(ptr) != cudaSuccess) return check_launch();
  return MF_OK;
}

int mf_peer_free(void* ptr) {
  if (!ptr) return MF_ERR_BAD_ARG;
  if (cudaFree(ptr) != cudaSuccess) return check_launch();
  return MF_OK;
}

}  // extern "C"

namespace {

int segment_summary_impl(int dtype, const void* mu0, const void* chol_p0, const void* a,
                         const void* b, const void* chol_q, const void* h, const void* obs,
                         const void* chol_r, void* out_elem, int64_t B, int64_t T, int64_t D,
                         int64_t m, int64_t h_batch, int64_t r_steps, int first_is_initial,
                         void* workspace, size_t workspace_bytes, void* stream, const KalmanPeerArgs* peers,
                         void* out_ell) {
  int st = check_args(mu0, chol_p0, a, b, chol_q, h, obs, chol_r, B, T, D, m, h_batch, r_steps,
                      first_is_initial);
  if (st != MF_OK) return st;
  if (B == 0) return MF_OK;
  if (!out_elem || !workspace || B > 65535) return MF_ERR_BAD_ARG;
  if (workspace_bytes < mf_kalman_workspace_bytes(dtype, B, T, D)) return MF_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  const SweepPlan sp = make_sweep_plan(B, T, D, m);
  if (sp.use) {
    if (dtype != MF_F32 && dtype != MF_F64) return MF_ERR_BAD_ARG;
    const SweepWs w = sweep_layout(dtype == MF_F64 ? 8 : 4, B, sp, D);
    char* ws = (char*)workspace;
    const KalmanRawArgs r = raw_args(dtype, mu0, chol_p0, a, b, chol_q, h, obs, chol_r, B, T, D,
                                     h_batch, r_steps, first_is_initial);
    // one pass: per-warp joins of the segment elements, then an ordered reduction
    if (sp.P == 1) {
      int rc = kalman_sweep_launch(1, r, 1, T, nullptr, nullptr, 1, 0, out_elem, s, 0);
      if (rc != MF_OK || !peers) return rc;
      return kalman_sweep_reduce(dtype, D, out_elem, out_elem, out_ell, B, 1, s, peers);
    }
    int rc = kalman_sweep_launch(1, r, sp.P, sp.L, nullptr, nullptr, sp.nblk, 0, ws + w.elems, s, 1);
    if (rc != MF_OK) return rc;
    return kalman_sweep_reduce(dtype, D, ws + w.elems, out_elem, out_ell, B, sp.P / 32, s, peers);
  }
  if (peers) return MF_ERR_UNSUPPORTED;  // the fused exchange lives in the sweep path (m = 1, D <= 4)
  return dispatch_small(dtype, D, [&](auto tt, auto dd) {
    using Tp = typename decltype(tt)::type;
    constexpr int kD = decltype(dd)::value;
    const Plan pl = make_plan(B, T);
    const Workspace w = layout(sizeof(Tp), B, pl.P, kD);
    char* ws = (char*)workspace;
    const auto g = make_args<Tp>(mu0, chol_p0, a, b, chol_q, h, obs, chol_r, B, T, m, h_batch,
                                 r_steps, first_is_initial);
    int rc = (m == 1) ? run_summaries<Tp, kD, true>(g, pl, (Tp*)(ws + w.summaries), s)
                      : run_summaries<Tp, kD, false>(g, pl, (Tp*)(ws + w.summaries), s);
    if (rc != MF_OK) return rc;
    kalman_summary_scan_kernel<Tp, kD, ScanThreads<kD>::value>
        <<<(unsigned)B, ScanThreads<kD>::value, 0, s>>>((const Tp*)(ws + w.summaries), nullptr,
                                                       (Tp*)(ws + w.seeds), (Tp*)out_elem, pl.P);
    return check_launch();
  });
}

}  // namespace

extern "C" {

int mf_kalman_fold_elements(int dtype, const void* elems, void* out, int64_t n, int64_t B,
                            int64_t D, void* stream) {
  if (n < 1 || B < 0 || D < 1 || !elems || !out) return MF_ERR_BAD_ARG;
  if (B == 0) return MF_OK;
  cudaStream_t s = (cudaStream_t)stream;
  return dispatch_small(dtype, D, [&](auto tt, auto dd) {
    using Tp = typename decltype(tt)::type;
    constexpr int kD = decltype(dd)::value;
    kalman_fold_elements_kernel<Tp, kD><<<grid_for(B, 32), 32, 0, s>>>((const Tp*)elems, (Tp*)out, n, B);
    return check_launch();
  });
}

int mf_kalman_log_likelihood_seeded(int dtype, const void* mu0, const void* chol_p0, const void* a,
                                    const void* b, const void* chol_q, const void* h,
                                    const void* obs, const void* chol_r, const void* prefix_elem,
                                    void* out, int64_t B, int64_t T, int64_t D, int64_t m,
                                    int64_t h_batch, int64_t r_steps, int first_is_initial,
                                    int summaries_valid, void* workspace, size_t workspace_bytes,
                                    void* stream) {
  int st = check_args(mu0, chol_p0, a, b, chol_q, h, obs, chol_r, B, T, D, m, h_batch, r_steps,
                      first_is_initial);
  if (st != MF_OK) return st;
  if (B == 0) return MF_OK;
  if (!out || !workspace || B > 65535) return MF_ERR_BAD_ARG;
  if (!first_is_initial && !prefix_elem) return MF_ERR_BAD_ARG;
  if (workspace_bytes < mf_kalman_workspace_bytes(dtype, B, T, D)) return MF_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  const SweepPlan sp = make_sweep_plan(B, T, D, m);
  if (sp.use) {
    if (dtype != MF_F32 && dtype != MF_F64) return MF_ERR_BAD_ARG;
    const SweepWs w = sweep_layout(dtype == MF_F64 ? 8 : 4, B, sp, D);
    char* ws = (char*)workspace;
    const KalmanRawArgs r = raw_args(dtype, mu0, chol_p0, a, b, chol_q, h, obs, chol_r, B, T, D,
                                     h_batch, r_steps, first_is_initial);
    // (mf_kalman_segment_summary leaves only per-warp joins behind: always redo the summaries)
    int rc = sweep_summaries(r, sp, w, ws, s);
    if (rc != MF_OK) return rc;
    return sweep_seeded(r, sp, w, ws, prefix_elem, out, s);
  }
  return dispatch_small(dtype, D, [&](auto tt, auto dd) {
    using Tp = typename decltype(tt)::type;
    constexpr int kD = decltype(dd)::value;
    const Plan pl = make_plan(B, T);
    const Workspace w = layout(sizeof(Tp), B, pl.P, kD);
    char* ws = (char*)workspace;
    const auto g = make_args<Tp>(mu0, chol_p0, a, b, chol_q, h, obs, chol_r, B, T, m, h_batch,
                                 r_steps, first_is_initial);
    int rc = MF_OK;
    if (!summaries_valid) {
      rc = (m == 1) ? run_summaries<Tp, kD, true>(g, pl, (Tp*)(ws + w.summaries), s)
                    : run_summaries<Tp, kD, false>(g, pl, (Tp*)(ws + w.summaries), s);
      if (rc != MF_OK) return rc;
    }
    kalman_summary_scan_kernel<Tp, kD, ScanThreads<kD>::value>
        <<<(unsigned)B, ScanThreads<kD>::value, 0, s>>>((const Tp*)(ws + w.summaries),
                                                       (const Tp*)prefix_elem, (Tp*)(ws + w.seeds),
                                                       nullptr, pl.P);
    if ((rc = check_launch()) != MF_OK) return rc;
    dim3 grid(grid_for(pl.P, 128), (unsigned)B);
    const int have_prefix = prefix_elem != nullptr;
    if (m == 1)
      kalman_seeded_filter_kernel<Tp, kD, true><<<grid, 128, 0, s>>>(
          g, (const Tp*)(ws + w.seeds), (Tp*)(ws + w.partial), pl.P, pl.L, have_prefix);
    else
      kalman_seeded_filter_kernel<Tp, kD, false><<<grid, 128, 0, s>>>(
          g, (const Tp*)(ws + w.seeds), (Tp*)(ws + w.partial), pl.P, pl.L, have_prefix);
    if ((rc = check_launch()) != MF_OK) return rc;
    kalman_partial_sum_kernel<Tp><<<(unsigned)B, 256, 0, s>>>((const Tp*)(ws + w.partial), (Tp*)out, pl.P);
    return check_launch();
  });
}

int mf_kalman_log_likelihood(int dtype, const void* mu0, const void* chol_p0, const void* a,
                             const void* b, const void* chol_q, const void* h, const void* obs,
                             const void* chol_r, void* out, int64_t B, int64_t T, int64_t D,
                             int64_t m, int64_t h_batch, int64_t r_steps, void* workspace,
                             size_t workspace_bytes, void* stream) {
  int st = check_args(mu0, chol_p0, a, b, chol_q, h, obs, chol_r, B, T, D, m, h_batch, r_steps, 1);
  if (st != MF_OK) return st;
  if (B == 0) return MF_OK;
  if (!out) return MF_ERR_BAD_ARG;
  if (mid_dim(D))
    return mid_kalman_log_likelihood(dtype, mu0, chol_p0, a, b, chol_q, h, obs, chol_r, out, B, T, D, m, h_batch,
                                     r_steps, (cudaStream_t)stream);
  const SweepPlan sp = make_sweep_plan(B, T, D, m);
  if (sp.use) {
    if (dtype != MF_F32 && dtype != MF_F64) return MF_ERR_BAD_ARG;
    const bool ws_ok = workspace && B <= 65535 &&
                       workspace_bytes >= mf_kalman_workspace_bytes(dtype, B, T, D);
    if (sp.P > 1 && ws_ok) {
      // parallel in time, ONE pass over the data: per-segment elements, then an ordered reduction
      // whose ell component is the log-likelihood
      const SweepWs w = sweep_layout(dtype == MF_F64 ? 8 : 4, B, sp, D);
      char* ws = (char*)workspace;
      const KalmanRawArgs r = raw_args(dtype, mu0, chol_p0, a, b, chol_q, h, obs, chol_r, B, T, D,
                                       h_batch, r_steps, 1);
      int rc = kalman_sweep_launch(1, r, sp.P, sp.L, nullptr, nullptr, sp.nblk, 0, ws + w.elems,
                                   (cudaStream_t)stream, 1);
      if (rc != MF_OK) return rc;
      return kalman_sweep_reduce(dtype, D, ws + w.elems, nullptr, out, B, sp.P / 32,
                                 (cudaStream_t)stream);
    }
    // batched filter: one virtual chain per series (also the no-workspace case)
    const KalmanRawArgs r = raw_args(dtype, mu0, chol_p0, a, b, chol_q, h, obs, chol_r, B, T, D,
                                     h_batch, r_steps, 1);
    return kalman_sweep_launch(0, r, 1, T, nullptr, nullptr, 1, 0, out, (cudaStream_t)stream);
  }
  const Plan pl = make_plan(B, T);
  if (pl.pscan && workspace && B <= 65535 &&
      workspace_bytes >= mf_kalman_workspace_bytes(dtype, B, T, D))
    return mf_kalman_log_likelihood_seeded(dtype, mu0, chol_p0, a, b, chol_q, h, obs, chol_r,
                                           nullptr, out, B, T, D, m, h_batch, r_steps, 1, 0,
                                           workspace, workspace_bytes, stream);
  cudaStream_t s = (cudaStream_t)stream;
  return dispatch_small(dtype, D, [&](auto tt, auto dd) {
    using Tp = typename decltype(tt)::type;
    constexpr int kD = decltype(dd)::value;
    const auto g = make_args<Tp>(mu0, chol_p0, a, b, chol_q, h, obs, chol_r, B, T, m, h_batch,
                                 r_steps, 1);
    if (m == 1)
      kalman_loglik_chain_kernel<Tp, kD, true><<<grid_for(B, 32), 32, 0, s>>>(g, (Tp*)out);
    else
      kalman_loglik_chain_kernel<Tp, kD, false><<<grid_for(B, 32), 32, 0, s>>>(g, (Tp*)out);
    return check_launch();
  });
}

}  // extern "C"
