// C-ABI entry points of the Matern-prior Kalman log-likelihood with in-kernel SSM construction
// (kalman_sde_sweep.cuh).
#include "dispatch.cuh"
#include "kalman_sde_sweep.cuh"
#include "kalman_sweep_api.h"

using namespace mf;

namespace {

size_t align_up(size_t v) { return (v + 255) & ~size_t(255); }

// Segments per series.  The recursion is arithmetic-bound (two values per step cross HBM), so the
// series are cut until every SM holds many compute warps.
// tuning knob 8: 0 = direct kernel (every warp computes), 1..4 = TMA chain-sweep geometries, 5 / 6 =
// direct kernel capped at the registers of 5 / 6 CTAs per SM (all measured slower);
// knob 9: virtual chains per SM aimed at, in warps (0 auto); knob 3: steps per segment override.
struct SdePlan {
  int64_t P, L;
};

SdePlan make_sde_plan(int64_t B, int64_t T) {
  SdePlan p;
  p.P = 1;
  p.L = T;
  const int warps_per_sm = tuning(9) > 0 ? tuning(9) : ((tuning(8) > 0 && tuning(8) < 5) ? 8 : 12);
  const int64_t wave = (int64_t)148 * 32 * warps_per_sm;
  int64_t ptarget = wave / (B > 0 ? B : 1);
  if (tuning(2) == 1) return p;
  if (ptarget < 32) {
    // segment slots come in whole warps (the join is a warp scan).  A batch that cannot fill the GPU
    // with one thread per series (4096 series = 128 warps on 148 SMs: 1.95 ms for T = 1e4, latency
    // bound) is still cut into 32 segments per series as long as those have >= 64 steps each.
    if (B * 4 > wave || T < 32 * 64) return p;  // enough series (or too short): one chain per series
    ptarget = 32;
  }
  int64_t L = (T + ptarget - 1) / ptarget;
  if (L < 64) L = 64;
  L = (L + 15) / 16 * 16;
  if (tuning(3) > 0) L = tuning(3);
  p.L = L;
  p.P = ((T + L - 1) / L + 31) / 32 * 32;  // whole warps of segment slots (empty = identity)
  return p;
}

template <class Core, int C, int K, int NSI>
int launch_fixed(const typename Core::Params& prm, int64_t nchains, cudaStream_t s) {
  static_assert(SweepCfg<Core, C, K, NSI, 2>::FITS, "sweep configuration does not fit");
  cudaError_t e = launch_chain_sweep<Core, C, K, NSI, 2>(prm, nchains, s);
  if (e != cudaSuccess) {
    set_last_error(cudaGetErrorString(e));
    return MF_ERR_CUDA;
  }
  return MF_OK;
}

template <typename T, int D, bool SUMMARY>
int launch_core(const KalmanSdeParams<T>& prm, int64_t nchains, cudaStream_t s) {
  using Core = typename std::conditional<SUMMARY, KalmanSdeSummaryCore<T, D>, KalmanSdeFilterCore<T, D>>::type;
  if constexpr (std::is_same<T, double>::value && D == 2) {
    switch (tuning(8)) {  // staged geometries, D = 2 f64 only (A/B against the direct kernel)
      case 1: return launch_fixed<Core, 32, 16, 3>(prm, nchains, s);
      case 2: return launch_fixed<Core, 64, 16, 2>(prm, nchains, s);
      case 3: return launch_fixed<Core, 128, 16, 2>(prm, nchains, s);
      case 4: return launch_fixed<Core, 64, 32, 2>(prm, nchains, s);
      case 5:  // register caps of the direct kernel: 5 / 6 CTAs of 128 threads per SM
        kalman_sde_direct_kernel<T, D, SUMMARY, 128, 5><<<grid_for(nchains, 128), 128, 0, s>>>(prm);
        return check_launch();
      case 6:
        kalman_sde_direct_kernel<T, D, SUMMARY, 128, 6><<<grid_for(nchains, 128), 128, 0, s>>>(prm);
        return check_launch();
      default: break;
    }
  }
  constexpr int NT = 128;
  kalman_sde_direct_kernel<T, D, SUMMARY, NT><<<grid_for(nchains, NT), NT, 0, s>>>(prm);
  return check_launch();
}

template <class F>
int dispatch_sde(int dtype, int64_t D, F&& f) {
  if (dtype != MF_F32 && dtype != MF_F64) return MF_ERR_BAD_ARG;
#define MF_SDE_CASE(n)                                                \
  case n:                                                             \
    if (dtype == MF_F64) return f(TypeTag<double>{}, IntTag<n>{});    \
    return f(TypeTag<float>{}, IntTag<n>{});
  switch (D) {
    MF_SDE_CASE(1) MF_SDE_CASE(2) MF_SDE_CASE(3)
    default: return MF_ERR_UNSUPPORTED;
  }
#undef MF_SDE_CASE
}

}  // namespace

extern "C" {

size_t mf_kalman_matern_workspace_bytes(int dtype, int64_t B, int64_t T, int64_t D) {
  if (B < 1 || T < 1 || D < 1 || D > 3) return 0;
  const size_t es = dtype == MF_F64 ? 8 : 4;
  const size_t N = 3 * D * D + 2 * D + 1;
  const SdePlan pl = make_sde_plan(B, T);
  const size_t slots = (size_t)(pl.P + 31) / 32;
  return align_up(es * (size_t)B * slots * N) + 256;
}

static int matern_impl(int dtype, const void* lengthscale, const void* variance, double jitter,
                       const void* time_deltas, const void* obs, const void* chol_r, void* out, void* out_elem,
                       int64_t B, int64_t T, int64_t D, int first_is_initial, void* workspace,
                       size_t workspace_bytes, void* stream, const KalmanPeerArgs* peers);

int mf_kalman_matern_log_likelihood(int dtype, const void* lengthscale, const void* variance,
                                    double jitter, const void* time_deltas, const void* obs,
                                    const void* chol_r, void* out, void* out_elem, int64_t B,
                                    int64_t T, int64_t D, int first_is_initial, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  return matern_impl(dtype, lengthscale, variance, jitter, time_deltas, obs, chol_r, out, out_elem, B, T, D,
                     first_is_initial, workspace, workspace_bytes, stream, nullptr);
}

int mf_kalman_matern_time_sharded_log_likelihood(int dtype, const void* lengthscale, const void* variance,
                                                 double jitter, const void* time_deltas, const void* obs,
                                                 const void* chol_r, void* out, void* out_elem, int64_t B,
                                                 int64_t T, int64_t D, int first_is_initial,
                                                 void* const* peer_regions, int rank, int world, uint64_t epoch,
                                                 void* workspace, size_t workspace_bytes, void* stream) {
  if (!out || !out_elem || !peer_regions || world < 1 || world > 8 || rank < 0 || rank >= world || epoch == 0)
    return MF_ERR_BAD_ARG;
  KalmanPeerArgs pa;
  pa.rank = rank; pa.world = world; pa.epoch = epoch;
  for (int r = 0; r < 8; ++r) pa.region[r] = r < world ? peer_regions[r] : nullptr;
  for (int r = 0; r < world; ++r)
    if (!pa.region[r]) return MF_ERR_BAD_ARG;
  return matern_impl(dtype, lengthscale, variance, jitter, time_deltas, obs, chol_r, out, out_elem, B, T, D,
                     first_is_initial, workspace, workspace_bytes, stream, &pa);
}

static int matern_impl(int dtype, const void* lengthscale, const void* variance, double jitter,
                       const void* time_deltas, const void* obs, const void* chol_r, void* out, void* out_elem,
                       int64_t B, int64_t T, int64_t D, int first_is_initial, void* workspace,
                       size_t workspace_bytes, void* stream, const KalmanPeerArgs* peers) {
  if (B < 0 || T < 1 || D < 1) return MF_ERR_BAD_ARG;
  if (D > 3) return MF_ERR_UNSUPPORTED;
  if (B == 0) return MF_OK;
  if (!lengthscale || !variance || !obs || !chol_r) return MF_ERR_BAD_ARG;
  if ((T - (first_is_initial ? 1 : 0)) > 0 && !time_deltas) return MF_ERR_BAD_ARG;
  if (!out && !out_elem) return MF_ERR_BAD_ARG;
  if (!first_is_initial && !out_elem) return MF_ERR_BAD_ARG;  // a later time segment has no ell of its own
  if (peers) first_is_initial = first_is_initial ? 1 : 0;
  cudaStream_t s = (cudaStream_t)stream;
  SdePlan pl = make_sde_plan(B, T);
  const size_t need = mf_kalman_matern_workspace_bytes(dtype, B, T, D);
  if (pl.P > 1 && (!workspace || workspace_bytes < need)) {
    if (out_elem) return MF_ERR_BAD_ARG;
    pl.P = 1;  // no scratch: one chain per series
    pl.L = T;
  }
  return dispatch_sde(dtype, D, [&](auto tt, auto dd) {
    using Tp = typename decltype(tt)::type;
    constexpr int kD = decltype(dd)::value;
    KalmanSdeParams<Tp> p;
    p.ls = (const Tp*)lengthscale; p.var = (const Tp*)variance; p.dt = (const Tp*)time_deltas;
    p.obs = (const Tp*)obs; p.chol_r = (const Tp*)chol_r; p.jitter = (Tp)jitter;
    p.B = B; p.Tn = T; p.P = pl.P; p.L = pl.L; p.first_is_initial = first_is_initial;
    p.reduce_warp = 0;
    if (pl.P == 1) {
      if (!out_elem) {
        p.out = (Tp*)out;
        return launch_core<Tp, kD, false>(p, B, s);
      }
      p.out = (Tp*)out_elem;
      int rc = launch_core<Tp, kD, true>(p, B, s);
      if (rc == MF_OK && peers) return kalman_sweep_reduce(dtype, kD, out_elem, out_elem, out, B, 1, s, peers);
      if (rc != MF_OK || !out) return rc;
      // ell is the last of the N values of an element
      constexpr int N = ScanElem<Tp, kD>::N;
      cudaError_t e = cudaMemcpy2DAsync(out, sizeof(Tp), (const Tp*)out_elem + (N - 1),
                                        sizeof(Tp) * N, sizeof(Tp), (size_t)B,
                                        cudaMemcpyDeviceToDevice, s);
      if (e != cudaSuccess) { set_last_error(cudaGetErrorString(e)); return MF_ERR_CUDA; }
      return MF_OK;
    }
    // parallel in time, ONE pass over (dt, y): per-warp joins of the segment elements, then an
    // ordered reduction whose ell component is the log-likelihood
    p.reduce_warp = 1;
    p.out = (Tp*)workspace;
    int rc = launch_core<Tp, kD, true>(p, B * pl.P, s);
    if (rc != MF_OK) return rc;
    return kalman_sweep_reduce(dtype, kD, workspace, out_elem, out, B, pl.P / 32, s, peers);
  });
}

}  // extern "C"
