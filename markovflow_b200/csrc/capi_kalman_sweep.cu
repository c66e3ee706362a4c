// TMA-sweep implementation of the m = 1 Kalman log-likelihood (kalman_sweep.cuh); called from
// capi_kalman.cu through kalman_sweep_api.h.
#include "dispatch.cuh"
#include "kalman_sweep.cuh"
#include "kalman_sweep_api.h"

#include <type_traits>

namespace mf {

namespace {

template <class Core, int C, int K, int NSI>
constexpr bool sweep_fits() { return SweepCfg<Core, C, K, NSI, 2>::FITS; }

// Bulk copies cost ~15 cycles of TMA issue each on top of their bytes (measured), so long tiles
// (K = 16 steps per copy) with two compute warps and a double-buffered ring come first.
template <class Core>
struct SweepPick {
  // measured on config 3 (ping-pong prefetch cores): (64, 16, 2) 0.213 ms < (64, 12, 2) 0.236 <
  // (96, 10, 2) 0.253 < (96, 8, 2) 0.267 < (128, 6, 2) 0.298: long tiles win (TMA issue per copy)
  static constexpr bool Z = sweep_fits<Core, 96, 8, 2>();
  static constexpr bool A = sweep_fits<Core, 64, 16, 2>();
  static constexpr bool B = sweep_fits<Core, 32, 16, 3>();
  static constexpr bool Cc = sweep_fits<Core, 32, 16, 2>();
  static constexpr bool Dd = sweep_fits<Core, 32, 8, 3>();
  static constexpr int C = A ? 64 : (Z ? 96 : 32);
  static constexpr int K = A ? 16 : (Z ? 8 : ((B || Cc) ? 16 : (Dd ? 8 : 4)));
  static constexpr int NSI = A ? 2 : (Z ? 2 : (B ? 3 : (Cc ? 2 : 3)));
  static_assert(sweep_fits<Core, C, K, NSI>(), "no sweep configuration fits");
};

template <int D> struct ScanNT { static constexpr int value = D <= 2 ? 256 : (D == 3 ? 128 : 64); };

template <class F>
int dispatch_sweep(int dtype, int64_t D, F&& f) {
  if (dtype != MF_F32 && dtype != MF_F64) return MF_ERR_BAD_ARG;
#define MF_KS_CASE(n)                                                 \
  case n:                                                             \
    if (dtype == MF_F64) return f(TypeTag<double>{}, IntTag<n>{});    \
    return f(TypeTag<float>{}, IntTag<n>{});
  switch (D) {
    MF_KS_CASE(1) MF_KS_CASE(2) MF_KS_CASE(3) MF_KS_CASE(4)
    default: return MF_ERR_UNSUPPORTED;
  }
#undef MF_KS_CASE
}

template <class Core, int C, int K, int NSI>
int launch_fixed(const typename Core::Params& prm, int64_t nchains, cudaStream_t s) {
  if constexpr (SweepCfg<Core, C, K, NSI, 2>::FITS) {
    cudaError_t e = launch_chain_sweep<Core, C, K, NSI, 2>(prm, nchains, s);
    if (e != cudaSuccess) {
      set_last_error(cudaGetErrorString(e));
      return MF_ERR_CUDA;
    }
    return MF_OK;
  } else {
    return MF_ERR_UNSUPPORTED;
  }
}

}  // namespace
namespace {

template <class Core>
int launch_core(const typename Core::Params& prm, int64_t nchains, cudaStream_t s) {
  using P = SweepPick<Core>;
  // experiment hook (tuning knob 5) for the config-3 summary core only
  if constexpr (std::is_same<Core, KalmanSummaryCore<double, 2, false>>::value) {
    switch (tuning(5)) {
      case 1: return launch_fixed<Core, 32, 32, 2>(prm, nchains, s);
      case 2: return launch_fixed<Core, 128, 8, 2>(prm, nchains, s);
      case 3: return launch_fixed<Core, 32, 16, 3>(prm, nchains, s);
      case 4: return launch_fixed<Core, 64, 8, 3>(prm, nchains, s);
      case 5: return launch_fixed<Core, 96, 8, 2>(prm, nchains, s);
      case 6: return launch_fixed<Core, 32, 8, 3>(prm, nchains, s);
      case 7: return launch_fixed<Core, 128, 4, 3>(prm, nchains, s);
      case 8: return launch_fixed<Core, 192, 4, 2>(prm, nchains, s);
      case 9: return launch_fixed<Core, 64, 4, 3>(prm, nchains, s);
      case 10: return launch_fixed<Core, 32, 4, 3>(prm, nchains, s);
      case 11: return launch_fixed<Core, 64, 8, 2>(prm, nchains, s);
      case 12: return launch_fixed<Core, 160, 4, 2>(prm, nchains, s);
      case 13: return launch_fixed<Core, 96, 10, 2>(prm, nchains, s);
      case 14: return launch_fixed<Core, 128, 6, 2>(prm, nchains, s);
      case 15: return launch_fixed<Core, 64, 16, 2>(prm, nchains, s);
      case 16: return launch_fixed<Core, 64, 12, 2>(prm, nchains, s);
      default: break;
    }
  }
  // few chains: one compute warp per CTA so that more SMs get a CTA
  if constexpr (P::C > 32 && (sweep_fits<Core, 32, 16, 3>() || sweep_fits<Core, 32, 16, 2>())) {
    if (nchains <= (int64_t)148 * 48)
      return launch_fixed<Core, 32, 16, sweep_fits<Core, 32, 16, 3>() ? 3 : 2>(prm, nchains, s);
  }
  return launch_fixed<Core, P::C, P::K, P::NSI>(prm, nchains, s);
}

}  // namespace

int kalman_sweep_chains_per_cta(int64_t D) {
  // must match SweepPick for the f64 cores (static_asserts below)
  return D <= 2 ? 64 : 32;
}

int kalman_sweep_scan_threads(int64_t D) { return D <= 2 ? 256 : (D == 3 ? 128 : 64); }

int kalman_sweep_launch(int mode, const KalmanRawArgs& r, int64_t P, int64_t L,
                        const void* local_prefix, const void* block_prefix, int64_t nblk,
                        int have_prefix, void* out, cudaStream_t s, int reduce_warp) {
  return dispatch_sweep(r.dtype, r.D, [&](auto tt, auto dd) {
    using Tp = typename decltype(tt)::type;
    constexpr int kD = decltype(dd)::value;
    KalmanSweepParams<Tp> p;
    p.g.mu0 = (const Tp*)r.mu0; p.g.chol_p0 = (const Tp*)r.chol_p0; p.g.a = (const Tp*)r.a;
    p.g.b = (const Tp*)r.b; p.g.chol_q = (const Tp*)r.chol_q; p.g.h = (const Tp*)r.h;
    p.g.obs = (const Tp*)r.obs; p.g.chol_r = (const Tp*)r.chol_r;
    p.g.B = r.B; p.g.Tn = r.T; p.g.Bh = r.h_batch; p.g.Tr = r.r_steps; p.g.m = 1;
    p.g.first_is_initial = r.first_is_initial;
    p.P = P; p.L = L; p.reduce_warp = reduce_warp;
    p.local_prefix = (const Tp*)local_prefix; p.block_prefix = (const Tp*)block_prefix;
    p.nblk = nblk; p.scan_nt = ScanNT<kD>::value; p.have_prefix = have_prefix;
    p.out = (Tp*)out;
    const int64_t nchains = r.B * P;
    const bool tvr = r.r_steps != 1;
    if (mode == 1) {
      if (tvr) return launch_core<KalmanSummaryCore<Tp, kD, true>>(p, nchains, s);
      return launch_core<KalmanSummaryCore<Tp, kD, false>>(p, nchains, s);
    }
    if (mode == 2) {
      if (tvr) return launch_core<KalmanFilterCore<Tp, kD, true, true>>(p, nchains, s);
      return launch_core<KalmanFilterCore<Tp, kD, false, true>>(p, nchains, s);
    }
    if (tvr) return launch_core<KalmanFilterCore<Tp, kD, true, false>>(p, nchains, s);
    return launch_core<KalmanFilterCore<Tp, kD, false, false>>(p, nchains, s);
  });
}

int kalman_sweep_block_scan(int dtype, int64_t D, void* elems, void* block_agg, int64_t B,
                            int64_t P, int64_t nblk, cudaStream_t s) {
  return dispatch_sweep(dtype, D, [&](auto tt, auto dd) {
    using Tp = typename decltype(tt)::type;
    constexpr int kD = decltype(dd)::value;
    dim3 grid((unsigned)nblk, (unsigned)B);
    kalman_block_scan_kernel<Tp, kD, ScanNT<kD>::value><<<grid, ScanNT<kD>::value, 0, s>>>(
        (Tp*)elems, (Tp*)block_agg, P, nblk);
    return check_launch();
  });
}

int kalman_sweep_top_scan(int dtype, int64_t D, const void* block_agg, const void* prefix_in,
                          void* block_prefix, void* total_out, void* ell_out, int64_t B,
                          int64_t nblk, cudaStream_t s) {
  return dispatch_sweep(dtype, D, [&](auto tt, auto dd) {
    using Tp = typename decltype(tt)::type;
    constexpr int kD = decltype(dd)::value;
    kalman_top_scan_kernel<Tp, kD, ScanNT<kD>::value><<<(unsigned)B, ScanNT<kD>::value, 0, s>>>(
        (const Tp*)block_agg, (const Tp*)prefix_in, (Tp*)block_prefix, (Tp*)total_out,
        (Tp*)ell_out, nblk);
    return check_launch();
  });
}

int kalman_sweep_reduce(int dtype, int64_t D, const void* elems, void* total_out, void* ell_out,
                        int64_t B, int64_t P, cudaStream_t s, const KalmanPeerArgs* peers) {
  return dispatch_sweep(dtype, D, [&](auto tt, auto dd) {
    using Tp = typename decltype(tt)::type;
    constexpr int kD = decltype(dd)::value;
    constexpr int NT = kD <= 2 ? 512 : 256;
    PeerExchange px;
    px.world = 1; px.rank = 0; px.epoch = 0; px.B = B;
    for (int r = 0; r < 8; ++r) px.region[r] = nullptr;
    if (peers && peers->world > 1) {
      px.world = peers->world; px.rank = peers->rank; px.epoch = peers->epoch;
      for (int r = 0; r < peers->world; ++r) px.region[r] = peers->region[r];
    }
    kalman_reduce_kernel<Tp, kD, NT><<<(unsigned)B, NT, 0, s>>>((const Tp*)elems, (Tp*)total_out,
                                                               (Tp*)ell_out, P, px);
    return check_launch();
  });
}

static_assert(SweepPick<KalmanSummaryCore<double, 1, false>>::C == 64, "chains per CTA table");
static_assert(SweepPick<KalmanSummaryCore<double, 2, false>>::C == 64, "chains per CTA table");
static_assert(SweepPick<KalmanSummaryCore<double, 3, false>>::C == 32, "chains per CTA table");
static_assert(SweepPick<KalmanSummaryCore<double, 4, false>>::C == 32, "chains per CTA table");
static_assert(SweepPick<KalmanFilterCore<double, 2, false, true>>::C == 64, "chains per CTA table");

}  // namespace mf
