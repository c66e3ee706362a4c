// Experimental launches of the sweep engines (selected with mf_set_tuning; used for A/B timing).
#include "btd_sweep.cuh"
#include "dispatch.cuh"
#include "kalman_sweep.cuh"
#include "sweep2.cuh"

namespace mf {

namespace {
template <class Core, int C, int K, int NSI, int NSO, int NLW, int NSW>
int run2(const typename Core::Params& p, int64_t n, cudaStream_t s) {
  if constexpr (Sweep2Cfg<Core, C, K, NSI, NSO, NLW, NSW>::FITS) {
    cudaError_t e = launch_chain_sweep2<Core, C, K, NSI, NSO, NLW, NSW>(p, n, s);
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

template <class Core, int C, int K, int NSI, int NSO>
static int run1(const typename Core::Params& p, int64_t n, cudaStream_t s) {
  if constexpr (SweepCfg<Core, C, K, NSI, NSO>::FITS) {
    cudaError_t e = launch_chain_sweep<Core, C, K, NSI, NSO>(p, n, s);
    if (e != cudaSuccess) {
      set_last_error(cudaGetErrorString(e));
      return MF_ERR_CUDA;
    }
    return MF_OK;
  } else {
    return MF_ERR_UNSUPPORTED;
  }
}

// Cholesky + solve, D = 3, f64 through sweep2; variant selects the ring geometry.
int exp_chol_d3(int variant, const void* diag, const void* sub, const void* rhs, void* od, void* os,
                void* ox, void* logdet, int32_t* info, int64_t B, int64_t T, cudaStream_t s) {
  using Core = CholSweepCore<double, 3, true>;
  CholSweepParams<double> p{(const double*)diag, (const double*)sub, (const double*)rhs, (double*)od,
                            (double*)os, (double*)ox, (double*)logdet, info, B, T};
  switch (variant) {
    case 0: return run2<Core, 64, 4, 2, 2, 4, 4>(p, B, s);
    case 1: return run2<Core, 32, 8, 3, 2, 4, 4>(p, B, s);
    case 2: return run2<Core, 32, 4, 3, 2, 2, 2>(p, B, s);
    case 3: return run2<Core, 64, 4, 2, 1, 4, 4>(p, B, s);
    case 4: return run2<Core, 64, 4, 2, 2, 6, 6>(p, B, s);
    case 5: return run2<Core, 64, 2, 3, 3, 4, 4>(p, B, s);
    case 6: return run2<Core, 96, 2, 3, 2, 4, 4>(p, B, s);
    // TMA engine (sweep.cuh)
    case 7: return run1<Core, 64, 4, 2, 2>(p, B, s);
    case 8: return run1<Core, 32, 8, 3, 2>(p, B, s);
    case 9: return run1<Core, 64, 4, 2, 1>(p, B, s);
    case 10: return run1<Core, 32, 16, 2, 1>(p, B, s);
    case 11: return run1<Core, 64, 6, 2, 1>(p, B, s);
    case 12: return run1<Core, 32, 12, 2, 2>(p, B, s);
    case 13: return run1<Core, 64, 2, 3, 2>(p, B, s);
    default: return MF_ERR_UNSUPPORTED;
  }
}

// Kalman summary sweep, D = 2, f64, shared noise, through sweep2.
int exp_kalman_summary_d2(int variant, const KalmanSweepParams<double>& p, int64_t nchains,
                          cudaStream_t s) {
  using Core = KalmanSummaryCore<double, 2, false>;
  switch (variant) {
    case 0: return run2<Core, 128, 4, 3, 2, 4, 0>(p, nchains, s);
    case 1: return run2<Core, 160, 4, 3, 2, 4, 0>(p, nchains, s);
    case 2: return run2<Core, 96, 8, 2, 2, 4, 0>(p, nchains, s);
    case 3: return run2<Core, 192, 4, 2, 2, 6, 0>(p, nchains, s);
    case 4: return run2<Core, 128, 8, 2, 2, 4, 0>(p, nchains, s);
    case 5: return run2<Core, 64, 16, 2, 2, 4, 0>(p, nchains, s);
    case 6: return run2<Core, 128, 4, 3, 2, 8, 0>(p, nchains, s);
    default: return MF_ERR_UNSUPPORTED;
  }
}

}  // namespace mf
