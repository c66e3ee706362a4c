// TMA-sweep implementations of LowerTriangularBlockTriDiagonal.solve, the sparse inverse subset and
// the U D U^T factorisation (btd_sweep_cores.cuh); called from capi_btd.cu.
#include "btd_sweep_cores.cuh"
#include "dispatch.cuh"
#include "ssm_sweep_api.h"

namespace mf {

namespace {

template <class F>
int dispatch_btd_sweep(int dtype, int64_t D, F&& f) {
  if (dtype != MF_F32 && dtype != MF_F64) return MF_ERR_BAD_ARG;
#define MF_BS_CASE(n)                                                 \
  case n:                                                             \
    if (dtype == MF_F64) return f(TypeTag<double>{}, IntTag<n>{});    \
    return f(TypeTag<float>{}, IntTag<n>{});
  switch (D) {
    MF_BS_CASE(1) MF_BS_CASE(2) MF_BS_CASE(3) MF_BS_CASE(4)
    default: return MF_ERR_UNSUPPORTED;
  }
#undef MF_BS_CASE
}

template <class Core>
int run(const typename Core::Params& p, int64_t nchains, cudaStream_t s) {
  if constexpr (SweepAuto<Core>::ok) {
    cudaError_t e = SweepAuto<Core>::launch(p, nchains, s);
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

int btd_sweep_solve(int dtype, int64_t D, const void* ld, const void* ls, const void* rhs, void* out,
                    int64_t n, int64_t Bm, int64_t T, int transpose, cudaStream_t s) {
  return dispatch_btd_sweep(dtype, D, [&](auto tt, auto dd) {
    using Tp = typename decltype(tt)::type;
    constexpr int kD = decltype(dd)::value;
    BtdSolveParams<Tp> p{(const Tp*)ld, (const Tp*)ls, (const Tp*)rhs, (Tp*)out, n, Bm, T};
    const int v = (transpose ? 4 : 0) | (ld ? 0 : 2) | (ls ? 1 : 0);
    switch (v) {
      case 0: return run<BtdSolveCore<Tp, kD, false, false, false>>(p, n, s);
      case 1: return run<BtdSolveCore<Tp, kD, false, false, true>>(p, n, s);
      case 3: return run<BtdSolveCore<Tp, kD, false, true, true>>(p, n, s);
      case 4: return run<BtdSolveCore<Tp, kD, true, false, false>>(p, n, s);
      case 5: return run<BtdSolveCore<Tp, kD, true, false, true>>(p, n, s);
      case 7: return run<BtdSolveCore<Tp, kD, true, true, true>>(p, n, s);
      default: return MF_ERR_UNSUPPORTED;  // identity matrix: nothing to sweep
    }
  });
}

int btd_sweep_inverse_subset(int dtype, int64_t D, const void* ld, const void* ls, void* od, void* os,
                             int64_t B, int64_t T, cudaStream_t s) {
  if (!ls) return MF_ERR_UNSUPPORTED;
  return dispatch_btd_sweep(dtype, D, [&](auto tt, auto dd) {
    using Tp = typename decltype(tt)::type;
    constexpr int kD = decltype(dd)::value;
    BtdInvSubsetParams<Tp> p{(const Tp*)ld, (const Tp*)ls, (Tp*)od, (Tp*)os, B, T};
    if (os) return run<BtdInvSubsetCore<Tp, kD, true>>(p, B, s);
    return run<BtdInvSubsetCore<Tp, kD, false>>(p, B, s);
  });
}

int btd_sweep_udu(int dtype, int64_t D, const void* diag, const void* sub, void* ou, void* ocd,
                  int32_t* info, int64_t B, int64_t T, cudaStream_t s) {
  return dispatch_btd_sweep(dtype, D, [&](auto tt, auto dd) {
    using Tp = typename decltype(tt)::type;
    constexpr int kD = decltype(dd)::value;
    BtdUduParams<Tp> p{(const Tp*)diag, (const Tp*)sub, (Tp*)ou, (Tp*)ocd, info, B, T};
    return run<BtdUduCore<Tp, kD>>(p, B, s);
  });
}

}  // namespace mf
