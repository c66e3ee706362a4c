// TMA-sweep implementations of LowerTriangularBlockTriDiagonal.solve, the sparse inverse subset and
// the U D U^T factorisation (btd_sweep_cores.cuh); called from capi_btd.cu.
#include <type_traits>

#include "btd_pit.cuh"
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

// segments per chain for a parallel-in-time sweep whose seeds are folded by one thread per chain
static void plan_pit(int64_t chains, int64_t T, int min_len, int64_t* P, int64_t* L,
                     bool sequential_fold = true) {
  // the three passes move ~1.5-2x the bytes of the sequential sweep: pays off below ~2000 chains
  const int64_t target = (int64_t)148 * 192;
  int64_t np = chains > 2048 ? 1 : (target + chains - 1) / chains;
  if (np > T / 64) np = T / 64;
  if (sequential_fold && np > 512) np = 512;
  if (tuning(3) > 1 && tuning(3) < T) np = (T + tuning(3) - 1) / tuning(3);
  if (np < 1) np = 1;
  int64_t l = (T + np - 1) / np;
  if (l < min_len) l = min_len;
  // a ragged last segment must still hold the parked vectors (backward sweeps summarise it)
  while (l < T && T % l != 0 && T % l < min_len) ++l;
  *L = l;
  *P = (T + l - 1) / l;
}

int btd_sweep_solve(int dtype, int64_t D, const void* ld, const void* ls, const void* rhs, void* out,
                    int64_t n, int64_t Bm, int64_t T, int transpose, cudaStream_t s) {
  return dispatch_btd_sweep(dtype, D, [&](auto tt, auto dd) {
    using Tp = typename decltype(tt)::type;
    constexpr int kD = decltype(dd)::value;
    BtdSolveParams<Tp> p{(const Tp*)ld, (const Tp*)ls, (const Tp*)rhs, (Tp*)out, n, Bm, T, 1, T};
    // few long chains with a recursion (sub-diagonal): parallel in time; the output slots serve as
    // scratch, so not when out aliases rhs
    // (block-diagonal matrices have no recursion at all: their segments are simply independent)
    if (tuning(2) != 1 && T >= 128 && (!ls || out != rhs)) plan_pit(n, T, kD + 2, &p.P, &p.L, false);
    const int v = (transpose ? 4 : 0) | (ld ? 0 : 2) | (ls ? 1 : 0);
    auto go = [&](auto tr, auto unit, auto sub) -> int {
      constexpr bool kT = decltype(tr)::value, kU = decltype(unit)::value, kS = decltype(sub)::value;
      if (p.P > 1 && kS) {
        int rc = run<BtdSolveCore<Tp, kD, kT, kU, kS, true>>(p, n * p.P, s);
        if (rc != MF_OK) return rc;
        if (warp_fold(p.P)) btd_solve_seed_kernel<Tp, kD, kT, true><<<grid_for(n * 32, 128), 128, 0, s>>>(p);
        else btd_solve_seed_kernel<Tp, kD, kT, false><<<grid_for(n, 128), 128, 0, s>>>(p);
        rc = check_launch();
        if (rc != MF_OK) return rc;
      }
      return run<BtdSolveCore<Tp, kD, kT, kU, kS, false>>(p, n * p.P, s);
    };
    using Y = std::true_type;
    using N = std::false_type;
    switch (v) {
      case 0: return go(N{}, N{}, N{});
      case 1: return go(N{}, N{}, Y{});
      case 3: return go(N{}, Y{}, Y{});
      case 4: return go(Y{}, N{}, N{});
      case 5: return go(Y{}, N{}, Y{});
      case 7: return go(Y{}, Y{}, Y{});
      default: return (int)MF_ERR_UNSUPPORTED;  // identity matrix: nothing to sweep
    }
  });
}

int btd_sweep_inverse_subset(int dtype, int64_t D, const void* ld, const void* ls, void* od, void* os,
                             int64_t B, int64_t T, cudaStream_t s) {
  if (!ls) return MF_ERR_UNSUPPORTED;
  return dispatch_btd_sweep(dtype, D, [&](auto tt, auto dd) {
    using Tp = typename decltype(tt)::type;
    constexpr int kD = decltype(dd)::value;
    BtdInvSubsetParams<Tp> p{(const Tp*)ld, (const Tp*)ls, (Tp*)od, (Tp*)os, B, T, 1, T};
    if (tuning(2) != 1 && T >= 128 && od != ld) plan_pit(B, T, 2, &p.P, &p.L, false);
    if (p.P > 1) {
      int rc = run<BtdInvSubsetCore<Tp, kD, false, true>>(p, B * p.P, s);
      if (rc != MF_OK) return rc;
      if (warp_fold(p.P)) btd_inv_subset_seed_kernel<Tp, kD, true><<<grid_for(B * 32, 128), 128, 0, s>>>(p);
      else btd_inv_subset_seed_kernel<Tp, kD, false><<<grid_for(B, 128), 128, 0, s>>>(p);
      rc = check_launch();
      if (rc != MF_OK) return rc;
    }
    if (os) return run<BtdInvSubsetCore<Tp, kD, true, false>>(p, B * p.P, s);
    return run<BtdInvSubsetCore<Tp, kD, false, false>>(p, B * p.P, s);
  });
}

int btd_sweep_udu(int dtype, int64_t D, const void* diag, const void* sub, void* ou, void* ocd,
                  int32_t* info, int64_t B, int64_t T, cudaStream_t s) {
  return dispatch_btd_sweep(dtype, D, [&](auto tt, auto dd) {
    using Tp = typename decltype(tt)::type;
    constexpr int kD = decltype(dd)::value;
    BtdUduParams<Tp> p{(const Tp*)diag, (const Tp*)sub, (Tp*)ou, (Tp*)ocd, info, B, T, 1, T};
    if (tuning(2) != 1 && T >= 128 && ocd != diag && ou != sub) plan_pit(B, T, 2, &p.P, &p.L, false);
    if (p.P > 1) {
      if (info && cudaMemsetAsync(info, 0, sizeof(int32_t) * B, s) != cudaSuccess) return check_launch();
      int rc = run<BtdUduCore<Tp, kD, true>>(p, B * p.P, s);
      if (rc != MF_OK) return rc;
      if (warp_fold(p.P)) btd_udu_seed_kernel<Tp, kD, true><<<grid_for(B * 32, 128), 128, 0, s>>>(p);
      else btd_udu_seed_kernel<Tp, kD, false><<<grid_for(B, 128), 128, 0, s>>>(p);
      rc = check_launch();
      if (rc != MF_OK) return rc;
    }
    return run<BtdUduCore<Tp, kD, false>>(p, B * p.P, s);
  });
}

// Few long chains: block Cholesky (+ solve) parallel in time (btd_pit.cuh).  MF_ERR_UNSUPPORTED when
// it does not apply (many chains, short chains, aliased outputs, D > 4): the caller then runs the
// sequential sweep.  out_logdet is filled from the factor afterwards by the caller.
int btd_sweep_cholesky_pit(int dtype, int64_t D, const void* diag, const void* sub, const void* rhs,
                           void* od, void* os, void* ox, int32_t* info, int64_t B, int64_t T,
                           cudaStream_t s) {
  if (!sub || !os || T < 128 || B > 1024 || tuning(2) == 1) return MF_ERR_UNSUPPORTED;
  if (od == diag || os == sub || (rhs && ox == rhs)) return MF_ERR_UNSUPPORTED;  // slots are scratch
  return dispatch_btd_sweep(dtype, D, [&](auto tt, auto dd) {
    using Tp = typename decltype(tt)::type;
    constexpr int kD = decltype(dd)::value;
    CholPitParams<Tp> p{(const Tp*)diag, (const Tp*)sub, (const Tp*)rhs, (Tp*)od, (Tp*)os, (Tp*)ox,
                        info, B, T, 1, T};
    const int64_t target = (int64_t)148 * 192;
    int64_t np = (target + B - 1) / B;
    if (np > T / 64) np = T / 64;
    if (tuning(3) > 1 && tuning(3) < T) np = (T + tuning(3) - 1) / tuning(3);
    if (np < 2) return (int)MF_ERR_UNSUPPORTED;
    p.L = (T + np - 1) / np;
    if (p.L < 2) p.L = 2;
    p.P = (T + p.L - 1) / p.L;
    if (p.P < 2) return (int)MF_ERR_UNSUPPORTED;
    if (info && cudaMemsetAsync(info, 0, sizeof(int32_t) * B, s) != cudaSuccess) return check_launch();
    int rc;
    if (rhs) rc = run<CholPitSummaryCore<Tp, kD, true>>(p, B * p.P, s);
    else rc = run<CholPitSummaryCore<Tp, kD, false>>(p, B * p.P, s);
    if (rc != MF_OK) return rc;
    // many segments per chain: the fold is a warp scan over the elements
    if (warp_fold(p.P)) {
      if (rhs) chol_pit_seed_kernel<Tp, kD, true, true><<<grid_for(B * 32, 128), 128, 0, s>>>(p);
      else chol_pit_seed_kernel<Tp, kD, false, true><<<grid_for(B * 32, 128), 128, 0, s>>>(p);
    } else {
      if (rhs) chol_pit_seed_kernel<Tp, kD, true, false><<<grid_for(B, 128), 128, 0, s>>>(p);
      else chol_pit_seed_kernel<Tp, kD, false, false><<<grid_for(B, 128), 128, 0, s>>>(p);
    }
    rc = check_launch();
    if (rc != MF_OK) return rc;
    if (rhs) return run<CholPitCore<Tp, kD, true>>(p, B * p.P, s);
    return run<CholPitCore<Tp, kD, false>>(p, B * p.P, s);
  });
}

}  // namespace mf
