// TMA-sweep implementations of the sequential StateSpaceModel / natural-parameter recurrences
// (ssm_sweep.cuh); called from capi_ssm.cu and capi_nat.cu through ssm_sweep_api.h.
#include <type_traits>

#include "dispatch.cuh"
#include "ssm_sweep.cuh"
#include "ssm_sweep_api.h"

namespace mf {

namespace {

template <class F>
int dispatch_ssm_sweep(int dtype, int64_t D, F&& f) {
  if (dtype != MF_F32 && dtype != MF_F64) return MF_ERR_BAD_ARG;
#define MF_SS_CASE(n)                                                 \
  case n:                                                             \
    if (dtype == MF_F64) return f(TypeTag<double>{}, IntTag<n>{});    \
    return f(TypeTag<float>{}, IntTag<n>{});
  switch (D) {
    MF_SS_CASE(1) MF_SS_CASE(2) MF_SS_CASE(3) MF_SS_CASE(4)
    default: return MF_ERR_UNSUPPORTED;
  }
#undef MF_SS_CASE
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

// Segments per chain for the parallel-in-time evaluation of an exact (affine) recurrence: enough
// virtual chains to occupy the GPU, segments of at least 64 steps.
static void plan_segments(int64_t B, int64_t T, int64_t* P, int64_t* L, int64_t target = (int64_t)148 * 192,
                          bool whole_waves = false) {
  // The three passes move ~1.5-2x the bytes of the sequential sweep, whose cost is T x (latency of a
  // step) whatever B: parallel in time pays off below ~2000 chains (measured: B = 4096 is slower).
  // whole_waves: `target` is the number of rows the resident CTAs hold at once -- never plan more (a few CTAs
  // spilling into one more wave cost a whole wave)
  int64_t p = B > 2048 ? 1 : (whole_waves ? target / B : (target + B - 1) / B);
  if (p > T / 64) p = T / 64;
  if (p < 1) p = 1;
  if (tuning(3) > 0 && tuning(3) < T) p = (T + tuning(3) - 1) / tuning(3);
  int64_t l = (T + p - 1) / p;
  // segments of a multiple of 4 steps start 16-byte aligned in every stream of either dtype (what the tensor-map
  // engine needs, sweep_tm.cuh); an explicit segment length (knob 3) is taken as given
  if (!(tuning(3) > 0 && tuning(3) < T) && l >= 16) l = (l + 3) / 4 * 4;
  if (l < 2) l = 2;
  *L = l;
  *P = (T + l - 1) / l;
}

int ssm_sweep_moments(int dtype, int64_t D, int expectations, const void* mu0, const void* chol_p0,
                      const void* a, const void* b, const void* chol_q, void* o_vec, void* o_diag,
                      void* o_sub, int64_t B, int64_t T, cudaStream_t s) {
  return dispatch_ssm_sweep(dtype, D, [&](auto tt, auto dd) {
    using Tp = typename decltype(tt)::type;
    constexpr int kD = decltype(dd)::value;
    SsmMomentsParams<Tp> p{(const Tp*)mu0, (const Tp*)chol_p0, (const Tp*)a, (const Tp*)b,
                           (const Tp*)chol_q, (Tp*)o_vec, (Tp*)o_diag, (Tp*)o_sub, B, T, 1, T,
                           (Tp*)o_vec, (Tp*)o_diag, 0};
    // float32 forward sweeps stay on the 1-D engine (`b`: [B, T-1, 2] floats, chains 8 bytes off a 16-byte stride,
    // cannot be tensor-mapped).  At D <= 2 its CTAs hold 64 rows and two are resident per SM: 148 x 128 rows are
    // ONE wave, and the segment count is rounded DOWN to stay inside it.  Measured on config 5 (1024 chains):
    // 18 segments (288 CTAs, one wave) 0.309 ms, 36 (two waves) 0.320 ms, 55 (three) 0.325 ms -- against
    // 0.372 ms for the 56 segments (896 CTAs: three waves and 8 CTAs) the rounded-up 148 x 384 rows gave.
    const bool one_wave = sizeof(Tp) == 4 && kD <= 2;
    const int64_t target = (int64_t)148 * (one_wave ? 128 : (sizeof(Tp) == 4 ? 384 : 192));
    if (tuning(2) != 1 && o_diag && (o_vec || !expectations)) plan_segments(B, T, &p.P, &p.L, target, one_wave);
    if (p.P > 1) {
      int rc = run<SsmMomSummaryCore<Tp, kD>>(p, B * p.P, s);
      if (rc != MF_OK) return rc;
      if (warp_fold(p.P)) ssm_moments_seed_kernel<Tp, kD, true><<<grid_for(B * 32, 128), 128, 0, s>>>(p);
      else ssm_moments_seed_kernel<Tp, kD, false><<<grid_for(B, 128), 128, 0, s>>>(p);
      rc = check_launch();
      if (rc != MF_OK) return rc;
    }
    if (expectations) return run<SsmMomentsCore<Tp, kD, true>>(p, B * p.P, s);
    return run<SsmMomentsCore<Tp, kD, false>>(p, B * p.P, s);
  });
}

int ssm_sweep_affine(int dtype, int64_t D, const void* mu0, const void* chol_p0, const void* a,
                     const void* b, const void* chol_q, const void* eps, void* out, int64_t n,
                     int64_t Bm, int64_t T, cudaStream_t s, int use_rng, unsigned long long seed) {
  return dispatch_ssm_sweep(dtype, D, [&](auto tt, auto dd) {
    using Tp = typename decltype(tt)::type;
    constexpr int kD = decltype(dd)::value;
    SsmAffineParams<Tp> p{(const Tp*)mu0, (const Tp*)chol_p0, (const Tp*)a, (const Tp*)b,
                          (const Tp*)chol_q, (const Tp*)eps, (Tp*)out, n, Bm, T, 1, T, seed};
    if (tuning(2) != 1 && T >= 128 && out != eps) {
      plan_segments(n, T, &p.P, &p.L);
      if (p.L < kD + 2) { p.P = 1; p.L = T; }
    }
    auto go = [&](auto noise, auto rng) -> int {
      constexpr bool kN = decltype(noise)::value;
      constexpr bool kR = decltype(rng)::value;
      if (p.P > 1) {
        int rc = run<SsmAffineCore<Tp, kD, kN, true, kR>>(p, n * p.P, s);
        if (rc != MF_OK) return rc;
        if (warp_fold(p.P)) ssm_affine_seed_kernel<Tp, kD, true><<<grid_for(n * 32, 128), 128, 0, s>>>(p);
        else ssm_affine_seed_kernel<Tp, kD, false><<<grid_for(n, 128), 128, 0, s>>>(p);
        rc = check_launch();
        if (rc != MF_OK) return rc;
      }
      return run<SsmAffineCore<Tp, kD, kN, false, kR>>(p, n * p.P, s);
    };
    if (use_rng) return go(std::true_type{}, std::true_type{});
    if (eps) return go(std::true_type{}, std::false_type{});
    return go(std::false_type{}, std::false_type{});
  });
}

int ssm_sweep_kl(int dtype, int64_t D, const void* q_mu0, const void* q_chol_p0, const void* q_a,
                 const void* q_b, const void* q_chol_q, const void* p_mu0, const void* p_chol_p0,
                 const void* p_a, const void* p_b, const void* p_chol_q, void* out, int64_t B,
                 int64_t T, cudaStream_t s) {
  return dispatch_ssm_sweep(dtype, D, [&](auto tt, auto dd) {
    using Tp = typename decltype(tt)::type;
    constexpr int kD = decltype(dd)::value;
    SsmKlParams<Tp> p{(const Tp*)q_mu0, (const Tp*)q_chol_p0, (const Tp*)q_a, (const Tp*)q_b,
                      (const Tp*)q_chol_q, (const Tp*)p_mu0, (const Tp*)p_chol_p0, (const Tp*)p_a,
                      (const Tp*)p_b, (const Tp*)p_chol_q, (Tp*)out, B, T, 1, T, nullptr, nullptr, nullptr};
    if (tuning(2) != 1 && T >= 128) plan_segments(B, T, &p.P, &p.L);
    if (p.P == 1) return run<SsmKlCore<Tp, kD>>(p, B, s);
    // few long chains: parallel in time.  Stream-ordered scratch for the elements / seeds / partials.
    const size_t n_el = (size_t)B * p.P * 2;
    const size_t bytes = sizeof(Tp) * (n_el * (kD + kD * kD) + (size_t)B * p.P);
    Tp* ws = nullptr;
    if (cudaMallocAsync((void**)&ws, bytes, s) != cudaSuccess) return check_launch();
    Tp* ws_vec = ws;
    Tp* ws_diag = ws + n_el * kD;
    p.seed_vec = ws_vec;
    p.seed_diag = ws_diag;
    p.partial = ws_diag + n_el * kD * kD;
    SsmMomentsParams<Tp> m{(const Tp*)q_mu0, (const Tp*)q_chol_p0, (const Tp*)q_a, (const Tp*)q_b,
                           (const Tp*)q_chol_q, nullptr, nullptr, nullptr, B, T, p.P, p.L, ws_vec, ws_diag, 1};
    int rc = run<SsmMomSummaryCore<Tp, kD>>(m, B * p.P, s);
    if (rc == MF_OK) {
      if (warp_fold(p.P)) ssm_moments_seed_kernel<Tp, kD, true><<<grid_for(B * 32, 128), 128, 0, s>>>(m);
      else ssm_moments_seed_kernel<Tp, kD, false><<<grid_for(B, 128), 128, 0, s>>>(m);
      rc = check_launch();
    }
    if (rc == MF_OK) rc = run<SsmKlCore<Tp, kD>>(p, B * p.P, s);
    if (rc == MF_OK) {
      ssm_kl_reduce_kernel<Tp><<<grid_for(B, 128), 128, 0, s>>>(p.partial, (Tp*)out, B, p.P);
      rc = check_launch();
    }
    cudaFreeAsync(ws, s);
    return rc;
  });
}

int nat_sweep_to_ssm(int dtype, int64_t D, const void* th_lin, const void* th_diag,
                     const void* th_sub, void* out_a, void* out_off, void* out_chol, int32_t* info,
                     int64_t B, int64_t T, cudaStream_t s) {
  return dispatch_ssm_sweep(dtype, D, [&](auto tt, auto dd) {
    using Tp = typename decltype(tt)::type;
    constexpr int kD = decltype(dd)::value;
    NatToSsmParams<Tp> p{(const Tp*)th_lin, (const Tp*)th_diag, (const Tp*)th_sub, (Tp*)out_a,
                         (Tp*)out_off, (Tp*)out_chol, info, B, T, 1, T};
    if (tuning(2) != 1 && out_a) {
      plan_segments(B, T, &p.P, &p.L);
    }
    if (p.P > 1) {
      if (info && cudaMemsetAsync(info, 0, sizeof(int32_t) * B, s) != cudaSuccess) return check_launch();
      int rc = run<NatSummaryCore<Tp, kD>>(p, B * p.P, s);
      if (rc != MF_OK) return rc;
      // many segments per chain: the fold is a warp scan over the elements
      if (warp_fold(p.P)) nat_seed_kernel<Tp, kD, true><<<grid_for(B * 32, 128), 128, 0, s>>>(p);
      else nat_seed_kernel<Tp, kD, false><<<grid_for(B, 128), 128, 0, s>>>(p);
      rc = check_launch();
      if (rc != MF_OK) return rc;
    }
    return run<NatToSsmCore<Tp, kD>>(p, B * p.P, s);
  });
}

}  // namespace mf
