// C-ABI entry points for the block-tridiagonal operators (include/markovflow_b200.h).
#include <cstdio>
#include <cstring>

#include "btd_direct.cuh"
#include "btd_staged.cuh"
#include "btd_tma.cuh"
#include "dispatch.cuh"
#include "mid_api.h"
#include "ssm_sweep_api.h"

namespace mf {

static thread_local char g_last_error[256] = "";

void set_last_error(const char* msg) {
  std::strncpy(g_last_error, msg, sizeof(g_last_error) - 1);
  g_last_error[sizeof(g_last_error) - 1] = 0;
}

int check_launch() {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error(cudaGetErrorString(e));
    return MF_ERR_CUDA;
  }
  return MF_OK;
}

}  // namespace mf

using namespace mf;

namespace {

// Tuning knobs (mf_set_tuning): 0 = variant of the Cholesky sweep (0 auto = TMA, 1 direct,
// 2 cp.async-staged),
// 1 = steps per shared-memory stage for the staged sweep (0 auto).
int g_tuning[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
}  // namespace
namespace mf {
int tuning(int knob) { return (knob >= 0 && knob < 16) ? g_tuning[knob] : 0; }
}  // namespace mf
namespace {

constexpr int kSmemBudget = 232448 - 1536;  // 227 KB opt-in maximum minus barrier/alignment slack

template <typename T, int D, bool RHS>
struct StagedPick {
  static constexpr int NSI = 3, NSO = 2;
  static constexpr int ETOT = 2 * D * D + (RHS ? D : 0);
  static constexpr int per_k(int c) { return (int)sizeof(T) * ETOT * (c + 1) * (NSI + NSO); }
  static constexpr int C = (2 * per_k(32) <= kSmemBudget) ? 32 : ((2 * per_k(16) <= kSmemBudget) ? 16 : 8);
  static constexpr int KMAX = kSmemBudget / per_k(C);
  static constexpr int K = KMAX >= 8 ? 8 : (KMAX >= 4 ? 4 : (KMAX >= 2 ? 2 : 1));
};

template <typename T, int D, bool RHS, int C, int K, int NSI, int NSO>
int launch_chol_staged(const void* diag, const void* sub, const void* rhs, void* od, void* os,
                       void* ox, void* logdet, int32_t* info, int64_t B, int64_t Tn,
                       cudaStream_t s) {
  using Cfg = CholStagedCfg<T, D, RHS, C, K, NSI, NSO>;
  auto kern = btd_chol_staged_kernel<T, D, RHS, C, K, NSI, NSO>;
  static SmemOnce once;  // per instantiation, per device
  if (ensure_smem(once, kern, Cfg::SMEM_BYTES) != cudaSuccess) return check_launch();
  kern<<<grid_for(B, C), Cfg::THREADS, Cfg::SMEM_BYTES, s>>>(
      (const T*)diag, (const T*)sub, (const T*)rhs, (T*)od, (T*)os, (T*)ox, (T*)logdet, info, B, Tn, tuning(12));
  return check_launch();
}

template <typename T, int D, bool RHS>
int launch_chol_staged_auto(const void* diag, const void* sub, const void* rhs, void* od, void* os,
                            void* ox, void* logdet, int32_t* info, int64_t B, int64_t Tn,
                            cudaStream_t s) {
  using P = StagedPick<T, D, RHS>;
  if (P::K >= 8 && g_tuning[1] == 4)
    return launch_chol_staged<T, D, RHS, P::C, (P::K >= 8 ? 4 : P::K), P::NSI, P::NSO>(
        diag, sub, rhs, od, os, ox, logdet, info, B, Tn, s);
  return launch_chol_staged<T, D, RHS, P::C, P::K, P::NSI, P::NSO>(diag, sub, rhs, od, os, ox,
                                                                  logdet, info, B, Tn, s);
}

// ---- TMA variant: pick chains/CTA and steps/stage so that the ring fits and stays 16B-aligned ----
template <typename T, int D, bool RHS, int C, int K>
constexpr bool tma_fits() {
  using Cfg = CholTmaCfg<T, D, RHS, C, K, 3, 2>;
  return Cfg::ALIGN_OK && Cfg::SMEM_BYTES <= (size_t)232448;
}

template <typename T, int D, bool RHS>
struct TmaPick {
  static constexpr int C = (tma_fits<T, D, RHS, 32, 4>() || tma_fits<T, D, RHS, 32, 8>())
                               ? 32
                               : ((tma_fits<T, D, RHS, 16, 4>() || tma_fits<T, D, RHS, 16, 8>()) ? 16 : 8);
  static constexpr int K = tma_fits<T, D, RHS, C, 8>() ? 8 : (tma_fits<T, D, RHS, C, 4>() ? 4 : 0);
};

template <typename T, int D, bool RHS, int C, int K>
int launch_chol_tma(const void* diag, const void* sub, const void* rhs, void* od, void* os,
                    void* ox, void* logdet, int32_t* info, int64_t B, int64_t Tn, cudaStream_t s) {
  using Cfg = CholTmaCfg<T, D, RHS, C, K, 3, 2>;
  auto kern = btd_chol_tma_kernel<T, D, RHS, C, K, 3, 2>;
  static SmemOnce once;  // per instantiation, per device
  if (ensure_smem(once, kern, Cfg::SMEM_BYTES) != cudaSuccess) return check_launch();
  kern<<<grid_for(B, C), Cfg::THREADS, Cfg::SMEM_BYTES, s>>>(
      (const T*)diag, (const T*)sub, (const T*)rhs, (T*)od, (T*)os, (T*)ox, (T*)logdet, info, B, Tn, tuning(12));
  return check_launch();
}

template <typename T, int D, bool RHS>
int launch_chol_fast(const void* diag, const void* sub, const void* rhs, void* od, void* os,
                     void* ox, void* logdet, int32_t* info, int64_t B, int64_t Tn, cudaStream_t s) {
  using P = TmaPick<T, D, RHS>;
  if constexpr (P::K > 0) {
    if (g_tuning[0] != 2)
      return launch_chol_tma<T, D, RHS, P::C, P::K>(diag, sub, rhs, od, os, ox, logdet, info, B, Tn, s);
  }
  return launch_chol_staged_auto<T, D, RHS>(diag, sub, rhs, od, os, ox, logdet, info, B, Tn, s);
}

}  // namespace

extern "C" {

int mf_set_tuning(int knob, int value) {
  if (knob < 0 || knob >= 16) return MF_ERR_BAD_ARG;
  g_tuning[knob] = value;
  return MF_OK;
}

int mf_version(void) { return 100; }

const char* mf_last_cuda_error(void) { return g_last_error; }

int mf_btd_cholesky(int dtype, const void* diag, const void* sub, const void* rhs, void* out_diag,
                    void* out_sub, void* out_x, void* out_logdet, int32_t* info, int64_t B,
                    int64_t T, int64_t D, void* stream) {
  if (B < 0 || T < 1 || D < 1) return MF_ERR_BAD_ARG;
  if (B == 0) return MF_OK;
  if (!diag || !out_diag) return MF_ERR_BAD_ARG;
  if ((sub != nullptr) != (out_sub != nullptr)) return MF_ERR_BAD_ARG;
  if ((rhs != nullptr) != (out_x != nullptr)) return MF_ERR_BAD_ARG;
  if (T == 1) { sub = nullptr; out_sub = nullptr; }
  cudaStream_t s = (cudaStream_t)stream;
  if (D > MF_SMALL_D_MAX)
    return big_cholesky(dtype, diag, sub, rhs, out_diag, out_sub, out_x, out_logdet, info, B, T, D, s);
  if (D <= kSsmSweepMaxD && tuning(4) != 1) {
    // few long chains: parallel in time (btd_pit.cuh); the log-determinant is read off the factor
    const int rc = btd_sweep_cholesky_pit(dtype, D, diag, sub, rhs, out_diag, out_sub, out_x, info, B, T, s);
    if (rc == MF_OK && out_logdet) return mf_btd_abs_log_det(dtype, out_diag, out_logdet, B, T, D, stream);
    if (rc != MF_ERR_UNSUPPORTED) return rc;
  }
  return dispatch_small(dtype, D, [&](auto tt, auto dd) {
    using Tp = typename decltype(tt)::type;
    constexpr int kD = decltype(dd)::value;
    const bool staged = (sub != nullptr) && g_tuning[0] != 1;
    if (staged) {
      if (rhs)
        return launch_chol_fast<Tp, kD, true>(diag, sub, rhs, out_diag, out_sub, out_x,
                                              out_logdet, info, B, T, s);
      return launch_chol_fast<Tp, kD, false>(diag, sub, rhs, out_diag, out_sub, out_x,
                                             out_logdet, info, B, T, s);
    }
    btd_chol_direct_kernel<Tp, kD><<<grid_for(B, 32), 32, 0, s>>>(
        (const Tp*)diag, (const Tp*)sub, (const Tp*)rhs, (Tp*)out_diag, (Tp*)out_sub, (Tp*)out_x,
        (Tp*)out_logdet, info, B, T);
    return check_launch();
  });
}

int mf_btd_solve(int dtype, const void* ld, const void* ls, const void* rhs, void* out,
                 int64_t n_rhs, int64_t Bm, int64_t T, int64_t D, int transpose, void* stream) {
  if (n_rhs < 0 || Bm < 1 || T < 1 || D < 1) return MF_ERR_BAD_ARG;
  if (n_rhs == 0) return MF_OK;
  if (!rhs || !out) return MF_ERR_BAD_ARG;  // ld == NULL: identity diagonal blocks
  if (T == 1) ls = nullptr;
  cudaStream_t s = (cudaStream_t)stream;
  if (D > MF_SMALL_D_MAX) return big_solve(dtype, ld, ls, rhs, out, n_rhs, Bm, T, D, transpose, s);
  if (D <= kSsmSweepMaxD && T > 1 && tuning(4) != 1) {
    const int rc = btd_sweep_solve(dtype, D, ld, ls, rhs, out, n_rhs, Bm, T, transpose, s);
    if (rc != MF_ERR_UNSUPPORTED) return rc;
  }
  return dispatch_small(dtype, D, [&](auto tt, auto dd) {
    using Tp = typename decltype(tt)::type;
    constexpr int kD = decltype(dd)::value;
    btd_solve_direct_kernel<Tp, kD><<<grid_for(n_rhs, 32), 32, 0, s>>>(
        (const Tp*)ld, (const Tp*)ls, (const Tp*)rhs, (Tp*)out, n_rhs, Bm, T, transpose);
    return check_launch();
  });
}

int mf_btd_inverse_subset(int dtype, const void* ld, const void* ls, void* out_diag,
                          void* out_sub, int64_t B, int64_t T, int64_t D, void* stream) {
  if (B < 0 || T < 1 || D < 1) return MF_ERR_BAD_ARG;
  if (B == 0) return MF_OK;
  if (!ld || !out_diag) return MF_ERR_BAD_ARG;
  if (T == 1) { ls = nullptr; out_sub = nullptr; }
  if (out_sub && !ls) return MF_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  if (mid_dim(D)) return mid_inverse_subset(dtype, ld, ls, out_diag, out_sub, B, T, D, s);
  if (D <= kSsmSweepMaxD && T > 1 && tuning(4) != 1) {
    const int rc = btd_sweep_inverse_subset(dtype, D, ld, ls, out_diag, out_sub, B, T, s);
    if (rc != MF_ERR_UNSUPPORTED) return rc;
  }
  return dispatch_small(dtype, D, [&](auto tt, auto dd) {
    using Tp = typename decltype(tt)::type;
    constexpr int kD = decltype(dd)::value;
    btd_inverse_subset_direct_kernel<Tp, kD><<<grid_for(B, 32), 32, 0, s>>>(
        (const Tp*)ld, (const Tp*)ls, (Tp*)out_diag, (Tp*)out_sub, B, T);
    return check_launch();
  });
}

int mf_btd_upper_diagonal_lower(int dtype, const void* diag, const void* sub, void* out_u,
                                void* out_chol_d, int32_t* info, int64_t B, int64_t T, int64_t D,
                                void* stream) {
  if (B < 0 || T < 2 || D < 1) return MF_ERR_BAD_ARG;
  if (B == 0) return MF_OK;
  if (!diag || !sub || !out_u || !out_chol_d) return MF_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  if (mid_dim(D)) return mid_udu(dtype, diag, sub, out_u, out_chol_d, info, B, T, D, s);
  if (D <= kSsmSweepMaxD && tuning(4) != 1) {
    const int rc = btd_sweep_udu(dtype, D, diag, sub, out_u, out_chol_d, info, B, T, s);
    if (rc != MF_ERR_UNSUPPORTED) return rc;
  }
  return dispatch_small(dtype, D, [&](auto tt, auto dd) {
    using Tp = typename decltype(tt)::type;
    constexpr int kD = decltype(dd)::value;
    btd_udu_direct_kernel<Tp, kD><<<grid_for(B, 32), 32, 0, s>>>(
        (const Tp*)diag, (const Tp*)sub, (Tp*)out_u, (Tp*)out_chol_d, info, B, T);
    return check_launch();
  });
}

int mf_btd_dense_mult(int dtype, const void* diag, const void* sub, const void* right, void* out,
                      int64_t n_rhs, int64_t Bm, int64_t T, int64_t D, int transpose,
                      int symmetric, void* stream) {
  if (n_rhs < 0 || Bm < 1 || T < 1 || D < 1) return MF_ERR_BAD_ARG;
  if (n_rhs == 0) return MF_OK;
  if (!diag || !right || !out) return MF_ERR_BAD_ARG;
  if (T == 1) sub = nullptr;
  cudaStream_t s = (cudaStream_t)stream;
  if (mid_dim(D)) return mid_dense_mult(dtype, diag, sub, right, out, n_rhs, Bm, T, D, transpose, symmetric, s);
  return dispatch_small(dtype, D, [&](auto tt, auto dd) {
    using Tp = typename decltype(tt)::type;
    constexpr int kD = decltype(dd)::value;
    btd_dense_mult_kernel<Tp, kD><<<grid_for(n_rhs * T, 128), 128, 0, s>>>(
        (const Tp*)diag, (const Tp*)sub, (const Tp*)right, (Tp*)out, n_rhs, Bm, T, transpose,
        symmetric);
    return check_launch();
  });
}

int mf_btd_abs_log_det(int dtype, const void* ld, void* out, int64_t B, int64_t T, int64_t D,
                       void* stream) {
  if (B < 0 || T < 1 || D < 1) return MF_ERR_BAD_ARG;
  if (B == 0) return MF_OK;
  if (!ld || !out) return MF_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  return dispatch_dtype(dtype, [&](auto tt) {
    using Tp = typename decltype(tt)::type;
    const int64_t seg_len = 8192;
    int64_t nseg = (T + seg_len - 1) / seg_len;
    if (nseg > 1) {
      if (cudaMemsetAsync(out, 0, sizeof(Tp) * B, s) != cudaSuccess) return check_launch();
    }
    for (int64_t b0 = 0; b0 < B; b0 += 65535) {
      const int64_t nb = (B - b0 < 65535) ? B - b0 : 65535;
      dim3 grid((unsigned)nseg, (unsigned)nb);
      btd_abs_log_det_kernel<Tp><<<grid, 256, 0, s>>>((const Tp*)ld + b0 * T * D * D,
                                                      (Tp*)out + b0, T, (int)D, seg_len,
                                                      nseg > 1 ? 1 : 0);
    }
    return check_launch();
  });
}

}  // extern "C"
