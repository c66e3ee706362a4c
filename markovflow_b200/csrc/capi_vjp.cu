// C-ABI entry points of the reverse-mode (VJP) kernels (include/markovflow_b200.h).
#include "dispatch.cuh"
#include "vjp_kernels.cuh"

using namespace mf;

extern "C" {

int mf_btd_cholesky_bwd(int dtype, const void* ld, const void* ls, const void* g_ld, const void* g_ls,
                        void* g_diag, void* g_sub, int64_t B, int64_t T, int64_t D, void* stream) {
  if (B < 0 || T < 1 || D < 1) return MF_ERR_BAD_ARG;
  if (B == 0) return MF_OK;
  if (!ld || !g_diag) return MF_ERR_BAD_ARG;
  if (T == 1) ls = nullptr;
  if (ls && !g_sub) return MF_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  return dispatch_small(dtype, D, [&](auto tt, auto dd) {
    using Tp = typename decltype(tt)::type;
    constexpr int kD = decltype(dd)::value;
    btd_cholesky_bwd_kernel<Tp, kD><<<grid_for(B, 32), 32, 0, s>>>(
        (const Tp*)ld, (const Tp*)ls, (const Tp*)g_ld, (const Tp*)g_ls, (Tp*)g_diag, (Tp*)g_sub, B, T);
    return check_launch();
  });
}

int mf_ssm_marginals_bwd(int dtype, const void* chol_p0, const void* a, const void* chol_q, const void* mean,
                         const void* cov, const void* g_mean, const void* g_cov, const void* g_sub,
                         void* g_mu0, void* g_chol_p0, void* g_a, void* g_b, void* g_chol_q, int64_t B,
                         int64_t T, int64_t D, void* stream) {
  if (B < 0 || T < 1 || D < 1) return MF_ERR_BAD_ARG;
  if (B == 0) return MF_OK;
  if (!chol_p0 || !mean || !cov || !g_mu0 || !g_chol_p0) return MF_ERR_BAD_ARG;
  if (T > 1 && (!a || !chol_q || !g_a || !g_b || !g_chol_q)) return MF_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  return dispatch_small(dtype, D, [&](auto tt, auto dd) {
    using Tp = typename decltype(tt)::type;
    constexpr int kD = decltype(dd)::value;
    ssm_marginals_bwd_kernel<Tp, kD><<<grid_for(B, 32), 32, 0, s>>>(
        (const Tp*)chol_p0, (const Tp*)a, (const Tp*)chol_q, (const Tp*)mean, (const Tp*)cov,
        (const Tp*)g_mean, (const Tp*)g_cov, (const Tp*)g_sub, (Tp*)g_mu0, (Tp*)g_chol_p0, (Tp*)g_a,
        (Tp*)g_b, (Tp*)g_chol_q, B, T);
    return check_launch();
  });
}

}  // extern "C"
