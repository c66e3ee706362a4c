// Host side of the mid-size (8 < D <= 32) kernels (mid_kernels.cuh): one warp per chain or per (chain, step),
// blocks in shared memory.  Called from the C-ABI entry points of capi_nat.cu / capi_ssm.cu when the state
// dimension exceeds what one thread's registers hold.
#include "dispatch.cuh"
#include "mid_api.h"
#include "mid_kernels.cuh"

namespace mf {

namespace {

template <class K>
int mid_prepare(SmemOnce& once, K kern, size_t bytes) {
  const cudaError_t e = ensure_smem(once, kern, bytes);
  if (e != cudaSuccess) {
    set_last_error(cudaGetErrorString(e));
    return MF_ERR_CUDA;
  }
  return MF_OK;
}

}  // namespace

int mid_nat_to_ssm(int dtype, const void* th_lin, const void* th_diag, const void* th_sub, void* out_a,
                   void* out_off, void* out_chol, int32_t* info, int64_t B, int64_t T, int64_t D, int smoothing,
                   cudaStream_t s) {
  if (!mid_dim(D)) return MF_ERR_UNSUPPORTED;
  return dispatch_dtype(dtype, [&](auto tt) {
    using Tp = typename decltype(tt)::type;
    if (smoothing) {
      auto kern = mid_nat_to_ssm_kernel<Tp>;
      static SmemOnce once;
      const size_t bytes = mid_smem_bytes(6, 2, sizeof(Tp));
      if (int rc = mid_prepare(once, kern, bytes)) return rc;
      kern<<<(unsigned)B, 32, bytes, s>>>((const Tp*)th_lin, (const Tp*)th_diag, (const Tp*)th_sub, (Tp*)out_a,
                                           (Tp*)out_off, (Tp*)out_chol, info, B, T, (int)D);
    } else {
      if (info && cudaMemsetAsync(info, 0, sizeof(int32_t) * B, s) != cudaSuccess) return check_launch();
      if (B * T > 0x7fffffffLL) return MF_ERR_UNSUPPORTED;
      auto kern = mid_nat_to_ssm_no_smoothing_kernel<Tp>;
      static SmemOnce once;
      const size_t bytes = mid_smem_bytes(4, 2, sizeof(Tp));
      if (int rc = mid_prepare(once, kern, bytes)) return rc;
      kern<<<(unsigned)(B * T), 32, bytes, s>>>((const Tp*)th_lin, (const Tp*)th_diag, (const Tp*)th_sub,
                                                 (Tp*)out_a, (Tp*)out_off, (Tp*)out_chol, info, B, T, (int)D);
    }
    return check_launch();
  });
}

int mid_ssm_to_naturals(int dtype, const void* mu0, const void* chol_p0, const void* a, const void* b,
                        const void* chol_q, void* th_lin, void* th_diag, void* th_sub, int64_t B, int64_t T,
                        int64_t D, int smoothing, cudaStream_t s) {
  if (!mid_dim(D)) return MF_ERR_UNSUPPORTED;
  if (B * T > 0x7fffffffLL) return MF_ERR_UNSUPPORTED;
  return dispatch_dtype(dtype, [&](auto tt) {
    using Tp = typename decltype(tt)::type;
    auto kern = mid_ssm_to_naturals_kernel<Tp>;
    static SmemOnce once;
    const size_t bytes = mid_smem_bytes(5, 1, sizeof(Tp));
    if (int rc = mid_prepare(once, kern, bytes)) return rc;
    kern<<<(unsigned)(B * T), 32, bytes, s>>>((const Tp*)mu0, (const Tp*)chol_p0, (const Tp*)a, (const Tp*)b,
                                               (const Tp*)chol_q, (Tp*)th_lin, (Tp*)th_diag, (Tp*)th_sub, B, T,
                                               (int)D, smoothing);
    return check_launch();
  });
}

int mid_ssm_moments(int dtype, int expectations, const void* mu0, const void* chol_p0, const void* a,
                    const void* b, const void* chol_q, void* o_vec, void* o_diag, void* o_sub, int64_t B,
                    int64_t T, int64_t D, cudaStream_t s) {
  if (!mid_dim(D)) return MF_ERR_UNSUPPORTED;
  return dispatch_dtype(dtype, [&](auto tt) {
    using Tp = typename decltype(tt)::type;
    const size_t bytes = mid_smem_bytes(5, 0, sizeof(Tp));
    if (expectations) {
      auto kern = mid_ssm_moments_kernel<Tp, true>;
      static SmemOnce once;
      if (int rc = mid_prepare(once, kern, bytes)) return rc;
      kern<<<(unsigned)B, 32, bytes, s>>>((const Tp*)mu0, (const Tp*)chol_p0, (const Tp*)a, (const Tp*)b,
                                           (const Tp*)chol_q, (Tp*)o_vec, (Tp*)o_diag, (Tp*)o_sub, B, T, (int)D);
    } else {
      auto kern = mid_ssm_moments_kernel<Tp, false>;
      static SmemOnce once;
      if (int rc = mid_prepare(once, kern, bytes)) return rc;
      kern<<<(unsigned)B, 32, bytes, s>>>((const Tp*)mu0, (const Tp*)chol_p0, (const Tp*)a, (const Tp*)b,
                                           (const Tp*)chol_q, (Tp*)o_vec, (Tp*)o_diag, (Tp*)o_sub, B, T, (int)D);
    }
    return check_launch();
  });
}

int mid_expectations_to_ssm(int dtype, const void* eta_lin, const void* eta_diag, const void* eta_sub,
                            void* out_a, void* out_off, void* out_chol, int32_t* info, int64_t B, int64_t T,
                            int64_t D, cudaStream_t s) {
  if (!mid_dim(D)) return MF_ERR_UNSUPPORTED;
  if (B * T > 0x7fffffffLL) return MF_ERR_UNSUPPORTED;
  return dispatch_dtype(dtype, [&](auto tt) {
    using Tp = typename decltype(tt)::type;
    auto kern = mid_expectations_to_ssm_kernel<Tp>;
    static SmemOnce once;
    const size_t bytes = mid_smem_bytes(6, 1, sizeof(Tp));
    if (int rc = mid_prepare(once, kern, bytes)) return rc;
    kern<<<(unsigned)(B * T), 32, bytes, s>>>((const Tp*)eta_lin, (const Tp*)eta_diag, (const Tp*)eta_sub,
                                               (Tp*)out_a, (Tp*)out_off, (Tp*)out_chol, info, B, T, (int)D);
    return check_launch();
  });
}

int mid_block_cholesky_or_zero(int dtype, const void* cov, void* out, int32_t* info, int64_t n, int64_t D,
                               cudaStream_t s) {
  if (!mid_dim(D)) return MF_ERR_UNSUPPORTED;
  if (n > 0x7fffffffLL) return MF_ERR_UNSUPPORTED;
  return dispatch_dtype(dtype, [&](auto tt) {
    using Tp = typename decltype(tt)::type;
    auto kern = mid_block_cholesky_or_zero_kernel<Tp>;
    static SmemOnce once;
    const size_t bytes = mid_smem_bytes(1, 1, sizeof(Tp));
    if (int rc = mid_prepare(once, kern, bytes)) return rc;
    kern<<<(unsigned)n, 32, bytes, s>>>((const Tp*)cov, (Tp*)out, info, n, (int)D);
    return check_launch();
  });
}

int mid_block_chol_of_inverse(int dtype, const void* chol, void* out, int64_t n, int64_t D, cudaStream_t s) {
  if (!mid_dim(D)) return MF_ERR_UNSUPPORTED;
  if (n > 0x7fffffffLL) return MF_ERR_UNSUPPORTED;
  return dispatch_dtype(dtype, [&](auto tt) {
    using Tp = typename decltype(tt)::type;
    auto kern = mid_block_chol_of_inverse_kernel<Tp>;
    static SmemOnce once;
    const size_t bytes = mid_smem_bytes(3, 1, sizeof(Tp));
    if (int rc = mid_prepare(once, kern, bytes)) return rc;
    kern<<<(unsigned)n, 32, bytes, s>>>((const Tp*)chol, (Tp*)out, n, (int)D);
    return check_launch();
  });
}

// one launch of a warp-per-block kernel with `nmat` matrix slots and `nvec` vector slots of shared memory
#define MF_MID_LAUNCH(KERN, NBLOCKS, NMAT, NVEC, ...)                          \
  do {                                                                         \
    auto kern = KERN<Tp>;                                                      \
    static SmemOnce once;                                                      \
    const size_t bytes = mid_smem_bytes(NMAT, NVEC, sizeof(Tp));               \
    if (int rc = mid_prepare(once, kern, bytes)) return rc;                    \
    kern<<<(unsigned)(NBLOCKS), 32, bytes, s>>>(__VA_ARGS__);                  \
    return check_launch();                                                     \
  } while (0)

int mid_build_precision(int dtype, const void* chol_p0, const void* a, const void* chol_q, const void* h,
                        const void* r_inv, void* out_diag, void* out_sub, int64_t B, int64_t T, int64_t D, int64_t m,
                        int64_t h_batch, int64_t r_steps, cudaStream_t s) {
  if (!mid_dim(D) || B * T > 0x7fffffffLL) return MF_ERR_UNSUPPORTED;
  return dispatch_dtype(dtype, [&](auto tt) {
    using Tp = typename decltype(tt)::type;
    MF_MID_LAUNCH(mid_build_precision_kernel, B * T, 5, 1, (const Tp*)chol_p0, (const Tp*)a, (const Tp*)chol_q,
                  (const Tp*)h, (const Tp*)r_inv, (Tp*)out_diag, (Tp*)out_sub, B, T, (int)D, (int)m, h_batch, r_steps);
  });
}

int mid_inverse_subset(int dtype, const void* ld, const void* ls, void* out_diag, void* out_sub, int64_t B,
                       int64_t T, int64_t D, cudaStream_t s) {
  if (!mid_dim(D)) return MF_ERR_UNSUPPORTED;
  return dispatch_dtype(dtype, [&](auto tt) {
    using Tp = typename decltype(tt)::type;
    MF_MID_LAUNCH(mid_inverse_subset_kernel, B, 6, 1, (const Tp*)ld, (const Tp*)ls, (Tp*)out_diag, (Tp*)out_sub, B, T,
                  (int)D);
  });
}

int mid_udu(int dtype, const void* diag, const void* sub, void* out_u, void* out_chol_d, int32_t* info, int64_t B,
            int64_t T, int64_t D, cudaStream_t s) {
  if (!mid_dim(D)) return MF_ERR_UNSUPPORTED;
  return dispatch_dtype(dtype, [&](auto tt) {
    using Tp = typename decltype(tt)::type;
    MF_MID_LAUNCH(mid_udu_kernel, B, 3, 1, (const Tp*)diag, (const Tp*)sub, (Tp*)out_u, (Tp*)out_chol_d, info, B, T,
                  (int)D);
  });
}

int mid_affine_scan(int dtype, const void* mu0, const void* chol_p0, const void* a, const void* b,
                    const void* chol_q, const void* eps, void* out, int64_t n, int64_t Bm, int64_t T, int64_t D,
                    cudaStream_t s) {
  if (!mid_dim(D) || n > 0x7fffffffLL) return MF_ERR_UNSUPPORTED;
  return dispatch_dtype(dtype, [&](auto tt) {
    using Tp = typename decltype(tt)::type;
    MF_MID_LAUNCH(mid_affine_scan_kernel, n, 2, 0, (const Tp*)mu0, (const Tp*)chol_p0, (const Tp*)a, (const Tp*)b,
                  (const Tp*)chol_q, (const Tp*)eps, (Tp*)out, n, Bm, T, (int)D);
  });
}

int mid_kalman_log_likelihood(int dtype, const void* mu0, const void* chol_p0, const void* a, const void* b,
                              const void* chol_q, const void* h, const void* obs, const void* chol_r, void* out,
                              int64_t B, int64_t T, int64_t D, int64_t m, int64_t h_batch, int64_t r_steps,
                              cudaStream_t s) {
  if (!mid_dim(D) || B > 0x7fffffffLL) return MF_ERR_UNSUPPORTED;
  return dispatch_dtype(dtype, [&](auto tt) {
    using Tp = typename decltype(tt)::type;
    MF_MID_LAUNCH(mid_kalman_loglik_kernel, B, 4, 0, (const Tp*)mu0, (const Tp*)chol_p0, (const Tp*)a, (const Tp*)b,
                  (const Tp*)chol_q, (const Tp*)h, (const Tp*)obs, (const Tp*)chol_r, (Tp*)out, B, T, (int)D, (int)m,
                  h_batch, r_steps);
  });
}

int mid_dense_mult(int dtype, const void* diag, const void* sub, const void* right, void* out, int64_t n_rhs,
                   int64_t Bm, int64_t T, int64_t D, int transpose, int symmetric, cudaStream_t s) {
  if (!mid_dim(D) || n_rhs * T > 0x7fffffffLL) return MF_ERR_UNSUPPORTED;
  return dispatch_dtype(dtype, [&](auto tt) {
    using Tp = typename decltype(tt)::type;
    MF_MID_LAUNCH(mid_dense_mult_kernel, n_rhs * T, 1, 0, (const Tp*)diag, (const Tp*)sub, (const Tp*)right, (Tp*)out,
                  n_rhs, Bm, T, (int)D, transpose, symmetric);
  });
}

int mid_log_pdf(int dtype, const void* mu0, const void* chol_p0, const void* a, const void* b, const void* chol_q,
                const void* states, void* out, int64_t n, int64_t Bm, int64_t T, int64_t D, cudaStream_t s) {
  if (!mid_dim(D) || n > 0x7fffffffLL) return MF_ERR_UNSUPPORTED;
  return dispatch_dtype(dtype, [&](auto tt) {
    using Tp = typename decltype(tt)::type;
    MF_MID_LAUNCH(mid_log_pdf_kernel, n, 2, 1, (const Tp*)mu0, (const Tp*)chol_p0, (const Tp*)a, (const Tp*)b,
                  (const Tp*)chol_q, (const Tp*)states, (Tp*)out, n, Bm, T, (int)D);
  });
}

int mid_kl_divergence(int dtype, const void* q_mu0, const void* q_chol_p0, const void* q_a, const void* q_b,
                      const void* q_chol_q, const void* p_mu0, const void* p_chol_p0, const void* p_a,
                      const void* p_b, const void* p_chol_q, void* out, int64_t B, int64_t T, int64_t D,
                      cudaStream_t s) {
  if (!mid_dim(D) || B > 0x7fffffffLL) return MF_ERR_UNSUPPORTED;
  return dispatch_dtype(dtype, [&](auto tt) {
    using Tp = typename decltype(tt)::type;
    MF_MID_LAUNCH(mid_kl_kernel, B, 8, 1, (const Tp*)q_mu0, (const Tp*)q_chol_p0, (const Tp*)q_a, (const Tp*)q_b,
                  (const Tp*)q_chol_q, (const Tp*)p_mu0, (const Tp*)p_chol_p0, (const Tp*)p_a, (const Tp*)p_b,
                  (const Tp*)p_chol_q, (Tp*)out, B, T, (int)D);
  });
}

int mid_pairwise_marginals(int dtype, const void* mean, const void* cov, const void* sub, const void* init_mean,
                           const void* init_cov, int64_t init_batch, void* out_mean, void* out_cov, int64_t B,
                           int64_t T, int64_t D, cudaStream_t s) {
  if (!mid_dim(D) || B * (T + 1) > 0x7fffffffLL) return MF_ERR_UNSUPPORTED;
  return dispatch_dtype(dtype, [&](auto tt) {
    using Tp = typename decltype(tt)::type;
    mid_pairwise_marginals_kernel<Tp><<<(unsigned)(B * (T + 1)), 128, 0, s>>>(
        (const Tp*)mean, (const Tp*)cov, (const Tp*)sub, (const Tp*)init_mean, (const Tp*)init_cov, init_batch,
        (Tp*)out_mean, (Tp*)out_cov, B, T, (int)D);
    return check_launch();
  });
}

int mid_conditional_statistics(int dtype, const void* a_mt, const void* q_mt, const void* a_tp, const void* q_tp,
                               void* out_p, void* out_t, int32_t* info, int return_precision, int64_t N, int64_t D,
                               cudaStream_t s) {
  if (!mid_dim(D) || N > 0x7fffffffLL) return MF_ERR_UNSUPPORTED;
  return dispatch_dtype(dtype, [&](auto tt) {
    using Tp = typename decltype(tt)::type;
    MF_MID_LAUNCH(mid_conditional_statistics_kernel, N, 9, 2, (const Tp*)a_mt, (const Tp*)q_mt, (const Tp*)a_tp,
                  (const Tp*)q_tp, (Tp*)out_p, (Tp*)out_t, info, return_precision, N, (int)D);
  });
}

int mid_conditional_predict(int dtype, const void* proj, const void* tcov, const void* pair_means,
                            const void* pair_covs, const int64_t* indices, void* out_mean, void* out_cov, int64_t B,
                            int64_t N, int64_t M, int64_t D, cudaStream_t s) {
  if (!mid_dim(D) || B * N > 0x7fffffffLL) return MF_ERR_UNSUPPORTED;
  return dispatch_dtype(dtype, [&](auto tt) {
    using Tp = typename decltype(tt)::type;
    MF_MID_LAUNCH(mid_conditional_predict_kernel, B * N, 5, 0, (const Tp*)proj, (const Tp*)tcov,
                  (const Tp*)pair_means, (const Tp*)pair_covs, indices, (Tp*)out_mean, (Tp*)out_cov, B, N, M, (int)D);
  });
}

#undef MF_MID_LAUNCH

}  // namespace mf
