// Internal (C++) interface to the mid-size (8 < D <= 32) warp-per-chain implementations (capi_mid.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace mf {

constexpr int kMidMaxD = 32;
inline bool mid_dim(int64_t D) { return D > 8 && D <= kMidMaxD; }

int mid_nat_to_ssm(int dtype, const void* th_lin, const void* th_diag, const void* th_sub, void* out_a,
                   void* out_off, void* out_chol, int32_t* info, int64_t B, int64_t T, int64_t D, int smoothing,
                   cudaStream_t s);
int mid_ssm_to_naturals(int dtype, const void* mu0, const void* chol_p0, const void* a, const void* b,
                        const void* chol_q, void* th_lin, void* th_diag, void* th_sub, int64_t B, int64_t T,
                        int64_t D, int smoothing, cudaStream_t s);
int mid_ssm_moments(int dtype, int expectations, const void* mu0, const void* chol_p0, const void* a,
                    const void* b, const void* chol_q, void* o_vec, void* o_diag, void* o_sub, int64_t B,
                    int64_t T, int64_t D, cudaStream_t s);
int mid_expectations_to_ssm(int dtype, const void* eta_lin, const void* eta_diag, const void* eta_sub,
                            void* out_a, void* out_off, void* out_chol, int32_t* info, int64_t B, int64_t T,
                            int64_t D, cudaStream_t s);

int mid_block_cholesky_or_zero(int dtype, const void* cov, void* out, int32_t* info, int64_t n, int64_t D,
                               cudaStream_t s);
int mid_block_chol_of_inverse(int dtype, const void* chol, void* out, int64_t n, int64_t D, cudaStream_t s);

int mid_build_precision(int dtype, const void* chol_p0, const void* a, const void* chol_q, const void* h,
                        const void* r_inv, void* out_diag, void* out_sub, int64_t B, int64_t T, int64_t D, int64_t m,
                        int64_t h_batch, int64_t r_steps, cudaStream_t s);
int mid_inverse_subset(int dtype, const void* ld, const void* ls, void* out_diag, void* out_sub, int64_t B,
                       int64_t T, int64_t D, cudaStream_t s);
int mid_udu(int dtype, const void* diag, const void* sub, void* out_u, void* out_chol_d, int32_t* info, int64_t B,
            int64_t T, int64_t D, cudaStream_t s);
int mid_affine_scan(int dtype, const void* mu0, const void* chol_p0, const void* a, const void* b,
                    const void* chol_q, const void* eps, void* out, int64_t n, int64_t Bm, int64_t T, int64_t D,
                    cudaStream_t s);
int mid_kalman_log_likelihood(int dtype, const void* mu0, const void* chol_p0, const void* a, const void* b,
                              const void* chol_q, const void* h, const void* obs, const void* chol_r, void* out,
                              int64_t B, int64_t T, int64_t D, int64_t m, int64_t h_batch, int64_t r_steps,
                              cudaStream_t s);

int mid_dense_mult(int dtype, const void* diag, const void* sub, const void* right, void* out, int64_t n_rhs,
                   int64_t Bm, int64_t T, int64_t D, int transpose, int symmetric, cudaStream_t s);
int mid_log_pdf(int dtype, const void* mu0, const void* chol_p0, const void* a, const void* b, const void* chol_q,
                const void* states, void* out, int64_t n, int64_t Bm, int64_t T, int64_t D, cudaStream_t s);

int mid_kl_divergence(int dtype, const void* q_mu0, const void* q_chol_p0, const void* q_a, const void* q_b,
                      const void* q_chol_q, const void* p_mu0, const void* p_chol_p0, const void* p_a,
                      const void* p_b, const void* p_chol_q, void* out, int64_t B, int64_t T, int64_t D,
                      cudaStream_t s);

int mid_pairwise_marginals(int dtype, const void* mean, const void* cov, const void* sub, const void* init_mean,
                           const void* init_cov, int64_t init_batch, void* out_mean, void* out_cov, int64_t B,
                           int64_t T, int64_t D, cudaStream_t s);
int mid_conditional_statistics(int dtype, const void* a_mt, const void* q_mt, const void* a_tp, const void* q_tp,
                               void* out_p, void* out_t, int32_t* info, int return_precision, int64_t N, int64_t D,
                               cudaStream_t s);
int mid_conditional_predict(int dtype, const void* proj, const void* tcov, const void* pair_means,
                            const void* pair_covs, const int64_t* indices, void* out_mean, void* out_cov, int64_t B,
                            int64_t N, int64_t M, int64_t D, cudaStream_t s);

}  // namespace mf
