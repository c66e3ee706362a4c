/*
 * markovflow_b200 -- C ABI of the B200-native structured linear-algebra hot path of Markovflow.
 *
 * This is the drop-in boundary.  The reference has no FFI layer in-tree: its seam is the Python
 * class API of markovflow/block_tri_diag.py and, beneath it, the TensorFlow custom ops of the
 * third-party package banded-matrices==0.0.6 (poetry.lock:133-144).  Each entry point below
 * replaces one such interface and cites it.  All entry points:
 *
 *   - take raw DEVICE pointers to C-contiguous arrays in the reference's public block layout
 *     (batch-major, chain-contiguous):  diag [B,T,D,D], sub [B,T-1,D,D], vectors [B,T,D];
 *   - never allocate, free or retain caller memory; outputs are caller-allocated;
 *   - are asynchronous on the cudaStream_t passed as `stream` (void* so the header needs no CUDA);
 *   - return MF_OK or an MF_ERR_* status; numerical failure (non-positive pivot) is reported per
 *     chain through `info[b]` = 1-based block index of the first failing pivot (0 = success),
 *     LAPACK convention, replacing TF's "Banded Cholesky decomposition failure" error;
 *   - support dtype MF_F32 / MF_F64 and 1 <= D <= MF_SMALL_D_MAX (thread-per-chain register
 *     kernels); the entry points listed at MF_BIG_D_MAX below also take MF_SMALL_D_MAX < D <=
 *     MF_BIG_D_MAX (one warp or half-warp per chain) -- MF_ERR_UNSUPPORTED otherwise.
 *
 * The mf_host_* variants (end of this header) take HOST pointers and perform the host<->device
 * copies themselves (chunked and pipelined on internal streams); they are what a host-resident
 * framework (TensorFlow CPU tensors) would bind.
 */
#ifndef MARKOVFLOW_B200_H
#define MARKOVFLOW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MF_OK 0
#define MF_ERR_BAD_ARG 1      /* null pointer / non-positive dimension / unknown dtype */
#define MF_ERR_UNSUPPORTED 2  /* dimension outside what the build supports */
#define MF_ERR_CUDA 3         /* a CUDA runtime call failed; see mf_last_cuda_error() */

#define MF_F32 0
#define MF_F64 1

#define MF_SMALL_D_MAX 8 /* thread-per-chain register kernels: every entry point */
#define MF_BIG_D_MAX 32  /* warp-per-chain kernels (blocks in shared memory / spread over lanes): mf_btd_cholesky,
                          * mf_btd_solve, mf_btd_inverse_subset, mf_btd_upper_diagonal_lower, mf_btd_dense_mult,
                          * mf_btd_abs_log_det, mf_ssm_build_precision, mf_ssm_marginals, mf_ssm_affine_scan,
                          * mf_ssm_log_pdf, mf_ssm_kl_divergence, mf_nat_to_ssm, mf_ssm_to_naturals, mf_ssm_to_expectations,
                          * mf_expectations_to_ssm, mf_block_cholesky_or_zero, mf_block_chol_of_inverse,
                          * mf_pairwise_marginals, mf_conditional_statistics, mf_conditional_predict,
                          * mf_kalman_log_likelihood */

/* Library / build information. */
int mf_version(void);
const char* mf_last_cuda_error(void);
/* Development/tuning knobs (kernel variant selection); knob 0: Cholesky sweep variant
 * (0 auto = TMA bulk-copy ring, 1 direct global-memory streaming, 2 cp.async element-staged ring),
 * knob 1: steps per stage of variant 2; knob 2: Kalman log-likelihood path (0 auto, 1 one thread
 * per chain, 2 parallel-in-time; also: 1 = no parallel-in-time evaluation of the moment recursions);
 * knob 3: steps per parallel-in-time segment (0 auto); knob 4: 1 = direct-load kernels instead of
 * the TMA chain sweeps; knob 7: large-block Cholesky (0/1 one warp per chain, 2 experimental
 * one-CTA-per-chain kernel); knob 8: Matern-prior Kalman kernel (0 all-warps-compute kernel,
 * 1..4 TMA chain-sweep geometries, 5/6 register-capped variants); knob 9: its virtual chains per SM, in warps (0 auto); knob 10: segments per chain from which
 * the parallel-in-time seed folds run as warp scans (0 auto = 8); knob 11: 1 = the default
 * 8-step tiles everywhere (no 16-step tiles for output-less float32 sweeps, no 4-step tiles for the
 * float64 D = 2 naturals -> SSM sweep); knob 7 (current meaning): large-block Cholesky for D <= 17 -- 0 auto
 * (half-warp-per-chain kernels, parallel in time for few chains), 1 never cut chains into segments, n >= 3
 * that many segments, 2 the round-1 warp-per-chain kernel; knob 12: 1 = ring producers wait for their element
 * cp.async themselves and arrive with a plain mbarrier arrive (the completion path compute-sanitizer's
 * racecheck models; default: cp.async.mbarrier.arrive.noinc); knob 13: 1 = no tensor-map (cp.async.bulk.tensor)
 * sweeps, always one 1-D bulk copy per chain, 2 = tensor-map sweeps also below their minimum problem size; knob 14: tensor-map tile geometry (0 auto: 4-step tiles, 1: 8-step,
 * 2: 4-step with one output stage, 3: 2-step, 4: 4-step; auto = 4-step tiles in float64, 8-step in float32). */
int mf_set_tuning(int knob, int value);
/* Number of sweeps launched on the tensor-map engine (csrc/sweep_tm.cuh) since the library was loaded; the tests
 * use it to check which engine served a call. */
int64_t mf_tm_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * Block-tridiagonal operators (markovflow/block_tri_diag.py)
 * ------------------------------------------------------------------------------------------- */

/* SymmetricBlockTriDiagonal.cholesky (block_tri_diag.py:423-436; banded op cholesky_band :436 plus
 * the block<->band repacks :206-237,549-592), optionally fused with the forward substitution
 * LowerTriangularBlockTriDiagonal.solve (:339-351; solve_triang_mat :350) and abs_log_det
 * (:353-366).  Reads only the lower triangle of each diagonal block.
 *   diag [B,T,D,D], sub [B,T-1,D,D] or NULL, rhs [B,T,D] or NULL
 *   out_diag [B,T,D,D] (upper triangles written as 0), out_sub [B,T-1,D,D] (NULL iff sub NULL),
 *   out_x [B,T,D] (NULL iff rhs NULL), out_logdet [B] or NULL (= sum log diag L), info [B].
 * out_* may alias the corresponding inputs (in-place factorisation). */
int mf_btd_cholesky(int dtype, const void* diag, const void* sub, const void* rhs, void* out_diag,
                    void* out_sub, void* out_x, void* out_logdet, int32_t* info, int64_t B,
                    int64_t T, int64_t D, void* stream);

/* LowerTriangularBlockTriDiagonal.solve (block_tri_diag.py:339-351; solve_triang_mat :350).
 *   ld [Bm,T,D,D] (lower triangle read; NULL = identity diagonal blocks, the a_inv_block of
 *   state_space_model.py:278-296), ls [Bm,T-1,D,D] or NULL, rhs/out [n_rhs,T,D].
 * Right-hand side chain c uses matrix chain c % Bm (leading sample dims of `right`,
 * block_tri_diag.py:261-287).  transpose != 0 solves with L^T (runs backwards in time). */
int mf_btd_solve(int dtype, const void* ld, const void* ls, const void* rhs, void* out,
                 int64_t n_rhs, int64_t Bm, int64_t T, int64_t D, int transpose, void* stream);

/* LowerTriangularBlockTriDiagonal.block_diagonal_of_inverse (block_tri_diag.py:318-337;
 * inverse_from_cholesky_band :331) and, with out_sub != NULL, the sub-diagonal blocks
 * Sigma_{k+1,k} that naturals_to_ssm_params needs (ssm_gaussian_transformations.py:443-458).
 *   out_diag [B,T,D,D] full symmetric blocks, out_sub [B,T-1,D,D] or NULL. */
int mf_btd_inverse_subset(int dtype, const void* ld, const void* ls, void* out_diag,
                          void* out_sub, int64_t B, int64_t T, int64_t D, void* stream);

/* SymmetricBlockTriDiagonal.upper_diagonal_lower (block_tri_diag.py:438-545).
 *   out_u [B,T-1,D,D] = sub-diagonal of U^T (identity diagonal implied),
 *   out_chol_d [B,T,D,D] = Cholesky factors of the block-diagonal D. */
int mf_btd_upper_diagonal_lower(int dtype, const void* diag, const void* sub, void* out_u,
                                void* out_chol_d, int32_t* info, int64_t B, int64_t T, int64_t D,
                                void* stream);

/* BlockTriDiagonal.dense_mult (block_tri_diag.py:175-199; product_band_mat :189).
 * symmetric != 0 treats the matrix as SymmetricBlockTriDiagonal (lower triangle mirrored);
 * otherwise it is lower triangular (upper triangles of diagonal blocks ignored) and
 * transpose != 0 multiplies by its transpose. */
int mf_btd_dense_mult(int dtype, const void* diag, const void* sub, const void* right, void* out,
                      int64_t n_rhs, int64_t Bm, int64_t T, int64_t D, int transpose,
                      int symmetric, void* stream);

/* LowerTriangularBlockTriDiagonal.abs_log_det (block_tri_diag.py:353-366): out [B]. */
int mf_btd_abs_log_det(int dtype, const void* ld, void* out, int64_t B, int64_t T, int64_t D,
                       void* stream);

/* ---------------------------------------------------------------------------------------------
 * StateSpaceModel operators (markovflow/state_space_model.py).  Parameter layout of the reference
 * (state_space_model.py:74-122), chain-contiguous, T = num_transitions + 1 states:
 *   mu0 [B,D], chol_p0 [B,D,D], a [B,T-1,D,D], b [B,T-1,D], chol_q [B,T-1,D,D]
 * ------------------------------------------------------------------------------------------- */

/* StateSpaceModel._build_precision (state_space_model.py:431-483): out_diag [B,T,D,D] (full
 * symmetric blocks), out_sub [B,T-1,D,D].  With h != NULL the likelihood precision H^T R^-1 H of
 * BaseKalmanFilter._k_inv_post (kalman_filter.py:85-101) is added in the same pass:
 *   h [h_batch,T,m,D] (h_batch = 1 or B), r_inv [r_steps,m,m] (r_steps = 1 or T). */
int mf_ssm_build_precision(int dtype, const void* chol_p0, const void* a, const void* chol_q,
                           const void* h, const void* r_inv, void* out_diag, void* out_sub,
                           int64_t B, int64_t T, int64_t D, int64_t m, int64_t h_batch,
                           int64_t r_steps, void* stream);

/* The affine recurrence behind marginal_means (state_space_model.py:231-251) and sample
 * (:298-324), i.e. a_inv_block.solve(.) (:278-296):
 *   x_0 = mu0 + chol_p0 eps_0,   x_k = a_{k-1} x_{k-1} + b_{k-1} + chol_q_{k-1} eps_k
 * eps [n,T,D] standard-normal draws, or NULL for the means.  out [n,T,D]; trajectory c uses SSM
 * chain c % Bm (leading sample dims). */
int mf_ssm_affine_scan(int dtype, const void* mu0, const void* chol_p0, const void* a,
                       const void* b, const void* chol_q, const void* eps, void* out, int64_t n,
                       int64_t Bm, int64_t T, int64_t D, void* stream);

/* StateSpaceModel.sample (state_space_model.py:298-324) with the standard normals drawn INSIDE the sweep
 * (SURVEY.md 8f-4): Philox4x32-10 keyed by (seed, trajectory, step), so the n*T*D draws are never written to
 * or read from memory and do not depend on how the launch is cut.  out [n,T,D]; trajectory c uses SSM chain
 * c % Bm.  mf_philox_normal writes the same stream out, eps [n,T,D]: mf_ssm_affine_scan on it reproduces the
 * sample (the reference's tf.random.normal stream cannot be pinned by any port). */
int mf_ssm_sample(int dtype, const void* mu0, const void* chol_p0, const void* a, const void* b,
                  const void* chol_q, uint64_t seed, void* out, int64_t n, int64_t Bm, int64_t T,
                  int64_t D, void* stream);
int mf_philox_normal(int dtype, uint64_t seed, void* out, int64_t n, int64_t T, int64_t D, void* stream);

/* marginal_means / marginal_covariances / covariance_blocks / subsequent_covariances
 * (state_space_model.py:231-275,326-341) in one forward sweep:  any of out_mean [B,T,D],
 * out_cov [B,T,D,D], out_sub [B,T-1,D,D] (= a_k Sigma_kk) may be NULL. */
int mf_ssm_marginals(int dtype, const void* mu0, const void* chol_p0, const void* a, const void* b,
                     const void* chol_q, void* out_mean, void* out_cov, void* out_sub, int64_t B,
                     int64_t T, int64_t D, void* stream);

/* StateSpaceModel.log_pdf (state_space_model.py:485-526): states [n,T,D] -> out [n]; trajectory c
 * is scored under SSM chain c % Bm. */
int mf_ssm_log_pdf(int dtype, const void* mu0, const void* chol_p0, const void* a, const void* b,
                   const void* chol_q, const void* states, void* out, int64_t n, int64_t Bm,
                   int64_t T, int64_t D, void* stream);

/* StateSpaceModel.kl_divergence (state_space_model.py:528-593): KL(q || p) -> out [B]. */
int mf_ssm_kl_divergence(int dtype, const void* q_mu0, const void* q_chol_p0, const void* q_a,
                         const void* q_b, const void* q_chol_q, const void* p_mu0,
                         const void* p_chol_p0, const void* p_a, const void* p_b,
                         const void* p_chol_q, void* out, int64_t B, int64_t T, int64_t D,
                         void* stream);

/* state_space_model_from_covariances.cholesky_or_zero (state_space_model.py:634-656): Cholesky of
 * n blocks [n,D,D]; all-zero blocks map to zero.  info (one int32, may be NULL) receives 1 + the
 * largest index of a block that is not positive definite, 0 if none. */
int mf_block_cholesky_or_zero(int dtype, const void* cov, void* out, int32_t* info, int64_t n,
                              int64_t D, void* stream);

/* out_k = chol((L_k L_k^T)^-1) for n lower factors [n,D,D]
 * (posterior_state_space_model, kalman_filter.py:170-174). */
int mf_block_chol_of_inverse(int dtype, const void* chol, void* out, int64_t n, int64_t D,
                             void* stream);

/* ---------------------------------------------------------------------------------------------
 * Kalman filter (markovflow/kalman_filter.py)
 *   h [h_batch,T,m,D] emission matrices (h_batch = 1 or B), obs [B,T,m],
 *   chol_r [r_steps,m,m] Cholesky factor(s) of the observation covariance (r_steps = 1:
 *   KalmanFilter, kalman_filter.py:301-348; r_steps = T: KalmanFilterWithSites :436-497).
 *   m <= 4.  With m = 1 an infinite chol_r entry marks a step without observation
 *   (KalmanFilterWithSparseSites :500-626).
 * ------------------------------------------------------------------------------------------- */

/* Scratch needed by the parallel-in-time path for B chains of T steps (bytes; 0 on bad input). */
size_t mf_kalman_workspace_bytes(int dtype, int64_t B, int64_t T, int64_t D);

/* BaseKalmanFilter.log_likelihood (kalman_filter.py:184-255), per chain: out [B] (the reference
 * returns their sum).  Many chains: one thread per chain.  Few long chains (and a workspace of at
 * least mf_kalman_workspace_bytes): parallel-in-time scan.  workspace may be NULL. */
int mf_kalman_log_likelihood(int dtype, const void* mu0, const void* chol_p0, const void* a,
                             const void* b, const void* chol_q, const void* h, const void* obs,
                             const void* chol_r, void* out, int64_t B, int64_t T, int64_t D,
                             int64_t m, int64_t h_batch, int64_t r_steps, void* workspace,
                             size_t workspace_bytes, void* stream);

/* Time-sharded evaluation of ONE long series over several GPUs (each rank holds a contiguous
 * segment of T local steps).  first_is_initial = 1: the segment starts at the prior (mu0, chol_p0)
 * and a/b/chol_q hold T-1 transitions; first_is_initial = 0: a/b/chol_q hold T transitions, a[k]
 * leading INTO local step k (mu0/chol_p0 unused).
 *   1. mf_kalman_segment_summary    -> out_elem [B, 3D^2+2D+1]: the segment's scan element
 *      (A, b, C, eta, J, ell) (Sarkka & Garcia-Fernandez 2021); leaves per-thread summaries in workspace
 *   2. all-gather the elements over ranks (NCCL), mf_kalman_fold_elements joins those of the
 *      earlier ranks: elems [n,B,3D^2+2D+1] -> out [B,3D^2+2D+1]
 *   3. mf_kalman_log_likelihood_seeded  -> out [B]: this segment's share of the log-likelihood,
 *      given prefix_elem (NULL on the first rank); summaries_valid = 1 reuses step 1's workspace. */
int mf_kalman_segment_summary(int dtype, const void* mu0, const void* chol_p0, const void* a,
                              const void* b, const void* chol_q, const void* h, const void* obs,
                              const void* chol_r, void* out_elem, int64_t B, int64_t T, int64_t D,
                              int64_t m, int64_t h_batch, int64_t r_steps, int first_is_initial,
                              void* workspace, size_t workspace_bytes, void* stream);
int mf_kalman_fold_elements(int dtype, const void* elems, void* out, int64_t n, int64_t B,
                            int64_t D, void* stream);
int mf_kalman_log_likelihood_seeded(int dtype, const void* mu0, const void* chol_p0, const void* a,
                                    const void* b, const void* chol_q, const void* h,
                                    const void* obs, const void* chol_r, const void* prefix_elem,
                                    void* out, int64_t B, int64_t T, int64_t D, int64_t m,
                                    int64_t h_batch, int64_t r_steps, int first_is_initial,
                                    int summaries_valid, void* workspace, size_t workspace_bytes,
                                    void* stream);

/* The same protocol in ONE call per rank with the exchange inside the reduction kernel: every rank owns a
 * peer-mapped region (mf_peer_alloc / mf_peer_open: cudaIpc, NVLink peer-to-peer) of
 * mf_kalman_peer_region_bytes; the thread that holds a series' local element stores it into EVERY rank's
 * region, raises a flag, waits for the flags of all ranks and joins the `world` elements in rank (= time) order.
 * No NCCL collective, no extra launch.  peer_regions[r] = rank r's region as mapped in this process
 * (peer_regions[rank] = the own allocation); epoch = call counter (> 0, the same on every rank; regions are
 * double-buffered on its parity).  out [B]: log-likelihood of the WHOLE series (the same on every rank),
 * out_elem [B, 3D^2+2D+1]: its joined element.  m = 1 and D <= 4 (MF_ERR_UNSUPPORTED otherwise); world <= 8. */
size_t mf_kalman_peer_region_bytes(int dtype, int64_t B, int64_t D, int world);
int mf_peer_alloc(size_t bytes, void** ptr, unsigned char* handle64);  /* zeroed device memory + its IPC handle */
int mf_peer_open(const unsigned char* handle64, void** ptr);           /* map a peer's region */
int mf_peer_close(void* ptr);
int mf_peer_free(void* ptr);
int mf_kalman_time_sharded_log_likelihood(int dtype, const void* mu0, const void* chol_p0, const void* a,
                                          const void* b, const void* chol_q, const void* h,
                                          const void* obs, const void* chol_r, void* out, void* out_elem,
                                          int64_t B, int64_t T, int64_t D, int64_t m, int64_t h_batch,
                                          int64_t r_steps, int first_is_initial, void* const* peer_regions,
                                          int rank, int world, uint64_t epoch, void* workspace,
                                          size_t workspace_bytes, void* stream);

/* KalmanFilter.log_likelihood of a stationary Matern prior with the state-space
 * model built INSIDE the kernel from the time deltas.  Replaces, in one launch (+ one reduction),
 *   SDEKernel.state_space_model (kernels/sde_kernel.py:153-171)
 *     -> StationaryKernel.transition_statistics (:421-446,  Q_k = Pinf - A_k Pinf A_k^T + jitter I)
 *     -> Matern12/32/52.state_transitions (kernels/matern.py:80-86, :299-324, :434-460)
 *   SDEKernel.generate_emission_model (sde_kernel.py:173-211,  H = [1, 0, ...])
 *   KalmanFilter.log_likelihood (kalman_filter.py:184-255)
 * so that a step reads two values (dt_k, y_k) instead of the 2D^2+2D+1 of the materialised SSM.
 *   D = 1: Matern12, 2: Matern32, 3: Matern52 (state_dim);  zero mean, output_dim 1.
 *   lengthscale [B], variance [B] (per series), jitter (the kernel's jitter, sde_kernel.py:83-96),
 *   time_deltas [B, T - first_is_initial], obs [B,T], chol_r [1].
 *   first_is_initial = 1: the series starts at the stationary prior N(0, Pinf + jitter I);
 *   first_is_initial = 0: a later time segment of a longer series (time_deltas[k] leads INTO local
 *   step k), only out_elem is meaningful -- the time-sharded protocol of mf_kalman_segment_summary.
 *   out [B] log-likelihoods (may be NULL), out_elem [B, 3D^2+2D+1] scan element (A, b, C, eta, J, ell)
 *   of the whole segment (may be NULL; join with mf_kalman_fold_elements).
 *   workspace: mf_kalman_matern_workspace_bytes; without it few long series are not cut in time. */
size_t mf_kalman_matern_workspace_bytes(int dtype, int64_t B, int64_t T, int64_t D);
int mf_kalman_matern_log_likelihood(int dtype, const void* lengthscale, const void* variance,
                                    double jitter, const void* time_deltas, const void* obs,
                                    const void* chol_r, void* out, void* out_elem, int64_t B,
                                    int64_t T, int64_t D, int first_is_initial, void* workspace,
                                    size_t workspace_bytes, void* stream);

/* mf_kalman_matern_log_likelihood of ONE time segment per rank with the exchange of the segment elements inside
 * the reduction kernel (see mf_kalman_time_sharded_log_likelihood): out [B] = log-likelihood of the whole series. */
int mf_kalman_matern_time_sharded_log_likelihood(int dtype, const void* lengthscale, const void* variance,
                                                 double jitter, const void* time_deltas, const void* obs,
                                                 const void* chol_r, void* out, void* out_elem, int64_t B,
                                                 int64_t T, int64_t D, int first_is_initial,
                                                 void* const* peer_regions, int rank, int world,
                                                 uint64_t epoch, void* workspace, size_t workspace_bytes,
                                                 void* stream);

/* ---------------------------------------------------------------------------------------------
 * Natural / expectation parameter transforms (markovflow/ssm_gaussian_transformations.py)
 *   naturals:     theta_lin [B,T,D], theta_diag [B,T,D,D], theta_sub [B,T-1,D,D]
 *   expectations: eta_lin, eta_diag, eta_sub with the same shapes
 *   SSM outputs in the concatenated layout the reference slices: out_a [B,T-1,D,D],
 *   out_offsets [B,T,D] = [mu0, b_1, ...], out_chols [B,T,D,D] = [chol P0, chol Q_1, ...];
 *   info [B] = 1-based step of the first non-positive pivot (0 = ok).
 * ------------------------------------------------------------------------------------------- */

/* naturals_to_ssm_params (ssm_gaussian_transformations.py:332-511), smoothing != 0, as one
 * backward U D U^T sweep; naturals_to_ssm_params_no_smoothing (:514-593), smoothing == 0. */
int mf_nat_to_ssm(int dtype, const void* theta_lin, const void* theta_diag, const void* theta_sub,
                  void* out_a, void* out_offsets, void* out_chols, int32_t* info, int64_t B,
                  int64_t T, int64_t D, int smoothing, void* stream);

/* ssm_to_naturals (:181-253) / ssm_to_naturals_no_smoothing (:256-329). */
int mf_ssm_to_naturals(int dtype, const void* mu0, const void* chol_p0, const void* a,
                       const void* b, const void* chol_q, void* theta_lin, void* theta_diag,
                       void* theta_sub, int64_t B, int64_t T, int64_t D, int smoothing,
                       void* stream);

/* ssm_to_expectations (:31-89): one forward sweep. */
int mf_ssm_to_expectations(int dtype, const void* mu0, const void* chol_p0, const void* a,
                           const void* b, const void* chol_q, void* eta_lin, void* eta_diag,
                           void* eta_sub, int64_t B, int64_t T, int64_t D, void* stream);

/* expectations_to_ssm_params (:92-178). */
int mf_expectations_to_ssm(int dtype, const void* eta_lin, const void* eta_diag,
                           const void* eta_sub, void* out_a, void* out_offsets, void* out_chols,
                           int32_t* info, int64_t B, int64_t T, int64_t D, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Conditionals on top of the marginals (markovflow/conditionals.py) -- SURVEY.md 8f-1
 * ------------------------------------------------------------------------------------------- */

/* pairwise_marginals (conditionals.py:423-485): joint mean / covariance of every pair of subsequent
 * states (x_{k-1}, x_k), k = 0..T, with the initial (prior) state at both ends.
 *   mean [B,T,D], cov [B,T,D,D], sub [B,T-1,D,D] = A_k Sigma_kk (as written by mf_ssm_marginals),
 *   init_mean [init_batch,D], init_cov [init_batch,D,D] (init_batch = 1 or B),
 *   out_mean [B,T+1,2D], out_cov [B,T+1,2D,2D]. */
int mf_pairwise_marginals(int dtype, const void* mean, const void* cov, const void* sub,
                          const void* init_mean, const void* init_cov, int64_t init_batch,
                          void* out_mean, void* out_cov, int64_t B, int64_t T, int64_t D, void* stream);

/* _conditional_statistics_from_transitions (conditionals.py:128-205): p(x_t | x_-, x_+) =
 * N(D_t x_- + E_t x_+, T_t) from the transitions into t (a_mt, q_mt) and out of t (a_tp, q_tp),
 * all [N,D,D].  out_p [N,D,2D] = [D_t | E_t], out_t [N,D,D] = T_t (the precision T_t^-1 when
 * return_precision != 0), info [N] or NULL (1 where a Cholesky pivot was not positive). */
int mf_conditional_statistics(int dtype, const void* a_mt, const void* q_mt, const void* a_tp,
                              const void* q_tp, void* out_p, void* out_t, int32_t* info,
                              int return_precision, int64_t N, int64_t D, void* stream);

/* conditional_predict / base_conditional_predict (conditionals.py:29-83, 380-420):
 *   mean = P m[idx], cov = T (+ P S[idx] P^T when pair_covs != NULL), the gather of the pairwise
 *   marginals by insertion index fused in.  proj [B,N,D,2D], tcov [B,N,D,D], pair_means [B,M,2D],
 *   pair_covs [B,M,2D,2D] or NULL, indices [B,N] (int64; NULL = identity, needs N == M),
 *   out_mean [B,N,D], out_cov [B,N,D,D]. */
int mf_conditional_predict(int dtype, const void* proj, const void* tcov, const void* pair_means,
                           const void* pair_covs, const int64_t* indices, void* out_mean,
                           void* out_cov, int64_t B, int64_t N, int64_t M, int64_t D, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Reverse mode (vector-Jacobian products) -- SURVEY.md 8f-3.  The reference differentiates through
 * every banded op with TensorFlow's tape (gradients registered by banded-matrices; callers
 * ssm_natgrad.py:142-172, tests/integration/models/test_gaussian_process_regression.py:117-130).
 * Each recurrence has an adjoint sweep that walks the chain in the opposite direction.
 * ------------------------------------------------------------------------------------------- */

/* Adjoint of mf_btd_cholesky: factor ld [B,T,D,D], ls [B,T-1,D,D] (or NULL) and the adjoints of its
 * blocks g_ld (lower triangles read; NULL = 0), g_ls (NULL = 0)  ->  g_diag [B,T,D,D], the gradient
 * with respect to the entries the forward pass READS (lower triangles; an off-diagonal entry stands
 * for both symmetric positions; upper triangles are written as 0), and g_sub [B,T-1,D,D]. */
int mf_btd_cholesky_bwd(int dtype, const void* ld, const void* ls, const void* g_ld, const void* g_ls,
                        void* g_diag, void* g_sub, int64_t B, int64_t T, int64_t D, void* stream);

/* Adjoint of mf_ssm_marginals: SSM parameters, the forward results mean [B,T,D], cov [B,T,D,D] and
 * the adjoints g_mean, g_cov, g_sub (each may be NULL)  ->  g_mu0 [B,D], g_chol_p0 [B,D,D],
 * g_a [B,T-1,D,D], g_b [B,T-1,D], g_chol_q [B,T-1,D,D] (lower triangles). */
int mf_ssm_marginals_bwd(int dtype, const void* chol_p0, const void* a, const void* chol_q,
                         const void* mean, const void* cov, const void* g_mean, const void* g_cov,
                         const void* g_sub, void* g_mu0, void* g_chol_p0, void* g_a, void* g_b,
                         void* g_chol_q, int64_t B, int64_t T, int64_t D, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Host-buffer variants: same operators, arrays in HOST memory (pinned for asynchronous copies;
 * pageable memory works, synchronously).  The library stages chunks through per-device slots it
 * keeps between calls (freed by mf_host_release), overlapping host->device copies, the sweeps and
 * device->host copies on three internal streams; the call returns when the results are in host
 * memory.  device < 0: the current device.  bytes_moved (may be NULL) receives
 * {host->device bytes, device->host bytes}.
 * ------------------------------------------------------------------------------------------- */

/* mf_btd_cholesky on host arrays, `chunk` chains per stage (<= 0: 128). */
int mf_host_btd_cholesky(int dtype, const void* diag, const void* sub, const void* rhs,
                         void* out_diag, void* out_sub, void* out_x, void* out_logdet,
                         int32_t* info, int64_t B, int64_t T, int64_t D, int64_t chunk, int device,
                         int64_t* bytes_moved);

/* mf_kalman_log_likelihood on host arrays (layouts as above; out [B] on the host).  The series are
 * cut into chunks of `chunk_steps` time steps (<= 0: auto); every chunk is reduced on the device to
 * one scan element while the next one is copied, and the elements are joined at the end, so
 * (2D^2+D+mD+m) values per step travel to the device and ONE value per series travels back. */
int mf_host_kalman_log_likelihood(int dtype, const void* mu0, const void* chol_p0, const void* a,
                                  const void* b, const void* chol_q, const void* h,
                                  const void* obs, const void* chol_r, void* out, int64_t B,
                                  int64_t T, int64_t D, int64_t m, int64_t h_batch,
                                  int64_t r_steps, int64_t chunk_steps, int device,
                                  int64_t* bytes_moved);

/* cudaHostRegister / cudaHostUnregister for callers whose framework owns pageable buffers. */
int mf_host_pin(void* ptr, size_t bytes);
int mf_host_unpin(void* ptr);
/* Free the staging slots of a device (device < 0: current). */
int mf_host_release(int device);

#ifdef __cplusplus
}
#endif
#endif /* MARKOVFLOW_B200_H */
