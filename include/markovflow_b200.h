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
 *   - support dtype MF_F32 / MF_F64 and any D >= 1 (thread-per-chain register kernels for
 *     D <= MF_SMALL_D_MAX, a cooperative shared-memory path above that).
 *
 * The mf_host_* variants take HOST pointers and perform the host<->device copies themselves
 * (chunked over the batch and pipelined on internal streams); they are what a host-resident
 * framework (TensorFlow CPU tensors) would bind.
 */
#ifndef MARKOVFLOW_B200_H
#define MARKOVFLOW_B200_H

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

#define MF_SMALL_D_MAX 8

/* Library / build information. */
int mf_version(void);
const char* mf_last_cuda_error(void);
/* Development/tuning knobs (kernel variant selection); knob 0: Cholesky sweep variant
 * (0 auto = TMA bulk-copy ring, 1 direct global-memory streaming, 2 cp.async element-staged ring),
 * knob 1: steps per stage of variant 2. */
int mf_set_tuning(int knob, int value);

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
 *   ld [Bm,T,D,D] (lower triangle read), ls [Bm,T-1,D,D] or NULL, rhs/out [n_rhs,T,D].
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

#ifdef __cplusplus
}
#endif
#endif /* MARKOVFLOW_B200_H */
