// Internal (C++) interface to the TMA-sweep implementations of the sequential StateSpaceModel /
// natural-parameter recurrences (capi_ssm_sweep.cu).  Every function returns MF_ERR_UNSUPPORTED
// when no ring geometry fits the (dtype, D) at hand; the caller then uses the direct-load kernel.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace mf {

constexpr int kSsmSweepMaxD = 4;

int ssm_sweep_moments(int dtype, int64_t D, int expectations, const void* mu0, const void* chol_p0,
                      const void* a, const void* b, const void* chol_q, void* o_vec, void* o_diag,
                      void* o_sub, int64_t B, int64_t T, cudaStream_t s);
int ssm_sweep_affine(int dtype, int64_t D, const void* mu0, const void* chol_p0, const void* a,
                     const void* b, const void* chol_q, const void* eps, void* out, int64_t n,
                     int64_t Bm, int64_t T, cudaStream_t s, int use_rng = 0, unsigned long long seed = 0);
int ssm_sweep_kl(int dtype, int64_t D, const void* q_mu0, const void* q_chol_p0, const void* q_a,
                 const void* q_b, const void* q_chol_q, const void* p_mu0, const void* p_chol_p0,
                 const void* p_a, const void* p_b, const void* p_chol_q, void* out, int64_t B,
                 int64_t T, cudaStream_t s);
int nat_sweep_to_ssm(int dtype, int64_t D, const void* th_lin, const void* th_diag,
                     const void* th_sub, void* out_a, void* out_off, void* out_chol, int32_t* info,
                     int64_t B, int64_t T, cudaStream_t s);

// block-tridiagonal recurrences (capi_btd_sweep.cu)
int btd_sweep_solve(int dtype, int64_t D, const void* ld, const void* ls, const void* rhs, void* out,
                    int64_t n, int64_t Bm, int64_t T, int transpose, cudaStream_t s);
int btd_sweep_inverse_subset(int dtype, int64_t D, const void* ld, const void* ls, void* od, void* os,
                             int64_t B, int64_t T, cudaStream_t s);
int btd_sweep_cholesky_pit(int dtype, int64_t D, const void* diag, const void* sub, const void* rhs,
                           void* od, void* os, void* ox, int32_t* info, int64_t B, int64_t T,
                           cudaStream_t s);
int btd_sweep_udu(int dtype, int64_t D, const void* diag, const void* sub, void* ou, void* ocd,
                  int32_t* info, int64_t B, int64_t T, cudaStream_t s);

}  // namespace mf
