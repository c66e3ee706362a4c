// Mid-size state dimensions, 8 < D <= 32: the natural / expectation parameter transforms and the moment recursion
// with ONE WARP per chain (sequential recursions) or per (chain, step) (per-step maps), blocks in shared memory
// (midmat.cuh).  Same algorithms, in the same order, as the one-thread-per-chain kernels of nat_kernels.cuh /
// ssm_kernels.cuh, which hold a block in registers and stop at D = 8.  Reference:
// markovflow/ssm_gaussian_transformations.py:31-593, state_space_model.py:202-275.
#pragma once
#include "kalman_core.cuh"
#include "midmat.cuh"

namespace mf {

// M += alpha * u v^T for vectors held one element per lane
template <typename T>
__device__ __forceinline__ void mid_rank1(T* __restrict__ m, T alpha, T u, T v, int d, int lane) {
  __syncwarp();
  const T au = alpha * u;
  for (int j = 0; j < d; ++j) {
    const T vj = __shfl_sync(0xffffffffu, v, j);
    if (lane < d) m[lane * MID_LD + j] = Num<T>::fma(au, vj, m[lane * MID_LD + j]);
  }
  __syncwarp();
}
// upper triangle <- lower triangle
template <typename T>
__device__ __forceinline__ void mid_mirror_lower(T* m, int d, int lane) {
  __syncwarp();
  if (lane < d)
    for (int j = lane + 1; j < d; ++j) m[lane * MID_LD + j] = m[j * MID_LD + lane];
  __syncwarp();
}

template <typename T>
struct MidSmem {
  T* base;
  __device__ __forceinline__ T* mat(int i) const { return base + i * MID_MAT; }
  __device__ __forceinline__ T* vec(int nmat, int i) const { return base + nmat * MID_MAT + i * 32; }
};
inline size_t mid_smem_bytes(int nmat, int nvec, size_t es) { return ((size_t)nmat * MID_MAT + (size_t)nvec * 32) * es; }

// ---- naturals_to_ssm_params, smoothing (nat_kernels.cuh::nat_to_ssm_kernel): backward sweep, warp per chain -------
template <typename T>
__global__ void __launch_bounds__(32)
mid_nat_to_ssm_kernel(const T* __restrict__ th_lin, const T* __restrict__ th_diag, const T* __restrict__ th_sub,
                      T* __restrict__ out_a, T* __restrict__ out_off, T* __restrict__ out_chol,
                      int32_t* __restrict__ info, int64_t B, int64_t Tn, int d) {
  extern __shared__ __align__(16) unsigned char mid_raw[];
  const MidSmem<T> sm{reinterpret_cast<T*>(mid_raw)};
  constexpr int NM = 6;
  T *Dk = sm.mat(0), *S = sm.mat(1), *A = sm.mat(2), *Th = sm.mat(3), *Qc = sm.mat(4), *W = sm.mat(5);
  T *rinv = sm.vec(NM, 0), *rinv2 = sm.vec(NM, 1);
  const int lane = threadIdx.x;
  const int64_t c = blockIdx.x;
  const int dd = d * d;
  const T* lp = th_lin + c * Tn * d;
  const T* dp = th_diag + c * Tn * dd;
  const T* sp = th_sub + c * (Tn - 1) * dd;
  T* ap = out_a + c * (Tn - 1) * dd;
  T* op = out_off + c * Tn * d;
  T* cp = out_chol + c * Tn * dd;
  int32_t fail = 0;
  T z = T(0);
  for (int64_t k = Tn - 1; k >= 0; --k) {
    mid_load<T>(Dk, dp + k * dd, d, lane);
    T th = lane < d ? lp[k * d + lane] : T(0);
    mid_scale<T>(Dk, T(-2), d, lane);
    if (k + 1 < Tn) {
      mid_load<T>(A, sp + k * dd, d, lane);
      mid_copy<T>(Th, A, d, lane);
      mid_trsm_l<T>(S, rinv, A, d, lane);
      mid_trsm_lt<T>(S, rinv, A, d, lane);  // A_k = D_{k+1}^{-1} theta_sub_k
      mid_store<T>(ap + k * dd, A, d, lane);
      mid_gemm<T, true, false, -1>(Dk, Th, A, d, lane);  // D_k = -2 theta_diag_k - theta_sub_k^T A_k
      th = mid_gemv<T, true, 1>(A, z, th, d, lane);      // z_k = theta_lin_k + A_k^T z_{k+1}
    }
    z = th;
    mid_copy<T>(S, Dk, d, lane);
    const bool ok = mid_chol<T>(S, rinv, d, lane);
    if (!ok && fail == 0) fail = (int32_t)(k + 1);
    T off = mid_trsv_l<T>(S, rinv, z, d, lane);
    off = mid_trsv_lt<T>(S, rinv, off, d, lane);  // offset_k = D_k^{-1} z_k
    if (lane < d) op[k * d + lane] = off;
    mid_chol_inverse<T>(Qc, S, rinv, W, d, lane);  // Q_k = D_k^{-1}
    mid_chol<T>(Qc, rinv2, d, lane);
    mid_store_lower<T>(cp + k * dd, Qc, d, lane);
  }
  if (info && lane == 0) info[c] = fail;
}

// ---- moment recursion (ssm_kernels.cuh::ssm_marginals_kernel, nat_kernels.cuh::ssm_to_expectations_kernel) ------
// EXPECT = false: (mu_k, Sigma_kk, A_k Sigma_kk);  true: (mu_k, Sigma_kk + mu mu^T, A_k Sigma_kk + mu_{k+1} mu_k^T).
// Any output may be NULL.
template <typename T, bool EXPECT>
__global__ void __launch_bounds__(32)
mid_ssm_moments_kernel(const T* __restrict__ mu0, const T* __restrict__ chol_p0, const T* __restrict__ a,
                       const T* __restrict__ b, const T* __restrict__ chol_q, T* __restrict__ o_vec,
                       T* __restrict__ o_diag, T* __restrict__ o_sub, int64_t B, int64_t Tn, int d) {
  extern __shared__ __align__(16) unsigned char mid_raw[];
  const MidSmem<T> sm{reinterpret_cast<T*>(mid_raw)};
  T *P = sm.mat(0), *A = sm.mat(1), *L = sm.mat(2), *AP = sm.mat(3), *E = sm.mat(4);
  const int lane = threadIdx.x;
  const int64_t c = blockIdx.x;
  const int dd = d * d;
  const T* ap = a + c * (Tn - 1) * dd;
  const T* bp = b + c * (Tn - 1) * d;
  const T* qp = chol_q + c * (Tn - 1) * dd;
  T mu = lane < d ? mu0[c * d + lane] : T(0);
  mid_load<T>(L, chol_p0 + c * dd, d, lane);
  mid_gemm<T, false, true, 0>(P, L, L, d, lane);
  for (int64_t k = 0; k < Tn; ++k) {
    if (o_vec && lane < d) o_vec[(c * Tn + k) * d + lane] = mu;
    if (o_diag) {
      if (EXPECT) {
        mid_copy<T>(E, P, d, lane);
        mid_rank1<T>(E, T(1), mu, mu, d, lane);
        mid_store<T>(o_diag + (c * Tn + k) * dd, E, d, lane);
      } else {
        mid_store<T>(o_diag + (c * Tn + k) * dd, P, d, lane);
      }
    }
    if (k + 1 == Tn) break;
    mid_load<T>(A, ap + k * dd, d, lane);
    mid_load<T>(L, qp + k * dd, d, lane);
    const T bk = lane < d ? bp[k * d + lane] : T(0);
    const T nmu = mid_gemv<T, false, 1>(A, mu, bk, d, lane);
    mid_gemm<T, false, false, 0>(AP, A, P, d, lane);
    if (o_sub) {
      if (EXPECT) {
        mid_copy<T>(E, AP, d, lane);
        mid_rank1<T>(E, T(1), nmu, mu, d, lane);
        mid_store<T>(o_sub + (c * (Tn - 1) + k) * dd, E, d, lane);
      } else {
        mid_store<T>(o_sub + (c * (Tn - 1) + k) * dd, AP, d, lane);
      }
    }
    mid_gemm<T, false, true, 0>(P, L, L, d, lane);   // Q_{k+1}
    mid_gemm<T, false, true, 1>(P, AP, A, d, lane);  // + A P A^T
    mid_mirror_lower<T>(P, d, lane);
    mu = nmu;
  }
}

// ---- expectations_to_ssm_params (nat_kernels.cuh::expectations_to_ssm_kernel): warp per (chain, step) ---------------
template <typename T>
__global__ void __launch_bounds__(32)
mid_expectations_to_ssm_kernel(const T* __restrict__ eta_lin, const T* __restrict__ eta_diag,
                               const T* __restrict__ eta_sub, T* __restrict__ out_a, T* __restrict__ out_off,
                               T* __restrict__ out_chol, int32_t* __restrict__ info, int64_t B, int64_t Tn, int d) {
  extern __shared__ __align__(16) unsigned char mid_raw[];
  const MidSmem<T> sm{reinterpret_cast<T*>(mid_raw)};
  constexpr int NM = 6;
  T *Sk = sm.mat(0), *Sp = sm.mat(1), *Lp = sm.mat(2), *X = sm.mat(3), *A = sm.mat(4), *W = sm.mat(5);
  T* rinv = sm.vec(NM, 0);
  const int lane = threadIdx.x;
  const int64_t idx = blockIdx.x;
  const int64_t c = idx / Tn, k = idx % Tn;
  const int dd = d * d;
  const T ek = lane < d ? eta_lin[idx * d + lane] : T(0);
  mid_load<T>(Sk, eta_diag + idx * dd, d, lane);
  mid_rank1<T>(Sk, T(-1), ek, ek, d, lane);  // Sigma_k
  bool ok = true;
  if (k == 0) {
    if (lane < d) out_off[idx * d + lane] = ek;
    ok = mid_chol<T>(Sk, rinv, d, lane);
    mid_store_lower<T>(out_chol + idx * dd, Sk, d, lane);
  } else {
    const T ep = lane < d ? eta_lin[(idx - 1) * d + lane] : T(0);
    mid_load<T>(Sp, eta_diag + (idx - 1) * dd, d, lane);
    mid_rank1<T>(Sp, T(-1), ep, ep, d, lane);  // Sigma_{k-1}
    mid_load<T>(W, eta_sub + (c * (Tn - 1) + k - 1) * dd, d, lane);
    mid_transpose<T>(X, W, d, lane);
    mid_rank1<T>(X, T(-1), ep, ek, d, lane);  // Sigma_{k-1,k} = eta_sub^T - eta_{k-1} eta_k^T
    mid_copy<T>(Lp, Sp, d, lane);
    ok = mid_chol<T>(Lp, rinv, d, lane);
    mid_trsm_l<T>(Lp, rinv, X, d, lane);
    mid_trsm_lt<T>(Lp, rinv, X, d, lane);  // Sigma_{k-1}^{-1} Sigma_{k-1,k} = A^T
    mid_transpose<T>(A, X, d, lane);
    mid_store<T>(out_a + (c * (Tn - 1) + k - 1) * dd, A, d, lane);
    const T off = mid_gemv<T, false, -1>(A, ep, ek, d, lane);  // b = eta_k - A eta_{k-1}
    if (lane < d) out_off[idx * d + lane] = off;
    mid_gemm<T, false, false, 0>(W, A, Sp, d, lane);
    mid_gemm<T, false, true, -1>(Sk, W, A, d, lane);  // Sigma_k - A Sigma_{k-1} A^T
    ok = mid_chol<T>(Sk, rinv, d, lane) && ok;
    mid_store_lower<T>(out_chol + idx * dd, Sk, d, lane);
  }
  if (info && !ok && lane == 0) atomicMax(info + c, (int32_t)(k + 1));
}

// ---- ssm_to_naturals (+ no smoothing) (nat_kernels.cuh::ssm_to_naturals_kernel): warp per (chain, step) -------------
template <typename T>
__global__ void __launch_bounds__(32)
mid_ssm_to_naturals_kernel(const T* __restrict__ mu0, const T* __restrict__ chol_p0, const T* __restrict__ a,
                           const T* __restrict__ b, const T* __restrict__ chol_q, T* __restrict__ th_lin,
                           T* __restrict__ th_diag, T* __restrict__ th_sub, int64_t B, int64_t Tn, int d,
                           int smoothing) {
  extern __shared__ __align__(16) unsigned char mid_raw[];
  const MidSmem<T> sm{reinterpret_cast<T*>(mid_raw)};
  constexpr int NM = 5;
  T *L = sm.mat(0), *Qi = sm.mat(1), *A = sm.mat(2), *X = sm.mat(3), *W = sm.mat(4);
  T* rinv = sm.vec(NM, 0);
  const int lane = threadIdx.x;
  const int64_t idx = blockIdx.x;
  const int64_t c = idx / Tn, k = idx % Tn;
  const int64_t tr = c * (Tn - 1);
  const int dd = d * d;
  mid_load<T>(L, k == 0 ? chol_p0 + c * dd : chol_q + (tr + k - 1) * dd, d, lane);
  T lin = lane < d ? (k == 0 ? mu0[c * d + lane] : b[(tr + k - 1) * d + lane]) : T(0);
  mid_diag_rcp<T>(L, rinv, d, lane);
  mid_chol_inverse<T>(Qi, L, rinv, W, d, lane);
  lin = mid_trsv_l<T>(L, rinv, lin, d, lane);
  lin = mid_trsv_lt<T>(L, rinv, lin, d, lane);  // Q_k^{-1} m_k
  if (k + 1 < Tn) {
    mid_load<T>(L, chol_q + (tr + k) * dd, d, lane);
    mid_load<T>(A, a + (tr + k) * dd, d, lane);
    const T nm = lane < d ? b[(tr + k) * d + lane] : T(0);
    mid_diag_rcp<T>(L, rinv, d, lane);
    mid_copy<T>(X, A, d, lane);
    mid_trsm_l<T>(L, rinv, X, d, lane);
    mid_trsm_lt<T>(L, rinv, X, d, lane);  // Q_{k+1}^{-1} A_k
    mid_store<T>(th_sub + (tr + k) * dd, X, d, lane);
    if (smoothing) {
      mid_gemm<T, true, false, 1>(Qi, A, X, d, lane);
      mid_mirror_lower<T>(Qi, d, lane);
      lin = mid_gemv<T, true, -1>(X, nm, lin, d, lane);  // - A^T Q^{-1} m_{k+1}
    }
  }
  mid_scale<T>(Qi, T(-0.5), d, lane);
  mid_store<T>(th_diag + idx * dd, Qi, d, lane);
  if (lane < d) th_lin[idx * d + lane] = lin;
}

// ---- naturals_to_ssm_params_no_smoothing (nat_kernels.cuh): warp per (chain, step) ----------------------------------
template <typename T>
__global__ void __launch_bounds__(32)
mid_nat_to_ssm_no_smoothing_kernel(const T* __restrict__ th_lin, const T* __restrict__ th_diag,
                                   const T* __restrict__ th_sub, T* __restrict__ out_a, T* __restrict__ out_off,
                                   T* __restrict__ out_chol, int32_t* __restrict__ info, int64_t B, int64_t Tn,
                                   int d) {
  extern __shared__ __align__(16) unsigned char mid_raw[];
  const MidSmem<T> sm{reinterpret_cast<T*>(mid_raw)};
  constexpr int NM = 4;
  T *S = sm.mat(0), *A = sm.mat(1), *Qc = sm.mat(2), *W = sm.mat(3);
  T *rinv = sm.vec(NM, 0), *rinv2 = sm.vec(NM, 1);
  const int lane = threadIdx.x;
  const int64_t idx = blockIdx.x;
  const int64_t c = idx / Tn, k = idx % Tn;
  const int dd = d * d;
  mid_load<T>(S, th_diag + idx * dd, d, lane);
  mid_scale<T>(S, T(-2), d, lane);
  const bool ok = mid_chol<T>(S, rinv, d, lane);
  T off = lane < d ? th_lin[idx * d + lane] : T(0);
  off = mid_trsv_l<T>(S, rinv, off, d, lane);
  off = mid_trsv_lt<T>(S, rinv, off, d, lane);
  if (lane < d) out_off[idx * d + lane] = off;
  if (k > 0) {
    mid_load<T>(A, th_sub + (c * (Tn - 1) + k - 1) * dd, d, lane);
    mid_trsm_l<T>(S, rinv, A, d, lane);
    mid_trsm_lt<T>(S, rinv, A, d, lane);
    mid_store<T>(out_a + (c * (Tn - 1) + k - 1) * dd, A, d, lane);
  }
  mid_chol_inverse<T>(Qc, S, rinv, W, d, lane);
  mid_chol<T>(Qc, rinv2, d, lane);
  mid_store_lower<T>(out_chol + idx * dd, Qc, d, lane);
  if (info && !ok && lane == 0) atomicMax(info + c, (int32_t)(k + 1));
}

// ---- per-block maps (ssm_kernels.cuh::block_cholesky_or_zero_kernel, block_chol_of_inverse_kernel): warp per block ---
template <typename T>
__global__ void __launch_bounds__(32)
mid_block_cholesky_or_zero_kernel(const T* __restrict__ cov, T* __restrict__ out, int32_t* __restrict__ info,
                                  int64_t n, int d) {
  extern __shared__ __align__(16) unsigned char mid_raw[];
  const MidSmem<T> sm{reinterpret_cast<T*>(mid_raw)};
  T* S = sm.mat(0);
  T* rinv = sm.vec(1, 0);
  const int lane = threadIdx.x;
  const int64_t i = blockIdx.x;
  const int dd = d * d;
  bool nz = false;
  for (int idx = lane; idx < dd; idx += 32) nz = nz || (cov[i * dd + idx] != T(0));
  const bool all_zero = !__any_sync(0xffffffffu, nz);
  mid_load<T>(S, cov + i * dd, d, lane);
  bool ok = true;
  if (!all_zero) {
    ok = mid_chol<T>(S, rinv, d, lane);
    mid_store_lower<T>(out + i * dd, S, d, lane);
  } else {
    mid_store<T>(out + i * dd, S, d, lane);
  }
  if (info && !ok && lane == 0) atomicMax(info, (int32_t)(i < 2147483647 ? i + 1 : 2147483647));
}

template <typename T>
__global__ void __launch_bounds__(32)
mid_block_chol_of_inverse_kernel(const T* __restrict__ l, T* __restrict__ out, int64_t n, int d) {
  extern __shared__ __align__(16) unsigned char mid_raw[];
  const MidSmem<T> sm{reinterpret_cast<T*>(mid_raw)};
  T *L = sm.mat(0), *S = sm.mat(1), *W = sm.mat(2);
  T* rinv = sm.vec(3, 0);
  const int lane = threadIdx.x;
  const int64_t i = blockIdx.x;
  const int dd = d * d;
  mid_load<T>(L, l + i * dd, d, lane);
  mid_diag_rcp<T>(L, rinv, d, lane);
  mid_chol_inverse<T>(S, L, rinv, W, d, lane);
  mid_chol<T>(S, rinv, d, lane);
  mid_store_lower<T>(out + i * dd, S, d, lane);
}

// ---- _build_precision (+ H^T R^-1 H) (ssm_kernels.cuh::ssm_build_precision_kernel): warp per (chain, step) ----------
template <typename T>
__global__ void __launch_bounds__(32)
mid_build_precision_kernel(const T* __restrict__ chol_p0, const T* __restrict__ a, const T* __restrict__ chol_q,
                           const T* __restrict__ h, const T* __restrict__ r_inv, T* __restrict__ out_diag,
                           T* __restrict__ out_sub, int64_t B, int64_t Tn, int d, int m, int64_t Bh, int64_t Tr) {
  extern __shared__ __align__(16) unsigned char mid_raw[];
  const MidSmem<T> sm{reinterpret_cast<T*>(mid_raw)};
  constexpr int NM = 5;
  T *L = sm.mat(0), *Qi = sm.mat(1), *A = sm.mat(2), *X = sm.mat(3), *W = sm.mat(4);
  T* rinv = sm.vec(NM, 0);
  const int lane = threadIdx.x;
  const int64_t idx = blockIdx.x;
  const int64_t c = idx / Tn, k = idx % Tn;
  const int dd = d * d;
  mid_load<T>(L, k == 0 ? chol_p0 + c * dd : chol_q + (c * (Tn - 1) + k - 1) * dd, d, lane);
  mid_diag_rcp<T>(L, rinv, d, lane);
  mid_chol_inverse<T>(Qi, L, rinv, W, d, lane);
  if (k + 1 < Tn) {
    mid_load<T>(L, chol_q + (c * (Tn - 1) + k) * dd, d, lane);
    mid_load<T>(A, a + (c * (Tn - 1) + k) * dd, d, lane);
    mid_diag_rcp<T>(L, rinv, d, lane);
    mid_copy<T>(X, A, d, lane);
    mid_trsm_l<T>(L, rinv, X, d, lane);
    mid_trsm_lt<T>(L, rinv, X, d, lane);  // X = Q_k^{-1} A_k
    mid_gemm<T, true, false, 1>(Qi, A, X, d, lane);
    mid_mirror_lower<T>(Qi, d, lane);
    mid_scale<T>(X, T(-1), d, lane);
    mid_store<T>(out_sub + (c * (Tn - 1) + k) * dd, X, d, lane);
  }
  if (h) {
    const T* hp = h + (((Bh == 1) ? 0 : c) * Tn + k) * (int64_t)m * d;
    const T* rp = r_inv + ((Tr == 1) ? 0 : k) * (int64_t)m * m;
    __syncwarp();
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < m; ++j) {
        const T rij = rp[i * m + j];
        if (lane < d) {
          const T hr = hp[i * d + lane] * rij;
          for (int q = 0; q < d; ++q) Qi[lane * MID_LD + q] = Num<T>::fma(hr, hp[j * d + q], Qi[lane * MID_LD + q]);
        }
      }
    __syncwarp();
  }
  mid_store<T>(out_diag + idx * dd, Qi, d, lane);
}

// ---- sparse inverse subset (btd_direct.cuh::btd_inverse_subset_direct_kernel): backward, warp per chain -----------
template <typename T>
__global__ void __launch_bounds__(32)
mid_inverse_subset_kernel(const T* __restrict__ ld, const T* __restrict__ ls, T* od, T* os, int64_t B, int64_t Tn,
                          int d) {
  extern __shared__ __align__(16) unsigned char mid_raw[];
  const MidSmem<T> sm{reinterpret_cast<T*>(mid_raw)};
  constexpr int NM = 6;
  T *L = sm.mat(0), *J = sm.mat(1), *sig = sm.mat(2), *loc = sm.mat(3), *ssub = sm.mat(4), *W = sm.mat(5);
  T* rinv = sm.vec(NM, 0);
  const int lane = threadIdx.x;
  const int64_t c = blockIdx.x;
  const int dd = d * d;
  const T* lp = ld + c * Tn * dd;
  const T* sp = ls ? ls + c * (Tn - 1) * dd : nullptr;
  T* odp = od + c * Tn * dd;
  T* osp = os ? os + c * (Tn - 1) * dd : nullptr;
  for (int64_t k = Tn - 1; k >= 0; --k) {
    mid_load<T>(L, lp + k * dd, d, lane);
    mid_diag_rcp<T>(L, rinv, d, lane);
    mid_chol_inverse<T>(loc, L, rinv, W, d, lane);
    if (sp && k + 1 < Tn) {
      mid_load<T>(J, sp + k * dd, d, lane);
      mid_trsm_right_l<T>(J, L, rinv, d, lane);          // J = Ls Ld^{-1}
      mid_gemm<T, false, false, 0>(ssub, sig, J, d, lane);  // Sigma_{k+1,k+1} J
      mid_scale<T>(ssub, T(-1), d, lane);
      if (osp) mid_store<T>(osp + k * dd, ssub, d, lane);
      mid_gemm<T, true, false, -1>(loc, J, ssub, d, lane);  // loc -= J^T ssub
      mid_mirror_lower<T>(loc, d, lane);
    }
    mid_store<T>(odp + k * dd, loc, d, lane);
    mid_copy<T>(sig, loc, d, lane);
  }
}

// ---- U D U^T (btd_direct.cuh::btd_udu_direct_kernel): backward, warp per chain --------------------------------------
template <typename T>
__global__ void __launch_bounds__(32)
mid_udu_kernel(const T* __restrict__ diag, const T* __restrict__ sub, T* ou, T* ocd, int32_t* __restrict__ info,
               int64_t B, int64_t Tn, int d) {
  extern __shared__ __align__(16) unsigned char mid_raw[];
  const MidSmem<T> sm{reinterpret_cast<T*>(mid_raw)};
  constexpr int NM = 3;
  T *C = sm.mat(0), *K = sm.mat(1), *X = sm.mat(2);
  T* rinv = sm.vec(NM, 0);
  const int lane = threadIdx.x;
  const int64_t c = blockIdx.x;
  const int dd = d * d;
  const T* dp = diag + c * Tn * dd;
  const T* sp = sub + c * (Tn - 1) * dd;
  T* oup = ou + c * (Tn - 1) * dd;
  T* ocp = ocd + c * Tn * dd;
  int32_t fail = 0;
  mid_load<T>(C, dp + (Tn - 1) * dd, d, lane);
  bool ok = mid_chol<T>(C, rinv, d, lane);
  if (!ok) fail = (int32_t)Tn;
  mid_store_lower<T>(ocp + (Tn - 1) * dd, C, d, lane);
  for (int64_t k = Tn - 2; k >= 0; --k) {
    mid_load<T>(K, sp + k * dd, d, lane);
    mid_copy<T>(X, K, d, lane);
    mid_trsm_l<T>(C, rinv, X, d, lane);
    mid_trsm_lt<T>(C, rinv, X, d, lane);  // X = D_{k+1}^{-1} K_{k+1,k}
    mid_store<T>(oup + k * dd, X, d, lane);
    mid_load<T>(C, dp + k * dd, d, lane);
    mid_gemm<T, true, false, -1>(C, K, X, d, lane);
    ok = mid_chol<T>(C, rinv, d, lane);
    if (!ok && fail == 0) fail = (int32_t)(k + 1);
    mid_store_lower<T>(ocp + k * dd, C, d, lane);
  }
  if (info && lane == 0) info[c] = fail;
}

// ---- affine recurrence (ssm_kernels.cuh::ssm_affine_scan_kernel): marginal means / sample, warp per trajectory -----
template <typename T>
__global__ void __launch_bounds__(32)
mid_affine_scan_kernel(const T* __restrict__ mu0, const T* __restrict__ chol_p0, const T* __restrict__ a,
                       const T* __restrict__ b, const T* __restrict__ chol_q, const T* __restrict__ eps,
                       T* __restrict__ out, int64_t n, int64_t Bm, int64_t Tn, int d) {
  extern __shared__ __align__(16) unsigned char mid_raw[];
  const MidSmem<T> sm{reinterpret_cast<T*>(mid_raw)};
  T *A = sm.mat(0), *L = sm.mat(1);
  const int lane = threadIdx.x;
  const int64_t c = blockIdx.x, cm = c % Bm;
  const int dd = d * d;
  const T* ep = eps ? eps + c * Tn * d : nullptr;
  T x = lane < d ? mu0[cm * d + lane] : T(0);
  if (ep) {
    mid_load<T>(L, chol_p0 + cm * dd, d, lane);
    const T e = lane < d ? ep[lane] : T(0);
    x = mid_gemv<T, false, 1>(L, e, x, d, lane);
  }
  if (lane < d) out[c * Tn * d + lane] = x;
  for (int64_t k = 1; k < Tn; ++k) {
    mid_load<T>(A, a + (cm * (Tn - 1) + k - 1) * dd, d, lane);
    const T bk = lane < d ? b[(cm * (Tn - 1) + k - 1) * d + lane] : T(0);
    x = mid_gemv<T, false, 1>(A, x, bk, d, lane);
    if (ep) {
      mid_load<T>(L, chol_q + (cm * (Tn - 1) + k - 1) * dd, d, lane);
      const T e = lane < d ? ep[k * d + lane] : T(0);
      x = mid_gemv<T, false, 1>(L, e, x, d, lane);
    }
    if (lane < d) out[(c * Tn + k) * d + lane] = x;
  }
}

// ---- Kalman log-likelihood, classical filter with scalar absorption (kalman_kernels.cuh::kalman_walk + FilterSink):
//      warp per series.  Arguments as KalmanArgs; first step is the prior.
template <typename T>
__global__ void __launch_bounds__(32)
mid_kalman_loglik_kernel(const T* __restrict__ mu0, const T* __restrict__ chol_p0, const T* __restrict__ a,
                         const T* __restrict__ b, const T* __restrict__ chol_q, const T* __restrict__ h,
                         const T* __restrict__ obs, const T* __restrict__ chol_r, T* __restrict__ out, int64_t B,
                         int64_t Tn, int d, int m, int64_t Bh, int64_t Tr) {
  extern __shared__ __align__(16) unsigned char mid_raw[];
  const MidSmem<T> sm{reinterpret_cast<T*>(mid_raw)};
  T *P = sm.mat(0), *F = sm.mat(1), *L = sm.mat(2), *FP = sm.mat(3);
  const int lane = threadIdx.x;
  const int64_t c = blockIdx.x;
  const int dd = d * d;
  const int64_t nt = Tn - 1;
  const T* hp = h + (Bh == 1 ? 0 : c) * Tn * (int64_t)m * d;
  const T* yp = obs + c * Tn * (int64_t)m;
  Whitener<T> wh;
  LogProd<T> wdet, det;
  wdet.init();
  det.init();
  T quad = T(0);
  int64_t nobs = 0;
  if (Tr == 1) wh.set(chol_r, m);
  T x = lane < d ? mu0[c * d + lane] : T(0);
  mid_load<T>(L, chol_p0 + c * dd, d, lane);
  mid_gemm<T, false, true, 0>(P, L, L, d, lane);
  for (int64_t k = 0; k < Tn; ++k) {
    if (k > 0) {
      mid_load<T>(F, a + (c * nt + k - 1) * dd, d, lane);
      mid_load<T>(L, chol_q + (c * nt + k - 1) * dd, d, lane);
      const T u = lane < d ? b[(c * nt + k - 1) * d + lane] : T(0);
      x = mid_gemv<T, false, 1>(F, x, u, d, lane);
      mid_gemm<T, false, false, 0>(FP, F, P, d, lane);
      mid_gemm<T, false, true, 0>(P, L, L, d, lane);
      mid_gemm<T, false, true, 1>(P, FP, F, d, lane);
      mid_mirror_lower<T>(P, d, lane);
    }
    if (Tr != 1) wh.set(chol_r + k * (int64_t)m * m, m);
    const T* hk = hp + k * (int64_t)m * d;
    const T* yk = yp + k * (int64_t)m;
    for (int i = 0; i < m; ++i) {
      const T wii = wh.W[i * kMaxObsDim + i];
      if (m == 1 && wii == T(0)) continue;  // an infinite noise scale marks a step without observation
      T hv = T(0), y = T(0);
      for (int j = 0; j <= i; ++j) {
        const T wij = wh.W[i * kMaxObsDim + j];
        y = Num<T>::fma(wij, yk[j], y);
        if (lane < d) hv = Num<T>::fma(wij, hk[j * d + lane], hv);
      }
      wdet.mul(wii);
      // absorb the whitened scalar observation y = hv . x + N(0, 1)
      const T g = mid_gemv<T, false, 0>(P, hv, T(0), d, lane);
      const T s = mid_dot<T>(hv, g, T(1), d);
      const T v = -mid_dot<T>(hv, x, -y, d);
      const T rs = Num<T>::rcp(s);
      const T vs = v * rs;
      quad = Num<T>::fma(v, vs, quad);
      det.mul_lazy(s);
      x = Num<T>::fma(g, vs, x);
      const T ki = g * rs;
      __syncwarp();
      for (int j = 0; j < d; ++j) {
        const T gj = __shfl_sync(0xffffffffu, g, j);
        if (lane < d && j <= lane) P[lane * MID_LD + j] = Num<T>::fma(-ki, gj, P[lane * MID_LD + j]);
      }
      mid_mirror_lower<T>(P, d, lane);
      ++nobs;
    }
    det.peel();
  }
  if (lane == 0)
    out[c] = T(-0.5) * (quad + det.log_abs()) + wdet.log_abs() - T(0.5 * 1.8378770664093454836) * T(nobs);
}

// ---- dense_mult (btd_direct.cuh::btd_dense_mult_kernel): warp per (rhs chain, block row) --------------------------
template <typename T>
__global__ void __launch_bounds__(32)
mid_dense_mult_kernel(const T* __restrict__ diag, const T* __restrict__ sub, const T* __restrict__ right,
                      T* __restrict__ out, int64_t n_rhs, int64_t Bm, int64_t Tn, int d, int transpose,
                      int symmetric) {
  extern __shared__ __align__(16) unsigned char mid_raw[];
  const MidSmem<T> sm{reinterpret_cast<T*>(mid_raw)};
  T* M = sm.mat(0);
  const int lane = threadIdx.x;
  const int64_t idx = blockIdx.x;
  const int64_t c = idx / Tn, k = idx % Tn;
  const int64_t cm = c % Bm;
  const int dd = d * d;
  const T* xp = right + (c * Tn + k) * d;
  mid_load<T>(M, diag + (cm * Tn + k) * dd, d, lane);
  T x = lane < d ? xp[lane] : T(0);
  T y;
  if (symmetric) {
    mid_mirror_lower<T>(M, d, lane);
    y = mid_gemv<T, false, 0>(M, x, T(0), d, lane);
  } else {
    __syncwarp();
    if (lane < d)
      for (int j = lane + 1; j < d; ++j) M[lane * MID_LD + j] = T(0);
    y = transpose ? mid_gemv<T, true, 0>(M, x, T(0), d, lane) : mid_gemv<T, false, 0>(M, x, T(0), d, lane);
  }
  if (sub) {
    const T* sp = sub + cm * (Tn - 1) * dd;
    if ((symmetric || !transpose) && k > 0) {  // + A_{k-1} x_{k-1}
      mid_load<T>(M, sp + (k - 1) * dd, d, lane);
      x = lane < d ? xp[lane - d] : T(0);
      y = mid_gemv<T, false, 1>(M, x, y, d, lane);
    }
    if ((symmetric || transpose) && k + 1 < Tn) {  // + A_k^T x_{k+1}
      mid_load<T>(M, sp + k * dd, d, lane);
      x = lane < d ? xp[lane + d] : T(0);
      y = mid_gemv<T, true, 1>(M, x, y, d, lane);
    }
  }
  if (lane < d) out[(c * Tn + k) * d + lane] = y;
}

// ---- log_pdf (ssm_kernels.cuh::ssm_log_pdf_kernel): warp per trajectory, the factors summed in time order ---------
template <typename T>
__global__ void __launch_bounds__(32)
mid_log_pdf_kernel(const T* __restrict__ mu0, const T* __restrict__ chol_p0, const T* __restrict__ a,
                   const T* __restrict__ b, const T* __restrict__ chol_q, const T* __restrict__ states,
                   T* __restrict__ out, int64_t n, int64_t Bm, int64_t Tn, int d) {
  extern __shared__ __align__(16) unsigned char mid_raw[];
  const MidSmem<T> sm{reinterpret_cast<T*>(mid_raw)};
  T *L = sm.mat(0), *A = sm.mat(1);
  T* rinv = sm.vec(2, 0);
  const int lane = threadIdx.x;
  const int64_t c = blockIdx.x, cm = c % Bm;
  const int dd = d * d;
  const T* xp = states + c * Tn * d;
  T acc = T(0);
  T xprev = T(0);
  for (int64_t k = 0; k < Tn; ++k) {
    const T xk = lane < d ? xp[k * d + lane] : T(0);
    T r;
    if (k == 0) {
      r = xk - (lane < d ? mu0[cm * d + lane] : T(0));
      mid_load<T>(L, chol_p0 + cm * dd, d, lane);
    } else {
      mid_load<T>(A, a + (cm * (Tn - 1) + k - 1) * dd, d, lane);
      r = xk - (lane < d ? b[(cm * (Tn - 1) + k - 1) * d + lane] : T(0));
      r = mid_gemv<T, false, -1>(A, xprev, r, d, lane);
      mid_load<T>(L, chol_q + (cm * (Tn - 1) + k - 1) * dd, d, lane);
    }
    mid_diag_rcp<T>(L, rinv, d, lane);
    r = mid_trsv_l<T>(L, rinv, r, d, lane);
    const T q = mid_dot<T>(r, r, T(0), d);
    T dprod = T(1);
    for (int i = 0; i < d; ++i) dprod *= L[i * MID_LD + i];
    acc += T(-0.5) * q - Num<T>::log(Num<T>::abs(dprod)) - T(0.5 * 1.8378770664093454836) * T(d);
    xprev = xk;
  }
  if (lane == 0) out[c] = acc;
}

// ---- KL(q || p), chain-rule form (ssm_kernels.cuh::ssm_kl_kernel): warp per chain --------------------------------
// sum of squares of the d x d entries of m, accumulated row by row (lane = row), then over the rows in order
template <typename T>
__device__ __forceinline__ T mid_sumsq(const T* __restrict__ m, int d, int lane, bool lower_only) {
  __syncwarp();
  T s = T(0);
  if (lane < d)
    for (int j = 0; j < (lower_only ? lane + 1 : d); ++j) s = Num<T>::fma(m[lane * MID_LD + j], m[lane * MID_LD + j], s);
  T tot = T(0);
  for (int q = 0; q < d; ++q) tot += __shfl_sync(0xffffffffu, s, q);
  return tot;
}

// 0.5 [ |Lp^-1 Lq|_F^2 + tr(G P G^T) + |Lp^-1 dmean|^2 - d ],  G = Lp^-1 dA ; log-dets go to `ratio`
template <typename T>
__device__ __forceinline__ T mid_kl_term(const T* __restrict__ Lp, const T* __restrict__ Lq, T* __restrict__ dA,
                                         T dmean, const T* __restrict__ P, T* __restrict__ W, T* __restrict__ GP,
                                         T* __restrict__ rinv, LogProd<T>& ratio, int d, int lane) {
  mid_diag_rcp<T>(Lp, rinv, d, lane);
  __syncwarp();
  if (lane < d)
    for (int j = 0; j < d; ++j) W[lane * MID_LD + j] = j <= lane ? Lq[lane * MID_LD + j] : T(0);
  mid_trsm_l<T>(Lp, rinv, W, d, lane);
  T s = mid_sumsq<T>(W, d, lane, false);
  const T e = mid_trsv_l<T>(Lp, rinv, dmean, d, lane);
  s = mid_dot<T>(e, e, s, d);
  if (dA) {
    mid_trsm_l<T>(Lp, rinv, dA, d, lane);  // G
    mid_gemm<T, false, false, 0>(GP, dA, P, d, lane);
    __syncwarp();
    T t = T(0);
    if (lane < d)
      for (int j = 0; j < d; ++j) t = Num<T>::fma(GP[lane * MID_LD + j], dA[lane * MID_LD + j], t);
    for (int q = 0; q < d; ++q) s += __shfl_sync(0xffffffffu, t, q);
  }
  for (int i = 0; i < d; ++i) ratio.mul(Lp[i * MID_LD + i] * Num<T>::rcp(Lq[i * MID_LD + i]));
  return T(0.5) * (s - T(d));
}

template <typename T>
__global__ void __launch_bounds__(32)
mid_kl_kernel(const T* __restrict__ q_mu0, const T* __restrict__ q_chol_p0, const T* __restrict__ q_a,
              const T* __restrict__ q_b, const T* __restrict__ q_chol_q, const T* __restrict__ p_mu0,
              const T* __restrict__ p_chol_p0, const T* __restrict__ p_a, const T* __restrict__ p_b,
              const T* __restrict__ p_chol_q, T* __restrict__ out, int64_t B, int64_t Tn, int d) {
  extern __shared__ __align__(16) unsigned char mid_raw[];
  const MidSmem<T> sm{reinterpret_cast<T*>(mid_raw)};
  constexpr int NM = 8;
  T *P = sm.mat(0), *Lq = sm.mat(1), *Lp = sm.mat(2), *Aq = sm.mat(3), *dA = sm.mat(4), *W = sm.mat(5),
    *GP = sm.mat(6), *AP = sm.mat(7);
  T* rinv = sm.vec(NM, 0);
  const int lane = threadIdx.x;
  const int64_t c = blockIdx.x;
  const int dd = d * d;
  LogProd<T> ratio;
  ratio.init();
  T mu = lane < d ? q_mu0[c * d + lane] : T(0);
  mid_load<T>(Lq, q_chol_p0 + c * dd, d, lane);
  mid_load<T>(Lp, p_chol_p0 + c * dd, d, lane);
  T dm = mu - (lane < d ? p_mu0[c * d + lane] : T(0));
  T kl = mid_kl_term<T>(Lp, Lq, nullptr, dm, nullptr, W, GP, rinv, ratio, d, lane);
  mid_gemm<T, false, true, 0>(P, Lq, Lq, d, lane);
  const int64_t off = c * (Tn - 1);
  for (int64_t k = 0; k + 1 < Tn; ++k) {
    mid_load<T>(Aq, q_a + (off + k) * dd, d, lane);
    mid_load<T>(dA, p_a + (off + k) * dd, d, lane);
    mid_load<T>(Lq, q_chol_q + (off + k) * dd, d, lane);
    mid_load<T>(Lp, p_chol_q + (off + k) * dd, d, lane);
    const T bq = lane < d ? q_b[(off + k) * d + lane] : T(0);
    const T bp = lane < d ? p_b[(off + k) * d + lane] : T(0);
    __syncwarp();
    if (lane < d)
      for (int j = 0; j < d; ++j) dA[lane * MID_LD + j] = Aq[lane * MID_LD + j] - dA[lane * MID_LD + j];
    dm = mid_gemv<T, false, 1>(dA, mu, bq - bp, d, lane);  // dA mu + db
    kl += mid_kl_term<T>(Lp, Lq, dA, dm, P, W, GP, rinv, ratio, d, lane);
    // advance q's marginal
    mid_gemm<T, false, false, 0>(AP, Aq, P, d, lane);
    mid_gemm<T, false, true, 0>(P, Lq, Lq, d, lane);
    mid_gemm<T, false, true, 1>(P, AP, Aq, d, lane);
    mid_mirror_lower<T>(P, d, lane);
    mu = mid_gemv<T, false, 1>(Aq, mu, bq, d, lane);
  }
  if (lane == 0) out[c] = kl + ratio.log_abs();
}

// ---- conditionals (capi_cond.cu kernels, conditionals.py:29-205,380-485) for 8 < D <= 32 ---------------------------
// pairwise marginals: pure assembly, one block per (chain, pair)
template <typename T>
__global__ void __launch_bounds__(128)
mid_pairwise_marginals_kernel(const T* __restrict__ mean, const T* __restrict__ cov, const T* __restrict__ sub,
                              const T* __restrict__ init_mean, const T* __restrict__ init_cov, int64_t init_batch,
                              T* __restrict__ o_mean, T* __restrict__ o_cov, int64_t B, int64_t Tn, int d) {
  const int64_t idx = blockIdx.x;
  const int64_t c = idx / (Tn + 1), k = idx % (Tn + 1);
  const int64_t ci = init_batch == 1 ? 0 : c;
  const int dd = d * d, d2 = 2 * d;
  const T* m0 = k == 0 ? init_mean + ci * d : mean + (c * Tn + k - 1) * d;
  const T* S0 = k == 0 ? init_cov + ci * dd : cov + (c * Tn + k - 1) * dd;
  const T* m1 = k == Tn ? init_mean + ci * d : mean + (c * Tn + k) * d;
  const T* S1 = k == Tn ? init_cov + ci * dd : cov + (c * Tn + k) * dd;
  const T* C = (k >= 1 && k < Tn) ? sub + (c * (Tn - 1) + k - 1) * dd : nullptr;
  T* om = o_mean + idx * d2;
  T* oc = o_cov + idx * (int64_t)d2 * d2;
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    om[i] = m0[i];
    om[d + i] = m1[i];
  }
  for (int i = threadIdx.x; i < dd; i += blockDim.x) {
    const int r = i / d, q = i % d;
    oc[r * d2 + q] = S0[i];
    oc[r * d2 + d + q] = C ? C[q * d + r] : T(0);
    oc[(d + r) * d2 + q] = C ? C[i] : T(0);
    oc[(d + r) * d2 + d + q] = S1[i];
  }
}

// conditional statistics: warp per point
template <typename T>
__global__ void __launch_bounds__(32)
mid_conditional_statistics_kernel(const T* __restrict__ a_mt, const T* __restrict__ q_mt, const T* __restrict__ a_tp,
                                  const T* __restrict__ q_tp, T* __restrict__ o_p, T* __restrict__ o_t,
                                  int32_t* __restrict__ info, int return_precision, int64_t N, int d) {
  extern __shared__ __align__(16) unsigned char mid_raw[];
  const MidSmem<T> sm{reinterpret_cast<T*>(mid_raw)};
  constexpr int NM = 9;
  T *Am = sm.mat(0), *Qm = sm.mat(1), *Ap = sm.mat(2), *Qp = sm.mat(3), *AQ = sm.mat(4), *L = sm.mat(5),
    *V = sm.mat(6), *E = sm.mat(7), *W = sm.mat(8);
  T *rinv = sm.vec(NM, 0), *r2 = sm.vec(NM, 1);
  const int lane = threadIdx.x;
  const int64_t n = blockIdx.x;
  const int dd = d * d, d2 = 2 * d;
  mid_load<T>(Am, a_mt + n * dd, d, lane);
  mid_load<T>(Qm, q_mt + n * dd, d, lane);
  mid_load<T>(Ap, a_tp + n * dd, d, lane);
  mid_load<T>(Qp, q_tp + n * dd, d, lane);
  mid_gemm<T, false, false, 0>(AQ, Ap, Qm, d, lane);  // A_tp Q_mt
  mid_gemm<T, false, true, 0>(L, AQ, Ap, d, lane);    // A_tp Q_mt A_tp^T
  mid_axpy<T, 1>(L, Qp, d, lane);
  bool ok = mid_chol<T>(L, rinv, d, lane);
  mid_copy<T>(V, AQ, d, lane);
  mid_trsm_l<T>(L, rinv, V, d, lane);  // V = L^-1 A_tp Q_mt
  mid_copy<T>(W, V, d, lane);
  mid_trsm_lt<T>(L, rinv, W, d, lane);  // L^-T V = E^T
  mid_transpose<T>(E, W, d, lane);
  mid_gemm<T, false, false, 0>(W, E, Ap, d, lane);   // E A_tp
  mid_gemm<T, false, false, 0>(AQ, W, Am, d, lane);  // E A_tp A_mt
  mid_axpy<T, 0>(AQ, Am, d, lane);                   // D = A_mt - E A_tp A_mt
  mid_store_pitched<T>(o_p + n * (int64_t)d * d2, AQ, d, d2, lane);
  mid_store_pitched<T>(o_p + n * (int64_t)d * d2 + d, E, d, d2, lane);
  if (return_precision) {
    ok = mid_chol<T>(Qm, rinv, d, lane) && ok;
    ok = mid_chol<T>(Qp, r2, d, lane) && ok;
    mid_chol_inverse<T>(L, Qm, rinv, W, d, lane);  // Q_mt^-1
    mid_trsm_l<T>(Qp, r2, Ap, d, lane);            // L_tp^-1 A_tp
    mid_gemm<T, true, false, 1>(L, Ap, Ap, d, lane);
    mid_store<T>(o_t + n * dd, L, d, lane);
  } else {
    mid_gemm<T, true, false, -1>(Qm, V, V, d, lane);  // T = Q_mt - V^T V
    mid_store<T>(o_t + n * dd, Qm, d, lane);
  }
  if (info && lane == 0) info[n] = ok ? 0 : 1;
}

// conditional prediction: warp per (chain, point);  mean = P m[idx],  cov = T (+ P S[idx] P^T), P = [P0 | P1]
template <typename T>
__global__ void __launch_bounds__(32)
mid_conditional_predict_kernel(const T* __restrict__ proj, const T* __restrict__ tcov,
                               const T* __restrict__ pair_means, const T* __restrict__ pair_covs,
                               const int64_t* __restrict__ indices, T* __restrict__ o_mean, T* __restrict__ o_cov,
                               int64_t B, int64_t N, int64_t M, int d) {
  extern __shared__ __align__(16) unsigned char mid_raw[];
  const MidSmem<T> sm{reinterpret_cast<T*>(mid_raw)};
  T *P0 = sm.mat(0), *P1 = sm.mat(1), *S = sm.mat(2), *W = sm.mat(3), *Cv = sm.mat(4);
  const int lane = threadIdx.x;
  const int64_t idx = blockIdx.x;
  const int64_t c = idx / N;
  int64_t j = indices ? indices[idx] : idx % N;
  if (j < 0) j = 0;
  if (j > M - 1) j = M - 1;
  const int dd = d * d, d2 = 2 * d;
  const T* P = proj + idx * (int64_t)d * d2;
  const T* m = pair_means + (c * M + j) * d2;
  mid_load_pitched<T>(P0, P, d, d2, lane);
  mid_load_pitched<T>(P1, P + d, d, d2, lane);
  const T m0 = lane < d ? m[lane] : T(0), m1 = lane < d ? m[d + lane] : T(0);
  T mean = mid_gemv<T, false, 0>(P0, m0, T(0), d, lane);
  mean = mid_gemv<T, false, 1>(P1, m1, mean, d, lane);
  if (lane < d) o_mean[idx * d + lane] = mean;
  mid_load<T>(Cv, tcov + idx * dd, d, lane);
  if (pair_covs) {
    const T* Sg = pair_covs + (c * M + j) * (int64_t)d2 * d2;
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b) {
        mid_load_pitched<T>(S, Sg + (int64_t)a * d * d2 + b * d, d, d2, lane);
        mid_gemm<T, false, false, 0>(W, a == 0 ? P0 : P1, S, d, lane);
        mid_gemm<T, false, true, 1>(Cv, W, b == 0 ? P0 : P1, d, lane);
      }
  }
  mid_store<T>(o_cov + idx * dd, Cv, d, lane);
}

}  // namespace mf
