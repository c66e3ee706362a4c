// Large-block path (MF_SMALL_D_MAX < D <= MF_BIG_D_MAX): ONE WARP PER CHAIN, lane i owns row i of
// the current blocks in registers.
//
// Cholesky(+solve) sweep, per step k (block form of cholesky_band / solve_triang_mat, reference
// block_tri_diag.py:436,350):
//     S = D_k - Ls_{k-1} Ls_{k-1}^T          (rows of Ls_{k-1} broadcast from shared memory)
//     r = b_k - Ls_{k-1} x_{k-1}
//   right-looking factorisation of the tall panel [S; A_k] with r carried along, column by column:
//     rinv = rsqrt(S_jj);  L_ij = S_ij rinv;  Ls_ij = A_ij rinv;  x_j = r_j rinv
//     S_ic -= L_ij L_cj,  A_ic -= Ls_ij L_cj  (c > j; column j of L broadcast through shared memory)
//     r_i  -= L_ij x_j
//   which yields Ld_k, Ls_k = A_k Ld_k^{-T} and x_k = Ld_k^{-1} r in one pass of depth D.
// The next step's blocks are prefetched into shared memory with cp.async while the current step is
// factorised (double buffer per warp), so HBM latency is off the sequential path.
#pragma once
#include <cstdint>

#include "pipe.cuh"
#include "ssm_kernels.cuh"

namespace mf {

constexpr int kBigWarps = 2;  // chains (warps) per CTA

template <typename T, int D>
struct BigCfg {
  static constexpr int DD = D * D;
  static constexpr int LDP = D | 1;                        // odd row stride: lane-per-row reads conflict-free
  static constexpr int BLK = D * LDP;                      // one padded block in shared memory
  static constexpr int STAGE = 2 * BLK + D;                // diag | sub | rhs of one step
  static constexpr int PER_WARP = 2 * STAGE + BLK + 2 * D;  // 2 stages + Ls_{k-1} + x_{k-1} + column
  static constexpr size_t SMEM_BYTES = sizeof(T) * (size_t)PER_WARP * kBigWarps;
};

template <typename T>
__device__ __forceinline__ void big_copy_async(T* dst, const T* src, int n, int lane) {
  for (int i = lane; i < n; i += 32) cp_async_elem<(int)sizeof(T)>(dst + i, src + i);
}

// D x D block, contiguous in global memory -> rows of stride LDP in shared memory
template <typename T, int D, int LDP>
__device__ __forceinline__ void big_block_async(T* dst, const T* src, int lane) {
  for (int i = lane; i < D * D; i += 32)
    cp_async_elem<(int)sizeof(T)>(dst + (i / D) * LDP + (i % D), src + i);
}

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <typename T, int D>
__global__ void __launch_bounds__(32 * kBigWarps)
btd_chol_big_kernel(const T* __restrict__ diag, const T* __restrict__ sub,
                    const T* __restrict__ rhs, T* od, T* os, T* ox, T* __restrict__ logdet,
                    int32_t* __restrict__ info, int64_t B, int64_t Tn) {
  using Cfg = BigCfg<T, D>;
  constexpr int DD = Cfg::DD, LDP = Cfg::LDP, BLK = Cfg::BLK;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t chain = (int64_t)blockIdx.x * kBigWarps + warp;
  if (chain >= B) return;
  T* base = reinterpret_cast<T*>(smem_raw) + (size_t)warp * Cfg::PER_WARP;
  T* stage[2] = {base, base + Cfg::STAGE};
  T* lp = base + 2 * Cfg::STAGE;  // Ls_{k-1}, rows of stride LDP
  T* xp = lp + BLK;               // x_{k-1}
  T* col = xp + D;                // column j of L
  const T* dp = diag + chain * Tn * DD;
  const T* sp = sub ? sub + chain * (Tn - 1) * DD : nullptr;
  const T* rp = rhs ? rhs + chain * Tn * D : nullptr;
  T* odp = od + chain * Tn * DD;
  T* osp = os ? os + chain * (Tn - 1) * DD : nullptr;
  T* oxp = ox ? ox + chain * Tn * D : nullptr;
  const bool row = lane < D;

  auto prefetch = [&](int64_t k, T* st) {
    big_block_async<T, D, LDP>(st, dp + k * DD, lane);
    if (sp && k + 1 < Tn) big_block_async<T, D, LDP>(st + BLK, sp + k * DD, lane);
    if (rp) big_copy_async<T>(st + 2 * BLK, rp + k * D, D, lane);
    cp_async_commit();
  };

  T S[D], A[D];  // this lane's rows of the panel
  T r = T(0);    // this lane's entry of the right-hand side
  LogProd<T> det;
  det.init();
  int32_t fail = 0;
  prefetch(0, stage[0]);
  for (int64_t k = 0; k < Tn; ++k) {
    T* st = stage[k & 1];
    if (k + 1 < Tn) {
      prefetch(k + 1, stage[(k + 1) & 1]);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncwarp();
    const bool has_sub = sp && (k + 1 < Tn);
    if (row) {
#pragma unroll
      for (int c = 0; c < D; ++c) {
        S[c] = st[lane * LDP + c];
        A[c] = has_sub ? st[BLK + lane * LDP + c] : T(0);
      }
      r = rp ? st[2 * BLK + lane] : T(0);
    }
    if (k > 0 && sp && row) {
      // Schur update with the previous sub-diagonal factor: own row of Ls_{k-1} from lp as well
      T mine[D];
#pragma unroll
      for (int q = 0; q < D; ++q) mine[q] = lp[lane * LDP + q];
#pragma unroll
      for (int c = 0; c < D; ++c) {
        T acc = T(0);
#pragma unroll
        for (int q = 0; q < D; ++q) acc = Num<T>::fma(mine[q], lp[c * LDP + q], acc);
        S[c] -= acc;
      }
      if (rp) {
        T acc = T(0);
#pragma unroll
        for (int q = 0; q < D; ++q) acc = Num<T>::fma(mine[q], xp[q], acc);
        r -= acc;
      }
    }
    __syncwarp();
    // right-looking factorisation of [S; A] with r carried along
    T xk = T(0);
#pragma unroll
    for (int j = 0; j < D; ++j) {
      const T piv = __shfl_sync(0xffffffffu, S[j], j);
      if (!(piv > T(0)) && fail == 0) fail = (int32_t)(k + 1);
      const T rinv = Num<T>::rsqrt_seq(piv);
      det.mul(piv);
      const T lij = S[j] * rinv;  // L_ij (i >= j); lane j: sqrt(piv)
      const T lsij = A[j] * rinv;
      S[j] = lij;
      A[j] = lsij;
      const T xj = __shfl_sync(0xffffffffu, r, j) * rinv;
      if (lane == j) xk = xj;
      if (lane > j) r = Num<T>::fma(-lij, xj, r);
      if (row) col[lane] = lij;
      __syncwarp();
#pragma unroll
      for (int c = j + 1; c < D; ++c) {
        const T lc = col[c];
        S[c] = Num<T>::fma(-lij, lc, S[c]);
        A[c] = Num<T>::fma(-lsij, lc, A[c]);
      }
      __syncwarp();
    }
    // write back: Ld row (upper triangle zero), Ls row, x; keep Ls / x for the next step
    if (row) {
#pragma unroll
      for (int c = 0; c < D; ++c) odp[k * DD + lane * D + c] = (c <= lane) ? S[c] : T(0);
      if (has_sub) {
#pragma unroll
        for (int c = 0; c < D; ++c) {
          osp[k * DD + lane * D + c] = A[c];
          lp[lane * LDP + c] = A[c];
        }
      }
      if (rp) {
        oxp[k * D + lane] = xk;
        xp[lane] = xk;
      }
    }
    __syncwarp();
  }
  if (lane == 0) {
    if (logdet) logdet[chain] = T(0.5) * det.log_abs();
    if (info) info[chain] = fail;
  }
}

// Triangular solve with a lower block-bidiagonal factor (forward) or its transpose (backward) for
// large blocks: one warp per right-hand-side chain, the step's blocks staged in shared memory
// (double-buffered cp.async), lane i owns entry i of the running vectors.
//   forward : x_k = Ld_k^{-1} (b_k - Ls_{k-1} x_{k-1})
//   backward: x_k = Ld_k^{-T} (b_k - Ls_k^T x_{k+1})
template <typename T, int D>
__global__ void __launch_bounds__(32 * kBigWarps)
btd_solve_big_kernel(const T* __restrict__ ld, const T* __restrict__ ls, const T* __restrict__ rhs,
                     T* out, int64_t n_rhs, int64_t Bm, int64_t Tn, int transpose) {
  using Cfg = BigCfg<T, D>;
  constexpr int DD = Cfg::DD, LDP = Cfg::LDP, BLK = Cfg::BLK;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t chain = (int64_t)blockIdx.x * kBigWarps + warp;
  if (chain >= n_rhs) return;
  const int64_t cm = chain % Bm;
  T* base = reinterpret_cast<T*>(smem_raw) + (size_t)warp * Cfg::PER_WARP;
  T* stage[2] = {base, base + Cfg::STAGE};
  T* xs = base + 2 * Cfg::STAGE;  // running x (D entries)
  const T* lp = ld ? ld + cm * Tn * DD : nullptr;
  const T* sp = ls ? ls + cm * (Tn - 1) * DD : nullptr;
  const T* rp = rhs + chain * Tn * D;
  T* op = out + chain * Tn * D;
  const bool row = lane < D;
  // step index k runs forwards or backwards; sub block coupling step k with the previous one:
  //   forward: Ls_{k-1} (entry k-1), backward: Ls_k (entry k)
  auto prefetch = [&](int64_t k, T* st) {
    if (lp) big_block_async<T, D, LDP>(st, lp + k * DD, lane);
    const int64_t se = transpose ? k : k - 1;
    if (sp && se >= 0 && se < Tn - 1) big_block_async<T, D, LDP>(st + BLK, sp + se * DD, lane);
    big_copy_async<T>(st + 2 * BLK, rp + k * D, D, lane);
    cp_async_commit();
  };
  const int64_t k_first = transpose ? Tn - 1 : 0, dk = transpose ? -1 : 1;
  prefetch(k_first, stage[0]);
  if (row) xs[lane] = T(0);
  for (int64_t it = 0; it < Tn; ++it) {
    const int64_t k = k_first + it * dk;
    T* st = stage[it & 1];
    if (it + 1 < Tn) {
      prefetch(k + dk, stage[(it + 1) & 1]);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncwarp();
    const int64_t se = transpose ? k : k - 1;
    const bool has_sub = sp && se >= 0 && se < Tn - 1 && it > 0;
    T r = row ? st[2 * BLK + lane] : T(0);
    if (has_sub && row) {
      T acc = T(0);
      if (!transpose) {
#pragma unroll
        for (int q = 0; q < D; ++q) acc = Num<T>::fma(st[BLK + lane * LDP + q], xs[q], acc);
      } else {
#pragma unroll
        for (int q = 0; q < D; ++q) acc = Num<T>::fma(st[BLK + q * LDP + lane], xs[q], acc);
      }
      r -= acc;
    }
    __syncwarp();
    T xk = r;
    if (lp) {
      xk = T(0);
      if (!transpose) {
        // column-oriented forward substitution: lane i holds row i of Ld
#pragma unroll
        for (int j = 0; j < D; ++j) {
          const T ljj = st[j * LDP + j];
          const T xj = __shfl_sync(0xffffffffu, r, j) / ljj;
          if (lane == j) xk = xj;
          if (lane > j && row) r = Num<T>::fma(-st[lane * LDP + j], xj, r);
        }
      } else {
        // backward substitution with Ld^T: lane i needs column i of Ld = L[j][i], j > i
#pragma unroll
        for (int j = D - 1; j >= 0; --j) {
          const T ljj = st[j * LDP + j];
          const T xj = __shfl_sync(0xffffffffu, r, j) / ljj;
          if (lane == j) xk = xj;
          if (lane < j) r = Num<T>::fma(-st[j * LDP + lane], xj, r);
        }
      }
    }
    if (row) {
      op[k * D + lane] = xk;
      xs[lane] = xk;
    }
    __syncwarp();
  }
}

}  // namespace mf
