// Large-block path for 9 <= D <= 17 (BASELINE config 4: D = 17): HALF A WARP PER CHAIN, parallel in time.
//
// A step of the block Cholesky(+solve) sweep (block_tri_diag.py:436,350) is the right-looking elimination of
// D pivots on the 2D x 2D matrix [[S_k, .], [A_k, D_{k+1}]]: it yields Ld_k, Ls_k = A_k Ld_k^-T, x_k and, in
// the trailing block, S_{k+1} = D_{k+1} - Ls_k Ls_k^T.  Layout: 16 lanes per chain, lane h owns ROW h of every
// block in registers; for D = 17 row 16 is a BORDER row stored column-distributed (lane c owns entry [16][c],
// the corner [16][16] is replicated), so both halves of a warp work on two chains at once with every lane busy
// (the one-warp-per-chain kernel of btd_big.cuh leaves 15 of 32 lanes idle at D = 17).  Columns are broadcast
// through shared memory (double-buffered: one __syncwarp per pivot), the next steps' blocks are prefetched
// with cp.async, results are staged in shared memory and written back coalesced.
//
// With 256 chains a sequential sweep is bound by the latency of one chain's step (T x ~9000 cycles, 0.08 of
// the HBM roofline).  The map (S_k, r_k) -> (S_{k+1}, r_{k+1}) is linear-fractional,
//     S_out = P - Q (S_in + R)^-1 Q^T,      r_out = p - Q (S_in + R)^-1 (r_in + r),
// so every chain is cut into P segments that are processed at once (exact, as btd_pit.cuh does for D <= 4):
//   pass 1  big2_element_kernel : segment 0 is factorised for good; segments 1..P-2 reduce their steps to
//                                 (P, Q, R, p, r) -- a step with three row blocks [S; A; Q^T] instead of two
//   pass 2  big2_fold_kernel    : per chain, apply the elements in order (each application IS one ordinary
//                                 step with S := S_in + R, A := Q, D_next := P) -> seed (S, r) of every segment
//   pass 3  big2_factor_kernel  : segments 1..P-1 run the ordinary sweep from their seeds.
// Elements and seeds live in a stream-ordered workspace of the library (in-place factorisation is allowed).
#pragma once
#include <cstdint>

#include "pipe.cuh"
#include "ssm_kernels.cuh"

namespace mf {

template <typename T, int D>
struct Big2 {
  static_assert(D >= 2 && D <= 17, "half-warp layout: at most 16 rows + one border row");
  static constexpr bool BORDER = D == 17;
  static constexpr int DM = BORDER ? 16 : D;  // rows owned one per lane
  static constexpr int DD = D * D;
  static constexpr int STAGE = (2 * DD + D + 1) / 2 * 2;  // diag | sub | rhs of one step, padded to 16 bytes
  static constexpr int NST = 2;                 // stages per chain: the next step is prefetched while one is
                                                // factorised; a consumed stage doubles as the output staging
  static constexpr int COLS = 3 * 2 * 16;       // 3 broadcast columns, double-buffered
  static constexpr int PER_CHAIN = NST * STAGE + COLS;
  static constexpr int ELEM = 3 * DD + 2 * D;   // P | Q | R | p | r
  static constexpr int SEED = DD + D;           // S | r
};

// two adjacent values of a broadcast column in one shared-memory load (LDS.128 for double)
template <typename T> struct Pair2;
template <> struct Pair2<double> { using type = double2; };
template <> struct Pair2<float> { using type = float2; };
template <typename T>
__device__ __forceinline__ typename Pair2<T>::type big2_ld2(const T* p) {
  return *reinterpret_cast<const typename Pair2<T>::type*>(p);
}

__device__ __forceinline__ void big2_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void big2_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// n contiguous values global -> shared, 16 lanes of a half-warp, 8-byte cp.async
template <typename T>
__device__ __forceinline__ void big2_copy_async(T* dst, const T* src, int n, int h) {
  for (int i = h; i < n; i += 16) cp_async_elem<(int)sizeof(T)>(dst + i, src + i);
}

// The registers of one chain's half-warp for the ordinary step.
template <typename T, int D>
struct Big2Rows {
  T S[D], A[D], Sn[D];      // row h of S_k / Ld_k, of A_k / Ls_k, of the accumulating -Ls Ls^T
  T r, rn;                   // entry h of the right-hand side, of the accumulating -Ls x
  T sb, ab, snb;             // border row 16: entry [16][h] of S / A / Sn      (D = 17 only)
  T sbb, abb, snbb, rb, rnb;  // corner [16][16] of S / A / Sn, entry 16 of r / rn (replicated)
};

// One ordinary step on registers: factorises [S; A] column by column and ACCUMULATES Sn -= Ls Ls^T,
// rn -= Ls x.  With `ob` != nullptr the finished columns of Ld / Ls and the entries of x are written to it as
// they appear (layout Ld [D,D] | Ls [D,D] | x [D], upper triangle of Ld zero), so the panel's registers die
// pivot by pivot.  `col` = 2 x 2 x 16 values of shared memory private to the half-warp.  Returns false on a
// non-positive pivot.
template <typename T, int D, bool WRITE, bool DET>
__device__ __forceinline__ bool big2_step(Big2Rows<T, D>& R, T* col, int h, T* ob, LogProd<T>* det) {
  using C = Big2<T, D>;
  constexpr int DM = C::DM, DD = C::DD;
  bool ok = true;
  T* obr = ob + h * D;  // this lane's row of the staged Ld (obr + DD: of Ls)
#pragma unroll
  for (int j = 0; j < DM; ++j) {
    T* cl = col + (j & 1) * 32;  // [0,16): column j of L, [16,32): column j of Ls
    const T piv = __shfl_sync(0xffffffffu, R.S[j], j, 16);
    ok = ok && (piv > T(0));
    const T rinv = Num<T>::rsqrt_seq(piv);
    if (DET) det->mul(piv);
    const T lij = R.S[j] * rinv, lsij = R.A[j] * rinv;
    const T xj = __shfl_sync(0xffffffffu, R.r, j, 16) * rinv;
    if (WRITE && (DM == 16 || h < DM)) {
      obr[j] = (h >= j) ? lij : T(0);
      obr[DD + j] = lsij;
      if (h == j) ob[2 * DD + j] = xj;
    }
    R.r = Num<T>::fma(-lij, xj, R.r);
    R.rn = Num<T>::fma(-lsij, xj, R.rn);
    T l16 = T(0), ls16 = T(0);
    if constexpr (C::BORDER) {
      l16 = __shfl_sync(0xffffffffu, R.sb, j, 16) * rinv;
      ls16 = __shfl_sync(0xffffffffu, R.ab, j, 16) * rinv;
      if (WRITE && h == j) {
        ob[16 * D + j] = l16;
        ob[DD + 16 * D + j] = ls16;
      }
      R.rb = Num<T>::fma(-l16, xj, R.rb);
      R.rnb = Num<T>::fma(-ls16, xj, R.rnb);
    }
    cl[h] = lij;
    cl[16 + h] = lsij;
    __syncwarp();
#pragma unroll
    for (int c = (j + 1) & ~1; c < DM; c += 2) {
      const auto lc = big2_ld2<T>(cl + c);
      if (c > j) {
        R.S[c] = Num<T>::fma(-lij, lc.x, R.S[c]);
        R.A[c] = Num<T>::fma(-lsij, lc.x, R.A[c]);
      }
      if (c + 1 < DM) {
        R.S[c + 1] = Num<T>::fma(-lij, lc.y, R.S[c + 1]);
        R.A[c + 1] = Num<T>::fma(-lsij, lc.y, R.A[c + 1]);
      }
    }
#pragma unroll
    for (int c = 0; c < DM; c += 2) {
      const auto l2 = big2_ld2<T>(cl + 16 + c);
      R.Sn[c] = Num<T>::fma(-lsij, l2.x, R.Sn[c]);
      if (c + 1 < DM) R.Sn[c + 1] = Num<T>::fma(-lsij, l2.y, R.Sn[c + 1]);
    }
    if constexpr (C::BORDER) {
      R.A[16] = Num<T>::fma(-lsij, l16, R.A[16]);
      if (h > j) {
        R.sb = Num<T>::fma(-l16, lij, R.sb);
        R.ab = Num<T>::fma(-ls16, lij, R.ab);
      }
      R.sbb = Num<T>::fma(-l16, l16, R.sbb);
      R.abb = Num<T>::fma(-ls16, l16, R.abb);
      R.snb = Num<T>::fma(-ls16, lsij, R.snb);
      R.snbb = Num<T>::fma(-ls16, ls16, R.snbb);
    }
  }
  if constexpr (C::BORDER) {  // the border pivot (column 16)
    T* cl = col + (DM & 1) * 32;
    const T piv = R.sbb;
    ok = ok && (piv > T(0));
    const T rinv = Num<T>::rsqrt_seq(piv);
    if (DET) det->mul(piv);
    const T ls = R.A[16] * rinv;  // Ls[h][16]
    R.abb = R.abb * rinv;         // Ls[16][16]
    const T xb = R.rb * rinv;
    if (WRITE) {
      obr[16] = T(0);
      obr[DD + 16] = ls;
      if (h == 0) {
        ob[16 * D + 16] = piv * rinv;
        ob[DD + 16 * D + 16] = R.abb;
        ob[2 * DD + 16] = xb;
      }
    }
    R.rn = Num<T>::fma(-ls, xb, R.rn);
    R.rnb = Num<T>::fma(-R.abb, xb, R.rnb);
    cl[16 + h] = ls;
    __syncwarp();
#pragma unroll
    for (int c = 0; c < DM; c += 2) {
      const auto l2 = big2_ld2<T>(cl + 16 + c);
      R.Sn[c] = Num<T>::fma(-ls, l2.x, R.Sn[c]);
      if (c + 1 < DM) R.Sn[c + 1] = Num<T>::fma(-ls, l2.y, R.Sn[c + 1]);
    }
    R.snb = Num<T>::fma(-R.abb, ls, R.snb);
    R.snbb = Num<T>::fma(-R.abb, R.abb, R.snbb);
  }
  return ok;
}

// rows of one step's blocks, shared memory (global layout) -> registers; `first` uses the blocks as they are,
// otherwise the accumulated -Ls_{k-1} Ls_{k-1}^T / -Ls_{k-1} x_{k-1} are added and the accumulators cleared
template <typename T, int D>
__device__ __forceinline__ void big2_load_rows(Big2Rows<T, D>& R, const T* dg, const T* sb, const T* rh, int h,
                                               bool has_sub, bool has_rhs, bool first) {
  using C = Big2<T, D>;
  const int row = h < C::DM ? h : 0;
#pragma unroll
  for (int c = 0; c < D; ++c) {
    const T add = first ? T(0) : R.Sn[c];
    R.S[c] = dg[row * D + c] + add;
    R.A[c] = has_sub ? sb[row * D + c] : T(0);
    R.Sn[c] = T(0);
  }
  R.r = (has_rhs ? rh[row] : T(0)) + (first ? T(0) : R.rn);
  R.rn = T(0);
  if constexpr (C::BORDER) {
    R.sb = dg[16 * D + h] + (first ? T(0) : R.snb);
    R.sbb = dg[16 * D + 16] + (first ? T(0) : R.snbb);
    R.ab = has_sub ? sb[16 * D + h] : T(0);
    R.abb = has_sub ? sb[16 * D + 16] : T(0);
    R.rb = (has_rhs ? rh[16] : T(0)) + (first ? T(0) : R.rnb);
    R.snb = R.snbb = R.rnb = T(0);
  }
}

// first failing step of a chain across its segments: smallest non-zero value wins
__device__ __forceinline__ void big2_atomic_min_nonzero(int32_t* a, int32_t v) {
  int32_t old = *reinterpret_cast<volatile int32_t*>(a);
  while (old == 0 || v < old) {
    const int32_t seen = atomicCAS(a, old, v);
    if (seen == old) break;
    old = seen;
  }
}

struct Big2Plan {
  int64_t P, L;  // segments per chain, steps per segment
};

// ---------------------------------------------------------------------------------------------------
// The ordinary sweep of segments [p_lo, p_hi) of every chain (virtual chains p-major: both halves of a warp
// work on the same segment index).  Segment 0 starts at the chain's first block, segment p > 0 from its seed.
// ---------------------------------------------------------------------------------------------------
template <typename T, int D, bool DET>
__global__ void __maxnreg__(224)  // no spills (a 200-register cap for 5 CTAs per SM spills into the pivot loop and
                                  // measured 20 % slower); 4 CTAs of 2 warps per SM
big2_factor_kernel(const T* __restrict__ diag, const T* __restrict__ sub, const T* __restrict__ rhs, T* od, T* os,
                   T* ox, T* __restrict__ logdet_part, int32_t* __restrict__ info, const T* __restrict__ seeds,
                   T* __restrict__ seed_out, int64_t B, int64_t Tn, Big2Plan plan, int64_t p_lo, int64_t p_hi) {
  using C = Big2<T, D>;
  constexpr int DD = C::DD, DM = C::DM;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, h = lane & 15;
  const int64_t vpair = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t v = vpair * 2 + (lane >> 4);  // p-major: v = (p - p_lo) * Bpad + b
  const int64_t Bpad = (B + 1) & ~int64_t(1);
  const int64_t p = p_lo + v / Bpad, b = v % Bpad;
  if (p >= p_hi) return;  // whole warp: both halves share p
  const bool live = b < B;
  const int64_t bc = live ? b : B - 1;  // the idle half of an odd batch shadows a real chain, writes nothing
  T* base = reinterpret_cast<T*>(smem_raw) + (size_t)((threadIdx.x >> 4)) * C::PER_CHAIN;
  T* col = base + C::NST * C::STAGE;
  const int64_t k0 = p * plan.L;
  int64_t n = Tn - k0;
  if (n > plan.L) n = plan.L;
  const T* dp = diag + bc * Tn * DD;
  const T* sp = sub ? sub + bc * (Tn - 1) * DD : nullptr;
  const T* rp = rhs ? rhs + bc * Tn * D : nullptr;
  T* odp = od + bc * Tn * DD;
  T* osp = os ? os + bc * (Tn - 1) * DD : nullptr;
  T* oxp = ox ? ox + bc * Tn * D : nullptr;
  auto prefetch = [&](int64_t k) {  // step k of the chain -> stage k % NST
    if (k < k0 + n) {
      T* st = base + (size_t)((k - k0) % C::NST) * C::STAGE;
      big2_copy_async<T>(st, dp + k * DD, DD, h);
      if (sp && k + 1 < Tn) big2_copy_async<T>(st + DD, sp + k * DD, DD, h);
      if (rp) big2_copy_async<T>(st + 2 * DD, rp + k * D, D, h);
    }
    big2_commit();
  };
  Big2Rows<T, D> R;
  LogProd<T> det;
  det.init();
  int32_t fail = 0;
  prefetch(k0);
  for (int64_t k = k0; k < k0 + n; ++k) {
    T* st = base + (size_t)((k - k0) % C::NST) * C::STAGE;
    T* ob = st;  // the stage is free once its rows are in registers: it stages this step's results
    big2_wait<0>();
    __syncwarp();
    const bool has_sub = sp && (k + 1 < Tn);
    const bool first = (k == k0);
    if (first && p > 0) {
      // seed (S, r) of the segment: computed by the fold, row-major like a block
      const T* sd = seeds + ((bc * plan.P + p) * C::SEED);
      big2_load_rows<T, D>(R, sd, st + DD, sd + DD, h, has_sub, true, true);
      if (!rp) {
        R.r = T(0);
        if constexpr (C::BORDER) R.rb = T(0);
      }
    } else {
      big2_load_rows<T, D>(R, st, st + DD, st + 2 * DD, h, has_sub, rp != nullptr, first);
    }
    __syncwarp();
    prefetch(k + 1);  // into the other stage: its results (step k-1) have been written back
    const bool ok = big2_step<T, D, true, DET>(R, col, h, ob, &det);
    if (!ok && fail == 0) fail = (int32_t)(k + 1);
    __syncwarp();
    if (live) {
      for (int i = h; i < DD; i += 16) odp[k * DD + i] = ob[i];
      if (has_sub && osp)
        for (int i = h; i < DD; i += 16) osp[k * DD + i] = ob[DD + i];
      if (oxp && h < D) oxp[k * D + h] = ob[2 * DD + h];
      if (oxp && C::BORDER && h == 0) oxp[k * D + 16] = ob[2 * DD + 16];
    }
    __syncwarp();
  }
  big2_wait<0>();
  // the seed of the NEXT segment: S = D_{k1+1} - Ls Ls^T, r = b_{k1+1} - Ls x of this segment's last step
  if (seed_out && live && k0 + n < Tn) {
    T* so = seed_out + ((bc * plan.P + p + 1) * C::SEED);
    const T* dn = dp + (k0 + n) * DD;
    const T* bn = rp ? rp + (k0 + n) * D : nullptr;
    if (h < DM) {
#pragma unroll
      for (int c = 0; c < D; ++c) so[h * D + c] = dn[h * D + c] + R.Sn[c];
      so[DD + h] = (bn ? bn[h] : T(0)) + R.rn;
    }
    if constexpr (C::BORDER) {
      so[16 * D + h] = dn[16 * D + h] + R.snb;
      if (h == 0) {
        so[16 * D + 16] = dn[16 * D + 16] + R.snbb;
        so[DD + 16] = (bn ? bn[16] : T(0)) + R.rnb;
      }
    }
  }
  if (live && h == 0) {
    if (DET && logdet_part) logdet_part[bc * plan.P + p] = T(0.5) * det.log_abs();
    if (info && fail) big2_atomic_min_nonzero(info + bc, fail);
  }
}

// ---------------------------------------------------------------------------------------------------
// pass 2: per chain, apply the elements of segments 1..P-2 in order.  Applying (P, Q, R, p, r) to (S_in, r_in)
// IS one ordinary step with S := S_in + R, A := Q, r := r_in + r: the trailing block gives S_out - P.
// ---------------------------------------------------------------------------------------------------
template <typename T, int D>
__global__ void __launch_bounds__(64)
big2_fold_kernel(const T* __restrict__ elems, T* seeds, int32_t* __restrict__ info, int64_t B, Big2Plan plan) {
  using C = Big2<T, D>;
  constexpr int DD = C::DD, DM = C::DM;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, h = lane & 15;
  const int64_t b = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 2 + (lane >> 4);
  const bool live = b < B;
  const int64_t bc = live ? b : B - 1;
  T* col = reinterpret_cast<T*>(smem_raw) + (size_t)(threadIdx.x >> 4) * C::COLS;
  const int row = h < DM ? h : 0;
  for (int64_t p = 1; p + 1 < plan.P; ++p) {
    const T* e = elems + (bc * plan.P + p) * C::ELEM;
    const T* sd = seeds + (bc * plan.P + p) * C::SEED;
    Big2Rows<T, D> R;
#pragma unroll
    for (int c = 0; c < D; ++c) {
      R.S[c] = sd[row * D + c] + e[2 * DD + row * D + c];
      R.A[c] = e[DD + row * D + c];
      R.Sn[c] = T(0);
    }
    R.r = sd[DD + row] + e[3 * DD + D + row];
    R.rn = T(0);
    if constexpr (C::BORDER) {
      R.sb = sd[16 * D + h] + e[2 * DD + 16 * D + h];
      R.sbb = sd[16 * D + 16] + e[2 * DD + 16 * D + 16];
      R.ab = e[DD + 16 * D + h];
      R.abb = e[DD + 16 * D + 16];
      R.rb = sd[DD + 16] + e[3 * DD + D + 16];
      R.snb = R.snbb = R.rnb = T(0);
    }
    const bool ok = big2_step<T, D, false, false>(R, col, h, (T*)nullptr, nullptr);
    // a failing pivot here means the chain's matrix is not positive definite somewhere before this segment's
    // end; the sweeps of pass 3 locate the block, this only guarantees the failure is not lost
    if (!ok && live && h == 0 && info) big2_atomic_min_nonzero(info + bc, (int32_t)((p + 1) * plan.L));
    __syncwarp();
    if (live) {
      T* so = seeds + (bc * plan.P + p + 1) * C::SEED;
      if (h < DM) {
#pragma unroll
        for (int c = 0; c < D; ++c) so[h * D + c] = e[h * D + c] + R.Sn[c];
        so[DD + h] = e[3 * DD + h] + R.rn;
      }
      if constexpr (C::BORDER) {
        so[16 * D + h] = e[16 * D + h] + R.snb;
        if (h == 0) {
          so[16 * D + 16] = e[16 * D + 16] + R.snbb;
          so[DD + 16] = e[3 * DD + 16] + R.rnb;
        }
      }
    }
    __syncwarp();
    __threadfence_block();
  }
}

// ---------------------------------------------------------------------------------------------------
// pass 1: the element (P, Q, R, p, r) of segments 1..P-2.  The boundary variable is the segment's first state;
// every later state k is eliminated from [[R, Q^T, 0], [Q, P, A_k^T], [0, A_k, D_{k+1}]]:
//     L = chol P,  Ls = A_k L^-T,  Vt = Q^T L^-T,  y = L^-1 p
//     P' = D_{k+1} - Ls Ls^T,  Q'^T = -Vt Ls^T,  R' = R - Vt Vt^T,  p' = b_{k+1} - Ls y,  r' = r - Vt y
// Lane h owns row h of P, A, Q^T (three "panel" rows that share the pivot columns) and of the three trailing
// blocks.  Phase 1 of a step factorises the panel and leaves ALL pivot columns of L / Ls / Vt in shared memory;
// phase 2 applies the three trailing updates from them (fewer registers live at once than a fused loop).
// ---------------------------------------------------------------------------------------------------
template <typename T, int D>
struct Big2E {
  using C = Big2<T, D>;
  static constexpr int NCOL = 3 * D * 16;  // columns j = 0..D-1 of L | Ls | Vt, 16 entries each
  static constexpr int PER_CHAIN = C::NST * C::STAGE + NCOL;
};

template <typename T, int D>
__global__ void __launch_bounds__(32)
big2_element_kernel(const T* __restrict__ diag, const T* __restrict__ sub, const T* __restrict__ rhs,
                    T* __restrict__ elems, int32_t* __restrict__ info, int64_t B, int64_t Tn, Big2Plan plan) {
  using C = Big2<T, D>;
  using E = Big2E<T, D>;
  constexpr int DD = C::DD, DM = C::DM;
  constexpr bool BORDER = C::BORDER;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, h = lane & 15;
  const int64_t vpair = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t v = vpair * 2 + (lane >> 4);
  const int64_t Bpad = (B + 1) & ~int64_t(1);
  const int64_t p = 1 + v / Bpad, b = v % Bpad;
  if (p + 1 >= plan.P) return;
  const bool live = b < B;
  const int64_t bc = live ? b : B - 1;
  T* base = reinterpret_cast<T*>(smem_raw) + (size_t)(threadIdx.x >> 4) * E::PER_CHAIN;
  T* cols = base + C::NST * C::STAGE;  // [3][D][16]
  const int64_t k0 = p * plan.L;
  const int64_t n = plan.L;  // segments 1..P-2 are full
  const T* dp = diag + bc * Tn * DD;
  const T* sp = sub + bc * (Tn - 1) * DD;
  const T* rp = rhs ? rhs + bc * Tn * D : nullptr;
  const int row = h < DM ? h : 0;
  // stage of step k holds D_{k+1}, A_k (transposed use at the first step), b_{k+1}
  auto prefetch = [&](int64_t k) {
    if (k < k0 + n) {
      T* st = base + (size_t)((k - k0) % C::NST) * C::STAGE;
      big2_copy_async<T>(st, dp + (k + 1) * DD, DD, h);
      big2_copy_async<T>(st + DD, sp + k * DD, DD, h);
      if (rp) big2_copy_async<T>(st + 2 * DD, rp + (k + 1) * D, D, h);
    }
    big2_commit();
  };
  // state: panel rows S (P), A, V (Q^T) and trailing rows Sn, Vn, Rr; vectors r (p), rn, rv
  T S[D], A[D], V[D], Sn[D], Vn[D], Rr[D];
  T r = T(0), rn = T(0), rv = T(0);
  T sb = T(0), ab = T(0), vb = T(0), snb = T(0), vnb = T(0), rrb = T(0);
  T sbb = T(0), abb = T(0), vbb = T(0), snbb = T(0), vnbb = T(0), rrbb = T(0), rb = T(0), rnb = T(0), rvb = T(0);
  bool ok = true;
  prefetch(k0);
  // ---- first step k0: the element after the boundary state alone is (D_{k0+1}, A_{k0}, 0, b_{k0+1}, 0) ----
  {
    const T* st = base;
    big2_wait<0>();
    __syncwarp();
#pragma unroll
    for (int c = 0; c < D; ++c) {
      Sn[c] = st[row * D + c];          // P row h (held in Sn: it becomes S of the next step below)
      Vn[c] = st[DD + c * D + row];     // Q^T row h = column h of A_{k0}
      Rr[c] = T(0);
    }
    rn = rp ? st[2 * DD + row] : T(0);
    if constexpr (BORDER) {
      snb = st[16 * D + h];
      snbb = st[16 * D + 16];
      vnb = st[DD + h * D + 16];        // Q^T[16][h] = A[h][16]
      vnbb = st[DD + 16 * D + 16];
      rnb = rp ? st[2 * DD + 16] : T(0);
    }
    __syncwarp();
    prefetch(k0 + 1);
  }
  for (int64_t k = k0 + 1; k < k0 + n; ++k) {
    const T* st = base + (size_t)((k - k0) % C::NST) * C::STAGE;
    big2_wait<0>();
    __syncwarp();
    // roll the element into the panel: S <- P (held in Sn), V <- Q^T (held in Vn); load A_k; reset trailing
#pragma unroll
    for (int c = 0; c < D; ++c) {
      S[c] = Sn[c];
      V[c] = Vn[c];
      A[c] = st[DD + row * D + c];
    }
    r = rn;
    rn = rp ? st[2 * DD + row] : T(0);
    if constexpr (BORDER) {
      sb = snb; sbb = snbb; vb = vnb; vbb = vnbb; rb = rnb;
      ab = st[DD + 16 * D + h];
      abb = st[DD + 16 * D + 16];
      rnb = rp ? st[2 * DD + 16] : T(0);
    }
    __syncwarp();
    prefetch(k + 1);
    // ---- phase 1: factorise the panel [S; A; V], pivot columns -> shared memory ----
#pragma unroll
    for (int j = 0; j < DM; ++j) {
      const T piv = __shfl_sync(0xffffffffu, S[j], j, 16);
      ok = ok && (piv > T(0));
      const T rinv = Num<T>::rsqrt_seq(piv);
      const T lij = S[j] * rinv, lsij = A[j] * rinv, vtij = V[j] * rinv;
      A[j] = lsij;
      V[j] = vtij;
      const T xj = __shfl_sync(0xffffffffu, r, j, 16) * rinv;
      r = Num<T>::fma(-lij, xj, r);
      rn = Num<T>::fma(-lsij, xj, rn);
      rv = Num<T>::fma(-vtij, xj, rv);
      T l16 = T(0), ls16 = T(0), vt16 = T(0);
      if constexpr (BORDER) {
        l16 = __shfl_sync(0xffffffffu, sb, j, 16) * rinv;
        ls16 = __shfl_sync(0xffffffffu, ab, j, 16) * rinv;
        vt16 = __shfl_sync(0xffffffffu, vb, j, 16) * rinv;
        if (h == j) { ab = ls16; vb = vt16; }
        rb = Num<T>::fma(-l16, xj, rb);
        rnb = Num<T>::fma(-ls16, xj, rnb);
        rvb = Num<T>::fma(-vt16, xj, rvb);
      }
      cols[(0 * D + j) * 16 + h] = lij;
      cols[(1 * D + j) * 16 + h] = lsij;
      cols[(2 * D + j) * 16 + h] = vtij;
      __syncwarp();
      const T* cl = cols + (0 * D + j) * 16;
#pragma unroll
      for (int c = (j + 1) & ~1; c < DM; c += 2) {
        const auto lc = big2_ld2<T>(cl + c);
        if (c > j) {
          S[c] = Num<T>::fma(-lij, lc.x, S[c]);
          A[c] = Num<T>::fma(-lsij, lc.x, A[c]);
          V[c] = Num<T>::fma(-vtij, lc.x, V[c]);
        }
        if (c + 1 < DM) {
          S[c + 1] = Num<T>::fma(-lij, lc.y, S[c + 1]);
          A[c + 1] = Num<T>::fma(-lsij, lc.y, A[c + 1]);
          V[c + 1] = Num<T>::fma(-vtij, lc.y, V[c + 1]);
        }
      }
      if constexpr (BORDER) {
        A[16] = Num<T>::fma(-lsij, l16, A[16]);
        V[16] = Num<T>::fma(-vtij, l16, V[16]);
        if (h > j) {
          sb = Num<T>::fma(-l16, lij, sb);
          ab = Num<T>::fma(-ls16, lij, ab);
          vb = Num<T>::fma(-vt16, lij, vb);
        }
        sbb = Num<T>::fma(-l16, l16, sbb);
        abb = Num<T>::fma(-ls16, l16, abb);
        vbb = Num<T>::fma(-vt16, l16, vbb);
      }
    }
    T xb = T(0);
    if constexpr (BORDER) {  // border pivot: column 16 of Ls / Vt
      const T piv = sbb;
      ok = ok && (piv > T(0));
      const T rinv = Num<T>::rsqrt_seq(piv);
      A[16] = A[16] * rinv;
      V[16] = V[16] * rinv;
      abb = abb * rinv;
      vbb = vbb * rinv;
      xb = rb * rinv;
      rn = Num<T>::fma(-A[16], xb, rn);
      rv = Num<T>::fma(-V[16], xb, rv);
      rnb = Num<T>::fma(-abb, xb, rnb);
      rvb = Num<T>::fma(-vbb, xb, rvb);
      cols[(1 * D + 16) * 16 + h] = A[16];
      cols[(2 * D + 16) * 16 + h] = V[16];
      __syncwarp();
    }
    // ---- phase 2: trailing updates from the stored columns: Sn -= Ls Ls^T, Vn -= Vt Ls^T, Rr -= Vt Vt^T ----
    // (the trailing block of the next state starts from D_{k+1}, still in this step's stage)
#pragma unroll
    for (int c = 0; c < D; ++c) {
      Sn[c] = st[row * D + c];
      Vn[c] = T(0);
    }
    if constexpr (BORDER) {
      snb = st[16 * D + h];
      snbb = st[16 * D + 16];
      vnb = T(0);
      vnbb = T(0);
    }
#pragma unroll
    for (int j = 0; j < D; ++j) {
      const T* c2 = cols + (1 * D + j) * 16;
      const T* c3 = cols + (2 * D + j) * 16;
      const T lsij = A[j], vtij = V[j];
#pragma unroll
      for (int c = 0; c < DM; c += 2) {
        const auto l2 = big2_ld2<T>(c2 + c);
        const auto l3 = big2_ld2<T>(c3 + c);
        Sn[c] = Num<T>::fma(-lsij, l2.x, Sn[c]);
        Vn[c] = Num<T>::fma(-vtij, l2.x, Vn[c]);
        Rr[c] = Num<T>::fma(-vtij, l3.x, Rr[c]);
        if (c + 1 < DM) {
          Sn[c + 1] = Num<T>::fma(-lsij, l2.y, Sn[c + 1]);
          Vn[c + 1] = Num<T>::fma(-vtij, l2.y, Vn[c + 1]);
          Rr[c + 1] = Num<T>::fma(-vtij, l3.y, Rr[c + 1]);
        }
      }
      if constexpr (BORDER) {
        // column 16 of the rows of Q'^T, and the border rows [16][h] of the three trailing blocks
        const T ls16 = (j < 16) ? __shfl_sync(0xffffffffu, ab, j, 16) : abb;
        const T vt16 = (j < 16) ? __shfl_sync(0xffffffffu, vb, j, 16) : vbb;
        Vn[16] = Num<T>::fma(-vtij, ls16, Vn[16]);
        snb = Num<T>::fma(-ls16, lsij, snb);
        snbb = Num<T>::fma(-ls16, ls16, snbb);
        vnb = Num<T>::fma(-vt16, lsij, vnb);
        vnbb = Num<T>::fma(-vt16, ls16, vnbb);
        rrb = Num<T>::fma(-vt16, vtij, rrb);
        rrbb = Num<T>::fma(-vt16, vt16, rrbb);
      }
    }
    __syncwarp();
  }
  big2_wait<0>();
  // ---- write the element: P | Q | R | p | r (Q = (Q^T)^T: lane h writes column h) ----
  if (live) {
    T* e = elems + (bc * plan.P + p) * C::ELEM;
    if (h < DM) {
#pragma unroll
      for (int c = 0; c < D; ++c) {
        e[h * D + c] = Sn[c];
        e[DD + c * D + h] = Vn[c];
        e[2 * DD + h * D + c] = Rr[c];
      }
      e[3 * DD + h] = rn;
      e[3 * DD + D + h] = rv;
    }
    if constexpr (BORDER) {
      e[16 * D + h] = snb;
      e[DD + h * D + 16] = vnb;  // Q[h][16] = Q^T[16][h]
      e[2 * DD + 16 * D + h] = rrb;
      if (h == 0) {
        e[16 * D + 16] = snbb;
        e[DD + 16 * D + 16] = vnbb;
        e[2 * DD + 16 * D + 16] = rrbb;
        e[3 * DD + 16] = rnb;
        e[3 * DD + D + 16] = rvb;
      }
    }
    if (!ok && h == 0 && info) big2_atomic_min_nonzero(info + bc, (int32_t)(k0 + n));
  }
}

}  // namespace mf
