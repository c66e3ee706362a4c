// Large-block path for FEW chains (MF_SMALL_D_MAX < D <= MF_BIG_D_MAX): ONE CTA ("team") PER CHAIN.
//
// With a few hundred chains of D x D blocks (config 4: B = 256, D = 17) the run time of the
// Cholesky(+solve) sweep (reference block_tri_diag.py:436,350 -> cholesky_band / solve_triang_mat)
// is  T x (dependent latency of one block step):  S_{k+1} = D_{k+1} - A_k S_k^{-1} A_k^T  cannot
// start before S_k is eliminated.  This kernel shortens that critical path instead of adding
// throughput:
//
//  * the step is ONE elimination of the D pivot columns of the augmented panel
//        [ S_k  ]   D x D   (lower triangle)  + the right-hand side as an extra row  b_k^T
//        [ A_k  ]   D x D
//    followed by the Schur update of the next block  [D_{k+1}; b_{k+1}^T] -= [Ls; x^T] Ls^T;
//  * FRACTION-FREE elimination: M' = (p M - u_r u_c) 2^{-e(p)}  (p = current pivot, 2^{-e} its
//    exact power-of-two normaliser).  No reciprocal / rsqrt sits on the dependent path; the true
//    factors are recovered off the path as  L_ij = M_ij rsqrt(sigma_j p_j)  with the running scale
//    sigma_{j+1} = sigma_j p_j 2^{-e_j}  (sigma stays within [1, 2^D), any input scale is safe);
//  * elements, not rows, are distributed over lanes (column-major over the lower triangle), so a
//    stage costs ceil(live elements / 32) updates per lane and column j+1 is published first;
//  * four compute warps on four SM sub-partitions: W0/W1 alternate between the pivot role ("F":
//    eliminates [S_k; b_k^T]) and the Schur role ("N": accumulates the next block in registers and
//    BECOMES the pivot warp of step k+1, so S never travels); A0/A1 own the rows of A_k.  Columns
//    travel through shared memory, ordered by one mbarrier per column;
//  * HBM <-> shared memory by TMA bulk copies (one per stream and tile of K steps) issued by two
//    producer warps; results are assembled in the output stage in the global layout and leave as
//    bulk stores.  In place (out == in) is allowed: a tile is always loaded before it is stored.
#pragma once
#include <cstdint>

#include "dispatch.cuh"
#include "ssm_kernels.cuh"
#include "sweep.cuh"

namespace mf {

#ifdef MF_TEAM_DEBUG
__device__ long long g_team_dbg[4 * 40];
#define MF_TEAM_STAMP(w, idx) \
  if ((k == 51 || (k == 52 && (idx) == 0)) && blockIdx.x == 0 && lane == 0) g_team_dbg[(w) * 40 + (idx)] = clock64();
#else
#define MF_TEAM_STAMP(w, idx)
#endif

template <int I, int N, typename F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < N) {
    f(IntTag<I>{});
    static_for<I + 1, N>(f);
  }
}

template <typename T>
__device__ __forceinline__ T pow2_inv(T p);
template <>
__device__ __forceinline__ double pow2_inv<double>(double p) {  // 2^-floor(log2 p), exact
  const int e = (__double2hiint(p) >> 20) & 0x7ff;
  return __hiloint2double((2046 - e) << 20, 0);
}
template <>
__device__ __forceinline__ float pow2_inv<float>(float p) {
  const int e = (__float_as_int(p) >> 23) & 0xff;
  return __int_as_float((254 - e) << 23);
}

template <typename T> struct TmVec2;
template <> struct TmVec2<double> { using type = double2; };
template <> struct TmVec2<float> { using type = float2; };
// two consecutive elements (even element index: 2 * sizeof(T)-aligned) in one shared-memory load
template <typename T>
__device__ __forceinline__ typename TmVec2<T>::type tm_ld2(const T* sm, int idx) {
  return *reinterpret_cast<const typename TmVec2<T>::type*>(sm + idx);
}

__device__ __forceinline__ void tm_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tm_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// 1/sqrt(x) for positive normal x without the library's special-case branch (one MUFU seed + a
// third-order correction, full double precision); anything else yields NaN/garbage, which is what a
// failed factorisation may hold.
__device__ __forceinline__ double tm_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(x, -(y * y), 1.0);
  const double c = fma(e, 0.375, 0.5);
  return fma(c, y * e, y);
}
__device__ __forceinline__ float tm_rsqrt(float x) { return rsqrtf(x); }

template <typename T, int D>
struct TeamCfg {
  static constexpr int ES = (int)sizeof(T);
  static constexpr int DD = D * D;
  static constexpr int UP = (D + 2) & ~1;  // entries of one published column (row D = right-hand side), even
  static constexpr int NSI = 3, NSO = 2;
  static constexpr int r16(int b) { return (b + 15) / 16 * 16; }
  static constexpr int reg_blk(int K) { return r16(K * DD * ES + 16); }
  static constexpr int reg_vec(int K) { return r16(K * D * ES + 16); }
  static constexpr int stage_bytes(int K) { return 2 * reg_blk(K) + reg_vec(K); }
  static constexpr int NBAR = 2 * NSI + 2 * NSO + 2 * D + 2;
  static constexpr int COLS_BYTES = r16(4 * D * UP * ES);  // ucol[2][D][UP], acol[2][D][UP]
  static constexpr int FIXED = COLS_BYTES + 8 * NBAR + 512;  // + dump slots
  static constexpr int total(int K) { return stage_bytes(K) * (NSI + NSO) + FIXED; }
  // two CTAs per SM when possible (113 KB each), else one
  static constexpr int K = total(4) <= 113 * 1024 ? 4 : (total(2) <= 113 * 1024 ? 2 : (total(2) <= 227 * 1024 ? 2 : 1));
  static constexpr int STAGE = stage_bytes(K);
  static constexpr int OFF_SUB = reg_blk(K), OFF_VEC = 2 * reg_blk(K);
  static constexpr size_t SMEM_BYTES = (size_t)total(K);
  // W-role: 2 x 2 register tiles over the lower triangle of the (D+1) x D panel [S; b^T]
  // (tile rows NTR, tile columns NTC, tiles with tR >= tC, column-major over tiles)
  static constexpr int NTR = (D + 2) / 2, NTC = (D + 1) / 2;
  static constexpr int tcolstart(int c) { return c * NTR - c * (c - 1) / 2; }
  static constexpr int NTW = tcolstart(NTC);
  static constexpr int NSW = (NTW + 31) / 32;
  static constexpr int wtcol(int e) {
    int c = 0;
    while (c + 1 < NTC && tcolstart(c + 1) <= e) ++c;
    return c;
  }
  static constexpr int wcmin(int s) { return wtcol(32 * s); }
  static constexpr int wcmax(int s) { return wtcol(32 * s + 31 < NTW ? 32 * s + 31 : NTW - 1); }
  // A-role: NR tile rows x NTC tile columns per warp (two warps), column-major over tiles
  static constexpr int NTRA = (D + 1) / 2;
  static constexpr int NR = (NTRA + 1) / 2;
  static constexpr int NTA = NR * NTC;
  static constexpr int NSA = (NTA + 31) / 32;
  static constexpr int acmin(int s) { return (32 * s) / NR; }
  static constexpr int acmax(int s) { return (32 * s + 31 < NTA ? 32 * s + 31 : NTA - 1) / NR; }
  static constexpr int THREADS = 32 * 7;
};

template <typename T, int D>
__global__ void __launch_bounds__(TeamCfg<T, D>::THREADS)
btd_chol_team_kernel(const T* __restrict__ diag, const T* __restrict__ sub,
                     const T* __restrict__ rhs, T* od, T* os, T* ox, T* __restrict__ logdet,
                     int32_t* __restrict__ info, int64_t B, int64_t Tn) {
  using Cfg = TeamCfg<T, D>;
  constexpr int ES = Cfg::ES, DD = Cfg::DD, UP = Cfg::UP, K = Cfg::K, NSI = Cfg::NSI, NSO = Cfg::NSO;
  constexpr int NSW = Cfg::NSW, NSA = Cfg::NSA, NR = Cfg::NR;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  char* in_stages = reinterpret_cast<char*>(smem_raw);
  char* out_stages = in_stages + (size_t)Cfg::STAGE * NSI;
  uint64_t* bars = reinterpret_cast<uint64_t*>(out_stages + (size_t)Cfg::STAGE * NSO + Cfg::COLS_BYTES);
  uint64_t* full_in = bars;
  uint64_t* consumed = bars + NSI;
  uint64_t* full_out = bars + 2 * NSI;
  uint64_t* empty_out = bars + 2 * NSI + NSO;
  uint64_t* bar_f = bars + 2 * NSI + 2 * NSO;  // raw column j of the pivot panel published (1 arrival)
  uint64_t* bar_n = bar_f + D;                 // raw columns j of both panels published (3 arrivals)
  uint64_t* o_done = bar_n + D;                // output warp finished the columns of a step (per parity)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t chain = blockIdx.x;
  const int64_t ntiles = (Tn + K - 1) / K;
  const bool has_sub = sub != nullptr && Tn > 1;

  const T* dp = diag + chain * Tn * DD;
  const T* sp = has_sub ? sub + chain * (Tn - 1) * DD : nullptr;
  const T* rp = rhs ? rhs + chain * Tn * D : nullptr;
  T* odp = od + chain * Tn * DD;
  T* osp = (os && has_sub) ? os + chain * (Tn - 1) * DD : nullptr;
  T* oxp = (ox && rhs) ? ox + chain * Tn * D : nullptr;
  // misalignment of every stream: data sits at +a0 inside its region so that bulk copies are
  // 16-byte aligned on both sides
  const int a0d = (int)(reinterpret_cast<uintptr_t>(dp) & 15);
  const int a0s = (int)(reinterpret_cast<uintptr_t>(sp) & 15);
  const int a0r = (int)(reinterpret_cast<uintptr_t>(rp) & 15);
  const int b0d = (int)(reinterpret_cast<uintptr_t>(odp) & 15);
  const int b0s = (int)(reinterpret_cast<uintptr_t>(osp) & 15);
  const int b0r = (int)(reinterpret_cast<uintptr_t>(oxp) & 15);

  // zero the output stages once: the upper triangles of Ld are never written again
  for (int i = threadIdx.x; i < Cfg::STAGE * NSO / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(out_stages)[i] = 0u;
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    for (int i = 0; i < NSI; ++i) {
      mbar_init(full_in + i, 6);       // 3 loader lanes x (expect_tx arrive + cp.async arrive)
      mbar_init(consumed + i, 4 * K);  // W0, W1, A0, A1 x K steps
    }
    for (int i = 0; i < NSO; ++i) {
      mbar_init(full_out + i, K);  // O x K steps
      mbar_init(empty_out + i, 3);
    }
    for (int j = 0; j < D; ++j) {
      mbar_init(bar_f + j, 1);
      mbar_init(bar_n + j, 3);
    }
    mbar_init(o_done, 1);
    mbar_init(o_done + 1, 1);
    mbar_fence_init();
  }
  __syncthreads();

  // ------------------------------------ producer warps ------------------------------------------
  if (warp == 5) {
    if (lane >= 3) return;
    const int E = lane == 2 ? D : DD;
    const int roff = lane == 0 ? 0 : (lane == 1 ? Cfg::OFF_SUB : Cfg::OFF_VEC);
    StreamGeom g;
    g.step0 = const_cast<char*>(reinterpret_cast<const char*>(lane == 0 ? dp : (lane == 1 ? sp : rp)));
    g.first = 0;
    g.end = lane == 1 ? Tn - 1 : Tn;
    const SweepSeg sg = make_seg(g, g.step0 != nullptr);
    auto issue_load = [&](int64_t t) {
      const int si = (int)(t % NSI);
      uint64_t* bar = full_in + si;
      const int64_t j0 = t * K;
      uint32_t tx = 0;
      int lo = 0, hi = 0, head = 0;
      if (sg.g) tx = sweep_ranges<ES, K>(sg, E, j0, lo, hi, head);
      mbar_arrive_expect_tx(bar, tx);
      if (sg.g && hi > lo) {
        char* sd = in_stages + (size_t)si * Cfg::STAGE + roff + sg.a0;
        const char* g0 = sg.g + j0 * (int64_t)(E * ES);
        if (tx) tma_load_1d(sd + lo + head, g0 + lo + head, tx, bar);
        for (int o = lo; o < lo + head; o += ES) cp_async_elem<ES>(sd + o, g0 + o);
        for (int o = lo + head + (int)tx; o < hi; o += ES) cp_async_elem<ES>(sd + o, g0 + o);
      }
      cp_async_arrive_noinc(bar);
    };
    for (int64_t t = 0; t < NSI && t < ntiles; ++t) issue_load(t);
    for (int64_t t = 0; t + NSI < ntiles; ++t) {
      mbar_wait(consumed + (int)(t % NSI), (uint32_t)((t / NSI) & 1));
      issue_load(t + NSI);
    }
    return;
  }
  if (warp == 6) {
    if (lane >= 3) return;
    const int E = lane == 2 ? D : DD;
    const int roff = lane == 0 ? 0 : (lane == 1 ? Cfg::OFF_SUB : Cfg::OFF_VEC);
    StreamGeom g;
    g.step0 = reinterpret_cast<char*>(lane == 0 ? odp : (lane == 1 ? osp : oxp));
    g.first = 0;
    g.end = lane == 1 ? Tn - 1 : Tn;
    const SweepSeg sg = make_seg(g, g.step0 != nullptr);
    for (int64_t t = 0; t < ntiles; ++t) {
      const int so = (int)(t % NSO);
      mbar_wait(full_out + so, (uint32_t)((t / NSO) & 1));
      if (sg.g) {
        const int64_t j0 = t * K;
        int lo, hi, head;
        const uint32_t tx = sweep_ranges<ES, K>(sg, E, j0, lo, hi, head);
        if (hi > lo) {
          const char* sd = out_stages + (size_t)so * Cfg::STAGE + roff + sg.a0;
          char* g0 = sg.g + j0 * (int64_t)(E * ES);
          if (tx) tma_store_1d(g0 + lo + head, sd + lo + head, tx);
          for (int o = lo; o < lo + head; o += ES)
            *reinterpret_cast<T*>(g0 + o) = *reinterpret_cast<const T*>(sd + o);
          for (int o = lo + head + (int)tx; o < hi; o += ES)
            *reinterpret_cast<T*>(g0 + o) = *reinterpret_cast<const T*>(sd + o);
        }
      }
      tma_store_commit();
      tma_store_wait_read<0>();
      mbar_arrive(empty_out + so);
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    return;
  }

  // ------------------------------------ compute warps -------------------------------------------
  // Barrier operations go through 32-bit shared addresses derived from one opaque base register
  // (the compiler otherwise re-derives the base from SR_CgaCtaId -- a slow S2R -- in front of every
  // barrier operation); data moves with plain indexed loads/stores that ptxas can schedule.
  uint32_t sbase;
  asm volatile("mov.u32 %0, %1;" : "=r"(sbase) : "r"(smem_u32(smem_raw)));
  T* const sm = reinterpret_cast<T*>(smem_raw);
  constexpr int I_OUT = Cfg::STAGE * NSI / ES;           // first output stage (element index)
  constexpr int I_UCOL = Cfg::STAGE * (NSI + NSO) / ES;  // ucol[2][D][UP]
  constexpr int PAR = D * UP;                            // elements of one parity buffer
  constexpr int ACOL = 2 * D * UP;                       // acol[2][D][UP] relative to ucol
  constexpr int SE = Cfg::STAGE / ES;                    // elements per stage
  constexpr int I_DUMP = (Cfg::STAGE * (NSI + NSO) + Cfg::COLS_BYTES + 8 * Cfg::NBAR) / ES;
  const uint32_t a_bars = sbase + (uint32_t)(Cfg::STAGE * (NSI + NSO) + Cfg::COLS_BYTES);
  const uint32_t a_full_in = a_bars, a_consumed = a_bars + 8 * NSI, a_full_out = a_bars + 16 * NSI,
                 a_empty_out = a_bars + 16 * NSI + 8 * NSO, a_bar_f = a_bars + 16 * NSI + 16 * NSO,
                 a_bar_n = a_bar_f + 8 * D, a_o_done = a_bar_n + 8 * D;

  int64_t in_ready = 0, out_ready = 0;  // tiles this warp has seen full (input) / empty (output)
  auto need_in = [&](int64_t tile) {
    while (in_ready <= tile) {
      tm_wait(a_full_in + 8 * (uint32_t)(in_ready % NSI), (uint32_t)((in_ready / NSI) & 1));
      ++in_ready;
    }
  };
  auto need_out = [&](int64_t tile) {
    while (out_ready <= tile) {
      tm_wait(a_empty_out + 8 * (uint32_t)(out_ready % NSO), (uint32_t)(((out_ready / NSO) & 1) ^ 1));
      ++out_ready;
    }
  };
  // bookkeeping arrivals of one warp for the data of step m
  auto done_in = [&](int64_t m) {
    __syncwarp();
    if (lane == 0) tm_arrive(a_consumed + 8 * (uint32_t)((m / K) % NSI));
  };
  auto done_out = [&](int64_t m) {
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) tm_arrive(a_full_out + 8 * (uint32_t)((m / K) % NSO));
  };
  const int64_t Tpad = ntiles * K;
  // every compute warp tracks the running scale sigma (same operations in the same order, so the
  // copies agree bit for bit): sigma <- sigma p 2^-e(p) per pivot, renormalised to [1,2) per step
  T sigma = T(1);
  auto renorm = [&]() -> T {
    const T s2 = pow2_inv<T>(sigma);
    sigma *= s2;
    return s2;
  };

  if (warp < 2) {
    // ===================== W warps: pivot role (F) / Schur role (N), alternating =================
    // slot s of this lane: tile (tR, tC) = rows 2tR, 2tR+1 x columns 2tC, 2tC+1; X[4s + 2a + b] = (r_a, c_b)
    int tC[NSW], r0[NSW];
    int iu_r[NSW], iu_c[NSW];  // element indices of the row pair / column pair inside column 0 of ucol[0]
    int win[NSW][4];           // input index of each element (step 0 of a tile), -1: structurally zero
    bool wvec[NSW];            // the tile's second... row D (right-hand side) is row a = vrow
    int vrow[NSW];
#pragma unroll
    for (int s = 0; s < NSW; ++s) {
      const int e = 32 * s + lane;
      int c = 0;
      while (c + 1 < Cfg::NTC && Cfg::tcolstart(c + 1) <= e) ++c;
      const int tr = c + (e - Cfg::tcolstart(c));
      const bool ok = e < Cfg::NTW;
      tC[s] = ok ? c : 255;  // 255: never a pivot column, never live
      r0[s] = ok ? 2 * tr : 0;
      iu_r[s] = I_UCOL + (ok ? 2 * tr : 0);
      iu_c[s] = I_UCOL + (ok ? 2 * c : 0);
      wvec[s] = ok && (2 * tr == D || 2 * tr + 1 == D);
      vrow[s] = (2 * tr == D) ? 0 : 1;
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const int r = 2 * tr + a, cc = 2 * c + b;
          const bool valid = ok && cc < D && r <= D && r >= cc;
          win[s][2 * a + b] = !valid ? -1 : (r == D ? (Cfg::OFF_VEC + a0r) / ES + cc : a0d / ES + r * D + cc);
        }
    }
    T X[NSW * 4];

    auto load_block = [&](int64_t m) {  // X <- sigma [D_m (lower); b_m^T]
      need_in(m / K);
      const int st = (int)((m / K) % NSI) * SE;
      const int sk = (int)(m % K);
#pragma unroll
      for (int s = 0; s < NSW; ++s)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const bool vec = wvec[s] && (q >> 1) == vrow[s];
          T v = T(0);
          if (win[s][q] >= 0 && (!vec || rp)) v = sm[st + win[s][q] + sk * (vec ? D : DD)];
          X[4 * s + q] = v * sigma;
        }
      done_in(m);
    };

    if (warp == 0) load_block(0);
    for (int64_t k = 0; k < Tn; ++k) {
      const int pb = (int)(k & 1) * PAR;  // parity buffer of this step's columns
      if ((int)(k & 1) == warp) {
        // ------------------------------- F: eliminate [S_k; b_k^T] -------------------------------
        MF_TEAM_STAMP(warp, 0)
        // Runtime loops keep the instruction footprint small (a fully unrolled sweep of D stages
        // per role does not fit the instruction caches and runs fetch-bound).  Phase ph covers the
        // tile columns in which slots >= ph are (partly) live; h = column inside the tile column.
        static_for<0, NSW>([&](auto pt) {
          constexpr int ph = decltype(pt)::value;
          constexpr int C_lo = ph == 0 ? 0 : Cfg::wcmax(ph - 1) + 1;
          constexpr int C_hi = Cfg::wcmax(ph);
#pragma unroll 1
          for (int Cj = C_lo; Cj <= C_hi; ++Cj) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int j = 2 * Cj + h;
              if (j < D) {
                const int jo = pb + j * UP;
                // publish raw column j (rows j..D); the right-hand-side entry also goes to acol row D
#pragma unroll
                for (int s = ph; s < NSW; ++s) {
                  const bool mine = tC[s] == Cj;
                  if (mine && r0[s] >= j) sm[iu_r[s] + jo] = X[4 * s + h];
                  if (mine && r0[s] + 1 <= D) sm[iu_r[s] + 1 + jo] = X[4 * s + 2 + h];
                  if (mine && wvec[s]) sm[I_UCOL + D + ACOL + jo] = X[4 * s + 2 * vrow[s] + h];
                }
                __syncwarp();
                if (lane == 0) {
                  tm_arrive(a_bar_f + 8 * j);
                  tm_arrive(a_bar_n + 8 * j);
                }
                const T p = sm[I_UCOL + jo + j];
                typename TmVec2<T>::type ur[NSW], uc[NSW];
#pragma unroll
                for (int s = ph; s < NSW; ++s) {
                  ur[s] = tm_ld2<T>(sm, iu_r[s] + jo);
                  uc[s] = tm_ld2<T>(sm, iu_c[s] + jo);
                }
                const T sc = pow2_inv<T>(p);
#pragma unroll
                for (int s = ph; s < NSW; ++s) {
                  const T v00 = Num<T>::fma(-ur[s].x, uc[s].x, p * X[4 * s + 0]) * sc;
                  const T v01 = Num<T>::fma(-ur[s].x, uc[s].y, p * X[4 * s + 1]) * sc;
                  const T v10 = Num<T>::fma(-ur[s].y, uc[s].x, p * X[4 * s + 2]) * sc;
                  const T v11 = Num<T>::fma(-ur[s].y, uc[s].y, p * X[4 * s + 3]) * sc;
                  const bool full = tC[s] > Cj && tC[s] != 255;
                  const bool half = full || (h == 0 && tC[s] == Cj);
                  if (full) { X[4 * s + 0] = v00; X[4 * s + 2] = v10; }
                  if (half) { X[4 * s + 1] = v01; X[4 * s + 3] = v11; }
                }
                sigma *= p * sc;
                MF_TEAM_STAMP(warp, j + 1)
              }
            }
          }
        });
        renorm();
      } else {
        // -------- N: fraction-free Schur update of sigma [D_{k+1}; b_{k+1}^T] with raw columns --------
        if (k + 1 < Tn) load_block(k + 1);
        need_in(k / K);
        done_in(k);  // this warp does not read the inputs of step k
        // the pivot role of step k+1 (this warp) overwrites the columns of step k-1: wait until the
        // output warp has read them
        if (k >= 1) tm_wait(a_o_done + 8 * (uint32_t)((k - 1) & 1), (uint32_t)(((k - 1) >> 1) & 1));
        if (k + 1 < Tn && has_sub) {
          const uint32_t par = (uint32_t)(k & 1);
#pragma unroll 1
          for (int j = 0; j < D; ++j) {
            const int jo = pb + j * UP;
            tm_wait(a_bar_n + 8 * j, par);
            const T p = sm[I_UCOL + jo + j];
            typename TmVec2<T>::type lr[NSW], lc[NSW];
#pragma unroll
            for (int s = 0; s < NSW; ++s) {
              lr[s] = tm_ld2<T>(sm, iu_r[s] + ACOL + jo);
              lc[s] = tm_ld2<T>(sm, iu_c[s] + ACOL + jo);
            }
            const T sc = pow2_inv<T>(p);
#pragma unroll
            for (int s = 0; s < NSW; ++s) {
              X[4 * s + 0] = Num<T>::fma(-lr[s].x, lc[s].x, p * X[4 * s + 0]) * sc;
              X[4 * s + 1] = Num<T>::fma(-lr[s].x, lc[s].y, p * X[4 * s + 1]) * sc;
              X[4 * s + 2] = Num<T>::fma(-lr[s].y, lc[s].x, p * X[4 * s + 2]) * sc;
              X[4 * s + 3] = Num<T>::fma(-lr[s].y, lc[s].y, p * X[4 * s + 3]) * sc;
            }
            sigma *= p * sc;
            MF_TEAM_STAMP(warp, 20 + j)
          }
          const T s2 = renorm();
#pragma unroll
          for (int q = 0; q < 4 * NSW; ++q) X[q] *= s2;
        } else {
          sigma = T(1);
        }
      }
    }
    for (int64_t m = Tn; m < Tpad; ++m) done_in(m);
    return;
  }

  if (warp < 4) {
    // ============================= A warps: rows of the A_k panel ================================
    const int aw = warp - 2;
    // slot s of this lane: tile rows i0, i0+1 x columns 2tC, 2tC+1 of the A_k panel
    int tC[NSA], i0[NSA];
    int ia_r[NSA], iu_c[NSA];  // own row pair inside column 0 of acol[0] / column pair of ucol[0]
    int ain[NSA][4];
#pragma unroll
    for (int s = 0; s < NSA; ++s) {
      const int e = 32 * s + lane;
      const int c = e / NR, tr = aw * NR + e % NR;
      const bool ok = e < Cfg::NTA && tr < Cfg::NTRA;
      tC[s] = ok ? c : 255;
      i0[s] = ok ? 2 * tr : 2 * D;
      ia_r[s] = I_UCOL + ACOL + (ok ? 2 * tr : 0);
      iu_c[s] = I_UCOL + (ok ? 2 * c : 0);
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const int i = 2 * tr + a, cc = 2 * c + b;
          ain[s][2 * a + b] = (ok && i < D && cc < D) ? (Cfg::OFF_SUB + a0s) / ES + i * D + cc : -1;
        }
    }
    T A[NSA * 4];
    for (int64_t k = 0; k < Tn; ++k) {
      const int pb = (int)(k & 1) * PAR;
      need_in(k / K);
      const bool active = has_sub && (k + 1 < Tn);
      if (active) {
        const int st = (int)((k / K) % NSI) * SE;
        const int sk = (int)(k % K);
#pragma unroll
        for (int s = 0; s < NSA; ++s)
#pragma unroll
          for (int q = 0; q < 4; ++q)
            A[4 * s + q] = (ain[s][q] >= 0 ? sm[st + ain[s][q] + sk * DD] : T(0)) * sigma;
      }
      done_in(k);
      // acol[parity] is rewritten now: the output warp must have finished step k-2
      if (k >= 2) tm_wait(a_o_done + 8 * (uint32_t)(k & 1), (uint32_t)(((k - 2) >> 1) & 1));
      if (active) {
        const uint32_t par = (uint32_t)(k & 1);
        MF_TEAM_STAMP(warp, 0)
        static_for<0, NSA>([&](auto pt) {
          constexpr int ph = decltype(pt)::value;
          constexpr int C_lo = ph == 0 ? 0 : Cfg::acmax(ph - 1) + 1;
          constexpr int C_hi = Cfg::acmax(ph);
#pragma unroll 1
          for (int Cj = C_lo; Cj <= C_hi; ++Cj) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int j = 2 * Cj + h;
              if (j < D) {
                const int jo = pb + j * UP;
                // own raw column j: final since the previous stage, independent of the pivot column j
#pragma unroll
                for (int s = ph; s < NSA; ++s) {
                  const bool mine = tC[s] == Cj;
                  if (mine && i0[s] < D) sm[ia_r[s] + jo] = A[4 * s + h];
                  if (mine && i0[s] + 1 < D) sm[ia_r[s] + 1 + jo] = A[4 * s + 2 + h];
                }
                __syncwarp();
                if (lane == 0) tm_arrive(a_bar_n + 8 * j);
                tm_wait(a_bar_f + 8 * j, par);
                const T p = sm[I_UCOL + jo + j];
                typename TmVec2<T>::type ar[NSA], uc[NSA];
#pragma unroll
                for (int s = ph; s < NSA; ++s) {
                  ar[s] = tm_ld2<T>(sm, ia_r[s] + jo);
                  uc[s] = tm_ld2<T>(sm, iu_c[s] + jo);
                }
                const T sc = pow2_inv<T>(p);
#pragma unroll
                for (int s = ph; s < NSA; ++s) {
                  const T v00 = Num<T>::fma(-ar[s].x, uc[s].x, p * A[4 * s + 0]) * sc;
                  const T v01 = Num<T>::fma(-ar[s].x, uc[s].y, p * A[4 * s + 1]) * sc;
                  const T v10 = Num<T>::fma(-ar[s].y, uc[s].x, p * A[4 * s + 2]) * sc;
                  const T v11 = Num<T>::fma(-ar[s].y, uc[s].y, p * A[4 * s + 3]) * sc;
                  const bool full = tC[s] > Cj && tC[s] != 255;
                  const bool half = full || (h == 0 && tC[s] == Cj);
                  if (full) { A[4 * s + 0] = v00; A[4 * s + 2] = v10; }
                  if (half) { A[4 * s + 1] = v01; A[4 * s + 3] = v11; }
                }
                sigma *= p * sc;
                MF_TEAM_STAMP(warp, j + 1)
              }
            }
          }
        });
        renorm();
      } else {
        sigma = T(1);
      }
    }
    for (int64_t m = Tn; m < Tpad; ++m) done_in(m);
    return;
  }

  // ====================== O warp: true Ld columns, x, log-determinant, info =======================
  {
    const int r = lane;  // row of the published pivot column (D = right-hand side)
    LogProd<T> det;
    det.init();
    int32_t fail = 0;
    for (int64_t k = 0; k < Tn; ++k) {
      const int pb = (int)(k & 1) * PAR;
      need_out(k / K);
      const int ost = I_OUT + (int)((k / K) % NSO) * SE;
      const int sk = (int)(k % K);
      int obase = r < D ? ost + b0d / ES + sk * DD + r * D
                        : ((r == D && oxp) ? ost + (Cfg::OFF_VEC + b0r) / ES + sk * D : I_DUMP + 16);
      asm volatile("" : "+r"(obase));
      const bool do_ls = has_sub && (k + 1 < Tn);
      int lbase = (r < D && do_ls && osp) ? ost + (Cfg::OFF_SUB + b0s) / ES + sk * DD + r * D : I_DUMP + 16;
      asm volatile("" : "+r"(lbase));
      const uint32_t par = (uint32_t)(k & 1);
      const int32_t kfail = (int32_t)(k + 1);
      T u_prev = T(0), a_prev = T(0), sp_prev = T(1), p_prev = T(1);
#pragma unroll 1
      for (int j = 0; j <= D; ++j) {
        T p = T(1), u = T(0), a = T(0);
        if (j < D) {
          const int jo = pb + j * UP;
          if (do_ls) tm_wait(a_bar_n + 8 * j, par);  // raw columns j of both panels
          else tm_wait(a_bar_f + 8 * j, par);
          p = sm[I_UCOL + jo + j];
          u = sm[I_UCOL + jo + (r <= D ? r : 0)];
          a = sm[I_UCOL + ACOL + jo + (r < D ? r : 0)];
        }
        if (j >= 1) {  // finish column j-1 while the loads of column j are in flight
          const T rsq = tm_rsqrt(sp_prev);
          const T v = u_prev * rsq;
          if (r >= j - 1 && r <= D) sm[obase + (j - 1)] = v;
          sm[lbase + (j - 1)] = a_prev * rsq;  // true Ls column (dump slot when there is none)
          det.mul(p_prev * rsq);               // true L_jj
          if (!(p_prev > T(0)) && fail == 0) fail = kfail;
        }
        if (j < D) {
          const T sc = pow2_inv<T>(p);
          sp_prev = sigma * p;
          sigma *= p * sc;
          u_prev = u;
          a_prev = a;
          p_prev = p;
        }
      }
      renorm();
      if (!(has_sub && k + 1 < Tn)) sigma = T(1);
      done_out(k);
      if (lane == 0) tm_arrive(a_o_done + 8 * (uint32_t)(k & 1));
    }
    for (int64_t m = Tn; m < Tpad; ++m) done_out(m);
    if (lane == 0) {
      if (logdet) logdet[chain] = det.log_abs();
      if (info) info[chain] = fail;
    }
  }
}

}  // namespace mf
