// Staged (shared-memory pipelined) Cholesky(+solve) sweep: the fast path of mf_btd_cholesky.
//
// CTA = 1 compute warp (lane == chain, blocks in registers) + NCW copy warps that stream each
// chain's next K steps through a transposed shared-memory ring (see pipe.cuh).  Per step the
// compute warp touches only shared memory and registers, so HBM latency never sits on the
// sequential Cholesky critical path; HBM sees only full 256-byte coalesced segments.
//
//   ring: NSI input stages (diag | sub | rhs per step), NSO output stages (Ld | Ls | x per step)
//   barriers: full_in[NSI]  <- cp.async completion of all copy threads
//             full_out[NSO] <- compute warp finished a tile (also frees that tile's input stage)
//             empty_out[NSO]<- copy threads drained an output stage
#pragma once
#include "chol_core.cuh"
#include "pipe.cuh"

namespace mf {

// transposed tile: element (s, e) of a stream at  s*ETOT*CP + e*CP  (+ lane)
template <int ETOT, int CP>
struct TransposedLayout {
  static constexpr int SS_M = ETOT * CP, ES_M = CP, SS_V = ETOT * CP, ES_V = CP;
};

template <typename T, int D, bool RHS, int C, int K, int NSI, int NSO>
struct CholStagedCfg {
  static constexpr int DD = D * D;
  static constexpr int ETOT = 2 * DD + (RHS ? D : 0);
  static constexpr int CP = C + 1;
  static constexpr int STAGE_ELEMS = K * ETOT * CP;
  static constexpr int NCW = 6;  // copy warps: warps 1,2,3,5,6,7 (warp 4 would share the compute
                                 // warp's scheduler, so it idles)
  static constexpr int THREADS = 256;
  static constexpr int NTAB = 3;  // chain-offset tables: diag-like, sub-like, vector-like streams
  static constexpr size_t SMEM_BYTES = sizeof(T) * (size_t)STAGE_ELEMS * (NSI + NSO) +
                                       sizeof(uint64_t) * (NSI + 2 * NSO) +
                                       sizeof(int64_t) * NTAB * C + 16;
};

template <typename T, int D, bool RHS, int C, int K, int NSI, int NSO>
__global__ void __launch_bounds__(256, 1)
btd_chol_staged_kernel(const T* __restrict__ diag, const T* __restrict__ sub,
                       const T* __restrict__ rhs, T* od, T* os, T* ox, T* __restrict__ logdet,
                       int32_t* __restrict__ info, int64_t B, int64_t Tn, const int elem_wait) {
  using Cfg = CholStagedCfg<T, D, RHS, C, K, NSI, NSO>;
  constexpr int DD = Cfg::DD, ETOT = Cfg::ETOT, CP = Cfg::CP, NCW = Cfg::NCW;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* in_tiles = reinterpret_cast<T*>(smem_raw);
  T* out_tiles = in_tiles + (size_t)Cfg::STAGE_ELEMS * NSI;
  uint64_t* bars = reinterpret_cast<uint64_t*>(
      (reinterpret_cast<uintptr_t>(out_tiles + (size_t)Cfg::STAGE_ELEMS * NSO) + 7) & ~uintptr_t(7));
  uint64_t* full_in = bars;
  uint64_t* full_out = bars + NSI;
  uint64_t* empty_out = bars + NSI + NSO;
  int64_t* off_diag = reinterpret_cast<int64_t*>(bars + NSI + 2 * NSO);
  int64_t* off_sub = off_diag + C;
  int64_t* off_vec = off_sub + C;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t chain0 = (int64_t)blockIdx.x * C;
  const int64_t ntiles = (Tn + K - 1) / K;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NSI; ++i) mbar_init(full_in + i, 32 * NCW);
    for (int i = 0; i < NSO; ++i) { mbar_init(full_out + i, 32); mbar_init(empty_out + i, 32 * NCW); }
    mbar_fence_init();
  }
  {
    // inputs and outputs share chain geometry here (no sample broadcast in the factorisation)
    Stream<const T> g_diag{diag, Tn, B}, g_sub{sub, Tn - 1, B}, g_vec{rhs, Tn, B};
    fill_chain_offsets<const T, DD, C>(off_diag, g_diag, chain0, B, threadIdx.x, blockDim.x);
    fill_chain_offsets<const T, DD, C>(off_sub, g_sub, chain0, B, threadIdx.x, blockDim.x);
    fill_chain_offsets<const T, D, C>(off_vec, g_vec, chain0, B, threadIdx.x, blockDim.x);
  }
  __syncthreads();

  if (warp == 4) return;
  if (warp > 0) {
    // ------------------------------- copy warps ------------------------------------------
    const int w = warp < 4 ? warp - 1 : warp - 2;
    auto issue_loads = [&](int64_t tile) {
      T* st = in_tiles + (size_t)(tile % NSI) * Cfg::STAGE_ELEMS;
      const int64_t k0 = tile * K;
      tile_load<T, DD, 0, ETOT, K, C, CP>(st, diag, off_diag, Tn, k0, w, NCW, lane);
      tile_load<T, DD, DD, ETOT, K, C, CP>(st, sub, off_sub, Tn - 1, k0, w, NCW, lane);
      if (RHS) tile_load<T, D, 2 * DD, ETOT, K, C, CP>(st, rhs, off_vec, Tn, k0, w, NCW, lane);
      cp_async_arrive(full_in + (tile % NSI), elem_wait);
    };
    for (int64_t t = 0; t < NSI && t < ntiles; ++t) issue_loads(t);
    for (int64_t t = 0; t < ntiles; ++t) {
      const int so = (int)(t % NSO);
      mbar_wait(full_out + so, (uint32_t)((t / NSO) & 1));
      if (t + NSI < ntiles) issue_loads(t + NSI);  // tile t's input stage is free now
      const T* st = out_tiles + (size_t)so * Cfg::STAGE_ELEMS;
      const int64_t k0 = t * K;
      tile_store<T, DD, 0, ETOT, K, C, CP>(st, od, off_diag, Tn, k0, w, NCW, lane);
      tile_store<T, DD, DD, ETOT, K, C, CP>(st, os, off_sub, Tn - 1, k0, w, NCW, lane);
      if (RHS) tile_store<T, D, 2 * DD, ETOT, K, C, CP>(st, ox, off_vec, Tn, k0, w, NCW, lane);
      mbar_arrive(empty_out + so);
    }
    return;
  }

  // --------------------------------- compute warp ------------------------------------------
  using Layout = TransposedLayout<ETOT, CP>;
  const bool valid = (lane < C) && (chain0 + lane < B);
  CholCore<T, D, RHS, Layout> core;
  core.init();
  for (int64_t t = 0; t < ntiles; ++t) {
    const int si = (int)(t % NSI), so = (int)(t % NSO);
    mbar_wait(full_in + si, (uint32_t)((t / NSI) & 1));
    mbar_wait(empty_out + so, (uint32_t)(((t / NSO) & 1) ^ 1));
    const T* ip = in_tiles + (size_t)si * Cfg::STAGE_ELEMS + lane;
    T* op = out_tiles + (size_t)so * Cfg::STAGE_ELEMS + lane;
    const int64_t k0 = t * K;
    const int ns = (int)((Tn - k0 < K) ? (Tn - k0) : K);
    if (valid)
      core.tile(ip, ip + DD * CP, ip + 2 * DD * CP, op, op + DD * CP, op + 2 * DD * CP, ns, k0, Tn,
                logdet != nullptr);
    mbar_arrive(full_out + so);
  }
  if (valid) {
    if (logdet) logdet[chain0 + lane] = core.log_det();
    if (info) info[chain0 + lane] = core.fail;
  }
}

}  // namespace mf
