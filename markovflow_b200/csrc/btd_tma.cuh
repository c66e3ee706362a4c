// TMA-fed Cholesky(+solve) sweep: the fastest path of mf_btd_cholesky.
//
// CTA = 1 compute warp (lane == chain, blocks in registers) + 1 producer warp (lane == chain)
// that moves each chain's next K steps with 1-D bulk async copies (cp.async.bulk, the TMA engine):
// one instruction per chain, stream and tile instead of hundreds of LSU element copies, so the
// LSU / shared-memory pipe serves only the compute warp.  Shared memory keeps the GLOBAL layout
// (chain-contiguous); each chain owns a region per stream whose stride is an odd multiple of 16
// bytes (spreads lanes over banks) and whose data start is shifted by the chain's global
// misalignment a0 = addr & 15, so that every bulk copy has 16-byte aligned source, destination
// and size; the at most (16/sizeof(T) - 1) leading / trailing elements of a misaligned segment
// travel as element-sized cp.async (loads) or plain stores (write-back).
//
//   ring: NSI input stages, NSO output stages
//   full_in[NSI]  : 32 x mbarrier.arrive.expect_tx (bulk bytes) + 32 x cp.async arrive (elements)
//   full_out[NSO] : compute lanes finished a tile (also releases that tile's input stage)
//   empty_out[NSO]: producer lanes' bulk stores have finished reading the stage
#pragma once
#include "chol_core.cuh"
#include "pipe.cuh"

namespace mf {

template <typename T, int D, bool RHS, int C, int K, int NSI, int NSO>
struct CholTmaCfg {
  static constexpr int DD = D * D;
  static constexpr int ES = (int)sizeof(T);
  static constexpr int odd16(int bytes) {
    int q = (bytes + 15) / 16;
    if (q % 2 == 0) q += 1;
    return q * 16;
  }
  static constexpr int SEG_M = K * DD * ES;  // bytes of one chain's tile segment, matrix streams
  static constexpr int SEG_V = K * D * ES;
  static constexpr int RS_M = odd16(SEG_M + 16);  // region stride (room for the a0 shift)
  static constexpr int RS_V = odd16(SEG_V + 16);
  static constexpr int STAGE_BYTES = C * (2 * RS_M + (RHS ? RS_V : 0));
  static constexpr int NSTREAM = 2 + (RHS ? 1 : 0);
  static constexpr int NJ = NSTREAM * C;          // (stream, chain) jobs per direction
  static constexpr int LPW = 32;                  // producer lanes used per warp
  static constexpr int NPW = (2 * NJ + LPW - 1) / LPW;  // producer warps: NJ loaders + NJ storers
  // warp w runs on SM sub-partition w % 4: warp 4 would share sub-partition 0 with the compute warp
  // (warp 0) and steal its issue slots with the UBLKCP loops, so it is left idle when NPW > 3
  static constexpr int SKIP = NPW > 3 ? 1 : 0;
  static constexpr int THREADS = 32 * (1 + NPW + SKIP);
  static constexpr size_t SMEM_BYTES =
      (size_t)STAGE_BYTES * (NSI + NSO) + sizeof(uint64_t) * (NSI + 2 * NSO) + 16;
  static constexpr bool ALIGN_OK = (SEG_M % 16 == 0) && (SEG_V % 16 == 0);
};

// chain-contiguous regions: element (s, e) of a stream at  s*E + e
template <int D>
struct ContiguousLayout {
  static constexpr int SS_M = D * D, ES_M = 1, SS_V = D, ES_V = 1;
};

// One chain's view of one stream (global side) for the producer.
struct ChainSeg {
  char* g;        // global byte address of the chain's step-0 record (nullptr: lane has no chain)
  int a0;         // g & 15
  int64_t len;    // steps in the stream
};

template <int ES, int E, int K>
__device__ __forceinline__ uint32_t seg_interior_bytes(const ChainSeg& sg, int64_t k0, int& head,
                                                       int& bytes) {
  const int64_t left = sg.len - k0;
  const int nst = (int)(left < K ? (left < 0 ? 0 : left) : K);
  bytes = nst * E * ES;
  head = (16 - sg.a0) & 15;
  if (head > bytes) head = bytes;
  return (uint32_t)((bytes - head) & ~15);
}

template <int ES, int E, int K>
__device__ __forceinline__ void seg_load(char* region, const ChainSeg& sg, int64_t k0,
                                         uint64_t* bar) {
  if (!sg.g) return;
  int head, bytes;
  const uint32_t interior = seg_interior_bytes<ES, E, K>(sg, k0, head, bytes);
  const char* g0 = sg.g + k0 * (int64_t)(E * ES);
  char* sd = region + sg.a0;
  if (interior) tma_load_1d(sd + head, g0 + head, interior, bar);
  for (int o = 0; o < head; o += ES) cp_async_elem<ES>(sd + o, g0 + o);
  for (int o = head + (int)interior; o < bytes; o += ES) cp_async_elem<ES>(sd + o, g0 + o);
}

template <typename T, int E, int K>
__device__ __forceinline__ void seg_store(const char* region, const ChainSeg& sg, int64_t k0) {
  constexpr int ES = (int)sizeof(T);
  if (!sg.g) return;
  int head, bytes;
  const uint32_t interior = seg_interior_bytes<ES, E, K>(sg, k0, head, bytes);
  char* g0 = sg.g + k0 * (int64_t)(E * ES);
  const char* sd = region + sg.a0;
  if (interior) tma_store_1d(g0 + head, sd + head, interior);
  for (int o = 0; o < head; o += ES)
    *reinterpret_cast<T*>(g0 + o) = *reinterpret_cast<const T*>(sd + o);
  for (int o = head + (int)interior; o < bytes; o += ES)
    *reinterpret_cast<T*>(g0 + o) = *reinterpret_cast<const T*>(sd + o);
}

template <typename T, int D, bool RHS, int C, int K, int NSI, int NSO>
__global__ void __launch_bounds__(CholTmaCfg<T, D, RHS, C, K, NSI, NSO>::THREADS, 1)
btd_chol_tma_kernel(const T* __restrict__ diag, const T* __restrict__ sub,
                    const T* __restrict__ rhs, T* od, T* os, T* ox, T* __restrict__ logdet,
                    int32_t* __restrict__ info, int64_t B, int64_t Tn, const int elem_wait) {
  using Cfg = CholTmaCfg<T, D, RHS, C, K, NSI, NSO>;
  constexpr int DD = Cfg::DD, ES = Cfg::ES;
  static_assert(Cfg::ALIGN_OK, "K * E * sizeof(T) must be a multiple of 16 for every stream");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  char* in_stages = reinterpret_cast<char*>(smem_raw);
  char* out_stages = in_stages + (size_t)Cfg::STAGE_BYTES * NSI;
  uint64_t* bars = reinterpret_cast<uint64_t*>(out_stages + (size_t)Cfg::STAGE_BYTES * NSO);
  uint64_t* full_in = bars;
  uint64_t* full_out = bars + NSI;
  uint64_t* empty_out = bars + NSI + NSO;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t chain0 = (int64_t)blockIdx.x * C;
  const int64_t chain = chain0 + lane;
  const bool valid = (lane < C) && (chain < B);
  const int64_t ntiles = (Tn + K - 1) / K;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NSI; ++i) mbar_init(full_in + i, 2 * Cfg::NJ);
    for (int i = 0; i < NSO; ++i) {
      mbar_init(full_out + i, 32);
      mbar_init(empty_out + i, Cfg::NJ);
    }
    mbar_fence_init();
  }
  __syncthreads();

  // region offsets inside a stage: [diag: C x RS_M][sub: C x RS_M][vec: C x RS_V]
  auto region_off = [&](int stream, int c) {
    return stream == 0 ? c * Cfg::RS_M
                       : (stream == 1 ? C * Cfg::RS_M + c * Cfg::RS_M
                                      : 2 * C * Cfg::RS_M + c * Cfg::RS_V);
  };
  auto seg_of = [&](const T* base, int64_t len, int E, int64_t ch) {
    ChainSeg sg;
    sg.g = (ch < B && base) ? (char*)const_cast<T*>(base) + ch * len * (int64_t)(E * ES) : nullptr;
    sg.a0 = (int)(reinterpret_cast<uintptr_t>(sg.g) & 15);
    sg.len = len;
    return sg;
  };

  if (warp >= 1) {
    // ---- producer threads: one per (direction, stream, chain) job; each issues ONE bulk copy
    //      per tile, so the per-lane UBLKCP issue cost is spread over NPW warps ----------------
    if (Cfg::SKIP && warp == 4) return;
    const int p = (warp - 1 - ((Cfg::SKIP && warp > 4) ? 1 : 0)) * Cfg::LPW + lane;
    if (lane >= Cfg::LPW || p >= 2 * Cfg::NJ) return;
    const bool storer = p >= Cfg::NJ;
    const int job = storer ? p - Cfg::NJ : p;
    const int stream = job / C, c = job % C;
    const int64_t ch = chain0 + c;
    const int roff = region_off(stream, c);
    if (!storer) {
      const ChainSeg sg = stream == 0 ? seg_of(diag, Tn, DD, ch)
                                      : (stream == 1 ? seg_of(sub, Tn - 1, DD, ch) : seg_of(rhs, Tn, D, ch));
      auto issue_load = [&](int64_t tile) {
        const int si = (int)(tile % NSI);
        char* st = in_stages + (size_t)si * Cfg::STAGE_BYTES + roff;
        uint64_t* bar = full_in + si;
        const int64_t k0 = tile * K;
        int h, b;
        uint32_t tx = 0;
        if (sg.g) tx = (stream == 2) ? seg_interior_bytes<ES, D, K>(sg, k0, h, b)
                                     : seg_interior_bytes<ES, DD, K>(sg, k0, h, b);
        mbar_arrive_expect_tx(bar, tx);
        if (stream == 2) seg_load<ES, D, K>(st, sg, k0, bar); else seg_load<ES, DD, K>(st, sg, k0, bar);
        cp_async_arrive(bar, elem_wait);
      };
      for (int64_t t = 0; t < NSI && t < ntiles; ++t) issue_load(t);
      for (int64_t t = 0; t + NSI < ntiles; ++t) {
        mbar_wait(full_out + (int)(t % NSO), (uint32_t)((t / NSO) & 1));  // tile t consumed
        issue_load(t + NSI);
      }
    } else {
      const ChainSeg sg = stream == 0 ? seg_of(od, Tn, DD, ch)
                                      : (stream == 1 ? seg_of(os, Tn - 1, DD, ch) : seg_of(ox, Tn, D, ch));
      for (int64_t t = 0; t < ntiles; ++t) {
        const int so = (int)(t % NSO);
        mbar_wait(full_out + so, (uint32_t)((t / NSO) & 1));
        const char* st = out_stages + (size_t)so * Cfg::STAGE_BYTES + roff;
        if (stream == 2) seg_store<T, D, K>(st, sg, t * K); else seg_store<T, DD, K>(st, sg, t * K);
        tma_store_commit();
        tma_store_wait_read<0>();
        mbar_arrive(empty_out + so);
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    return;
  }

  // compute lane geometry: misalignment shifts of its own chain's streams
  const ChainSeg i_diag = seg_of(diag, Tn, DD, chain), i_sub = seg_of(sub, Tn - 1, DD, chain),
                 i_rhs = seg_of(rhs, Tn, D, chain);
  const ChainSeg o_diag = seg_of(od, Tn, DD, chain), o_sub = seg_of(os, Tn - 1, DD, chain),
                 o_x = seg_of(ox, Tn, D, chain);
  const int off_m0 = region_off(0, lane), off_m1 = region_off(1, lane), off_v = region_off(2, lane);

  // --------------------------------- compute warp ------------------------------------------
  using Layout = ContiguousLayout<D>;
  CholCore<T, D, RHS, Layout> core;
  core.init();
  for (int64_t t = 0; t < ntiles; ++t) {
    const int si = (int)(t % NSI), so = (int)(t % NSO);
    mbar_wait(full_in + si, (uint32_t)((t / NSI) & 1));
    mbar_wait(empty_out + so, (uint32_t)(((t / NSO) & 1) ^ 1));
    const char* ist = in_stages + (size_t)si * Cfg::STAGE_BYTES;
    char* ost = out_stages + (size_t)so * Cfg::STAGE_BYTES;
    const int64_t k0 = t * K;
    const int ns = (int)((Tn - k0 < K) ? (Tn - k0) : K);
    if (valid)
      core.tile(reinterpret_cast<const T*>(ist + off_m0 + i_diag.a0),
                reinterpret_cast<const T*>(ist + off_m1 + i_sub.a0),
                reinterpret_cast<const T*>(ist + off_v + i_rhs.a0),
                reinterpret_cast<T*>(ost + off_m0 + o_diag.a0),
                reinterpret_cast<T*>(ost + off_m1 + o_sub.a0),
                reinterpret_cast<T*>(ost + off_v + o_x.a0), ns, k0, Tn, logdet != nullptr);
    fence_proxy_async_smem();  // our shared-memory writes -> visible to the TMA store
    mbar_arrive(full_out + so);
  }
  if (valid) {
    if (logdet) logdet[chain] = core.log_det();
    if (info) info[chain] = core.fail;
  }
}

}  // namespace mf
