// Chain sweep, second engine: the same Core policies and ring protocol as sweep.cuh, but the data
// moves with warp-cooperative 16-byte copies instead of one bulk copy per (stream, chain):
//
//   loads : loader warps issue cp.async.cg 16 B (SASS LDGSTS.128), consecutive lanes on consecutive
//           16-byte chunks of a chain's run, so every warp instruction reads whole 32-byte sectors
//           of at most a few chains; completion through cp.async.mbarrier.arrive.noinc
//   stores: storer warps read the output stage with LDS.128 and write STG.128 the same way
//
// A bulk copy costs the SM's TMA unit ~7.5 issue cycles whatever its size (measured), which forces
// long tiles (K >= 8..16 steps) and therefore few chains per SM; with per-lane 16-byte copies the
// tile can be short (K = 4), shared memory holds 2-4x more chains and the SM runs 2-5 compute warps
// instead of 1-2 -- the sweeps are bound by dependent-issue latency per warp, so throughput follows.
//
// Shared memory keeps the global chain-contiguous layout.  Per stage, stream i of chain c owns
// RS_i bytes, RS_i an ODD multiple of sizeof(T) (lane == chain accesses are bank-conflict free);
// the run is placed at the shift that makes shared and global addresses congruent mod 16, so the
// 16-byte chunk grid is the GLOBAL 16-byte grid for any pointer, T and D.  Chunks cut by the ends of
// the run travel element-wise.
#pragma once
#include "sweep.cuh"

namespace mf {

template <class Core, int C, int K, int NSI, int NSO, int NLW, int NSW>
struct Sweep2Cfg {
  using T = typename Core::T;
  static constexpr int ES = (int)sizeof(T);
  static constexpr int NIN = Core::NIN, NOUT = Core::NOUT;
  static constexpr int odd_es(int bytes) {  // smallest odd multiple of ES >= bytes
    int q = (bytes + ES - 1) / ES;
    if (q % 2 == 0) q += 1;
    return q * ES;
  }
  static constexpr int rs_in(int i) { return odd_es(K * Core::ein(i) * ES + 16 - ES); }
  static constexpr int rs_out(int i) { return odd_es(K * Core::eout(i) * ES + 16 - ES); }
  static constexpr int nq_in(int i) { return K * Core::ein(i) * ES / 16 + 1; }   // chunks per run
  static constexpr int nq_out(int i) { return K * Core::eout(i) * ES / 16 + 1; }
  static constexpr int off_in(int i) {
    int o = 0;
    for (int q = 0; q < i; ++q) o += C * rs_in(q);
    return o;
  }
  static constexpr int off_out(int i) {
    int o = 0;
    for (int q = 0; q < i; ++q) o += C * rs_out(q);
    return o;
  }
  static constexpr int pad16(int b) { return (b + 15) / 16 * 16; }
  static constexpr int STAGE_IN = pad16(off_in(NIN) + 16);
  static constexpr int STAGE_OUT = pad16(off_out(NOUT) + 16);
  static constexpr int NSO_EFF = NOUT > 0 ? NSO : 0;
  static constexpr int NSW_EFF = NOUT > 0 ? NSW : 0;
  static constexpr int NCW = C / 32;
  static constexpr int THREADS = 32 * (NCW + NLW + NSW_EFF);
  static constexpr int NBAR = 2 * NSI + 2 * NSO_EFF;
  static constexpr int SEG_BYTES = 16 * C * (NIN + NOUT);
  static constexpr size_t SMEM_BYTES = (size_t)STAGE_IN * NSI + (size_t)STAGE_OUT * NSO_EFF +
                                       SEG_BYTES + sizeof(uint64_t) * NBAR + 16;
  static constexpr bool align_ok() {
    for (int i = 0; i < NIN; ++i)
      if ((K * Core::ein(i) * ES) % 16 != 0) return false;
    for (int i = 0; i < NOUT; ++i)
      if ((K * Core::eout(i) * ES) % 16 != 0) return false;
    return true;
  }
  static constexpr bool FITS = align_ok() && SMEM_BYTES <= (size_t)232448 && THREADS <= 1024 &&
                               C % 32 == 0;
};

struct alignas(16) Seg2 {
  char* g;   // virtual global address of the chain's local step 0 (nullptr: nothing to move)
  int first, end;
};

// shift of a run inside its region so that shared and global addresses agree mod 16
__device__ __forceinline__ int sweep2_shift(const void* g, uint32_t region_smem_addr) {
  return (int)(((uint32_t)reinterpret_cast<uintptr_t>(g) - region_smem_addr) & 15u);
}

__device__ __forceinline__ void cp_async_16(uint32_t smem_addr, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gmem) : "memory");
}
template <int BYTES>
__device__ __forceinline__ void cp_async_small(uint32_t smem_addr, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_addr), "l"(gmem), "n"(BYTES)
               : "memory");
}

template <class Core, int C, int K, int NSI, int NSO, int NLW, int NSW>
__global__ void __launch_bounds__(Sweep2Cfg<Core, C, K, NSI, NSO, NLW, NSW>::THREADS)
chain_sweep2_kernel(const typename Core::Params prm) {
  using Cfg = Sweep2Cfg<Core, C, K, NSI, NSO, NLW, NSW>;
  using T = typename Core::T;
  constexpr int ES = Cfg::ES, NIN = Cfg::NIN, NOUT = Cfg::NOUT, NSOE = Cfg::NSO_EFF;
  static_assert(Cfg::FITS, "sweep2 configuration does not fit");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  char* in_stages = reinterpret_cast<char*>(smem_raw);
  char* out_stages = in_stages + (size_t)Cfg::STAGE_IN * NSI;
  Seg2* seg_in = reinterpret_cast<Seg2*>(out_stages + (size_t)Cfg::STAGE_OUT * NSOE);
  Seg2* seg_out = seg_in + NIN * C;
  uint64_t* bars = reinterpret_cast<uint64_t*>(seg_out + NOUT * C);
  uint64_t* full_in = bars;
  uint64_t* consumed = bars + NSI;
  uint64_t* full_out = bars + 2 * NSI;
  uint64_t* empty_out = bars + 2 * NSI + NSOE;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t nchains = Core::num_chains(prm);
  const int64_t chain0 = (int64_t)blockIdx.x * C;
  const int64_t nsteps = Core::max_steps(prm);
  const int64_t ntiles = (nsteps + K - 1) / K;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NSI; ++i) {
      mbar_init(full_in + i, 32 * NLW);
      mbar_init(consumed + i, C);
    }
    for (int i = 0; i < NSOE; ++i) {
      mbar_init(full_out + i, C);
      mbar_init(empty_out + i, 32 * Cfg::NSW_EFF);
    }
    mbar_fence_init();
  }
  // per-(stream, chain) geometry table
  for (int idx = threadIdx.x; idx < (NIN + NOUT) * C; idx += blockDim.x) {
    const bool is_in = idx < NIN * C;
    const int job = is_in ? idx : idx - NIN * C;
    const int stream = job / C, c = job % C;
    const int64_t ch = chain0 + c;
    Seg2 sg;
    sg.g = nullptr;
    sg.first = 0;
    sg.end = 0;
    if (ch < nchains) {
      const StreamGeom g = is_in ? Core::in_geom(prm, stream, ch) : Core::out_geom(prm, stream, ch);
      sg.g = g.step0;
      sg.first = (int)(g.first < 0 ? 0 : (g.first > nsteps ? nsteps : g.first));
      sg.end = (int)(g.end < 0 ? 0 : (g.end > nsteps ? nsteps : g.end));
    }
    (is_in ? seg_in : seg_out)[job] = sg;
  }
  __syncthreads();

  auto tile_id = [&](int64_t t) { return Core::BACKWARD ? ntiles - 1 - t : t; };
  const uint32_t in_base = smem_u32(in_stages), out_base = smem_u32(out_stages);

  if (warp >= Cfg::NCW && warp < Cfg::NCW + NLW) {
    // ------------------------------------ loader warps ----------------------------------------
    const int lt = (warp - Cfg::NCW) * 32 + lane;
    constexpr int NLT = 32 * NLW;
    auto issue_tile = [&](int64_t t) {
      const int si = (int)(t % NSI);
      const int64_t j0 = tile_id(t) * K;
      const uint32_t stage = in_base + (uint32_t)si * Cfg::STAGE_IN;
#pragma unroll
      for (int i = 0; i < NIN; ++i) {
        constexpr int dummy = 0;
        (void)dummy;
        const int E = Core::ein(i), NQ = Cfg::nq_in(i), RS = Cfg::rs_in(i), OFF = Cfg::off_in(i);
        const int row = E * ES;  // bytes per step
        for (int item = lt; item < C * NQ; item += NLT) {
          const int c = item / NQ, q = item - c * NQ;
          const Seg2 sg = seg_in[i * C + c];
          if (!sg.g) continue;
          int64_t a = sg.first > j0 ? sg.first - j0 : 0;
          int64_t b = sg.end - j0;
          if (b > K) b = K;
          if (b <= a) continue;
          const uint32_t region = stage + OFF + c * RS;
          const char* g0 = sg.g + j0 * (int64_t)row;            // global address of the tile's step 0
          const int sh = sweep2_shift(g0, region);               // == shift of step 0 in the region
          const int mis = (int)(reinterpret_cast<uintptr_t>(g0) & 15);
          // chunk q covers global [gal + 16q, +16), gal = g0 - mis; valid bytes g0 + [lo, hi)
          const int lo = (int)a * row, hi = (int)b * row;
          int s = 16 * q - mis, e = s + 16;                      // chunk range relative to g0
          const uint32_t sdst = region + sh;                     // shared address of g0
          if (s >= lo && e <= hi) {
            cp_async_16(sdst + s, g0 + s);
          } else {
            if (s < lo) s = lo;
            if (e > hi) e = hi;
            for (int o = s; o < e; o += ES) cp_async_small<ES>(sdst + o, g0 + o);
          }
        }
      }
      cp_async_arrive_noinc(full_in + si);
    };
    for (int64_t t = 0; t < NSI && t < ntiles; ++t) issue_tile(t);
    for (int64_t t = 0; t + NSI < ntiles; ++t) {
      mbar_wait(consumed + (int)(t % NSI), (uint32_t)((t / NSI) & 1));
      issue_tile(t + NSI);
    }
    return;
  }
  if (NOUT > 0 && warp >= Cfg::NCW + NLW) {
    // ------------------------------------ storer warps ----------------------------------------
    const int st = (warp - Cfg::NCW - NLW) * 32 + lane;
    constexpr int NST = 32 * (Cfg::NSW_EFF > 0 ? Cfg::NSW_EFF : 1);
    for (int64_t t = 0; t < ntiles; ++t) {
      const int so = (int)(t % NSO);
      mbar_wait(full_out + so, (uint32_t)((t / NSO) & 1));
      const int64_t j0 = tile_id(t) * K;
      const char* stage = out_stages + (size_t)so * Cfg::STAGE_OUT;
      const uint32_t stage_u = out_base + (uint32_t)so * Cfg::STAGE_OUT;
#pragma unroll
      for (int i = 0; i < NOUT; ++i) {
        const int E = Core::eout(i), NQ = Cfg::nq_out(i), RS = Cfg::rs_out(i), OFF = Cfg::off_out(i);
        const int row = E * ES;
        for (int item = st; item < C * NQ; item += NST) {
          const int c = item / NQ, q = item - c * NQ;
          const Seg2 sg = seg_out[i * C + c];
          if (!sg.g) continue;
          int64_t a = sg.first > j0 ? sg.first - j0 : 0;
          int64_t b = sg.end - j0;
          if (b > K) b = K;
          if (b <= a) continue;
          char* g0 = sg.g + j0 * (int64_t)row;
          const int sh = sweep2_shift(g0, stage_u + OFF + c * RS);
          const int mis = (int)(reinterpret_cast<uintptr_t>(g0) & 15);
          const int lo = (int)a * row, hi = (int)b * row;
          int s = 16 * q - mis, e = s + 16;
          const char* ssrc = stage + OFF + c * RS + sh;
          if (s >= lo && e <= hi) {
            const int4 v = *reinterpret_cast<const int4*>(ssrc + s);
            __stcs(reinterpret_cast<int4*>(g0 + s), v);
          } else {
            if (s < lo) s = lo;
            if (e > hi) e = hi;
            for (int o = s; o < e; o += ES)
              *reinterpret_cast<T*>(g0 + o) = *reinterpret_cast<const T*>(ssrc + o);
          }
        }
      }
      mbar_arrive(empty_out + so);
    }
    return;
  }
  if (warp >= Cfg::NCW) return;

  // --------------------------------- compute threads ------------------------------------------
  const int c = warp * 32 + lane;
  const int64_t chain = chain0 + c;
  const bool valid = chain < nchains;
  int in_off[NIN > 0 ? NIN : 1], out_off[NOUT > 0 ? NOUT : 1];
#pragma unroll
  for (int i = 0; i < NIN; ++i) {
    const uint32_t region = (uint32_t)(Cfg::off_in(i) + c * Cfg::rs_in(i));
    // stage bases are multiples of 16, so the shift is the same in every stage
    in_off[i] = (int)region + sweep2_shift(seg_in[i * C + c].g, in_base + region);
  }
#pragma unroll
  for (int i = 0; i < NOUT; ++i) {
    const uint32_t region = (uint32_t)(Cfg::off_out(i) + c * Cfg::rs_out(i));
    out_off[i] = (int)region + sweep2_shift(seg_out[i * C + c].g, out_base + region);
  }
  Core core;
  if (valid) core.init(prm, chain);
  for (int64_t t = 0; t < ntiles; ++t) {
    const int si = (int)(t % NSI);
    mbar_wait(full_in + si, (uint32_t)((t / NSI) & 1));
    const char* ist = in_stages + (size_t)si * Cfg::STAGE_IN;
    char* ost = nullptr;
    int so = 0;
    if (NOUT > 0) {
      so = (int)(t % NSO);
      mbar_wait(empty_out + so, (uint32_t)(((t / NSO) & 1) ^ 1));
      ost = out_stages + (size_t)so * Cfg::STAGE_OUT;
    }
    const int64_t j0 = tile_id(t) * K;
    const int ns = (int)((nsteps - j0 < K) ? (nsteps - j0) : K);
    if (valid) {
      const T* in[NIN > 0 ? NIN : 1];
      T* out[NOUT > 0 ? NOUT : 1];
#pragma unroll
      for (int i = 0; i < NIN; ++i) in[i] = reinterpret_cast<const T*>(ist + in_off[i]);
#pragma unroll
      for (int i = 0; i < NOUT; ++i) out[i] = reinterpret_cast<T*>(ost + out_off[i]);
      core.tile(prm, in, out, j0, ns);
    }
    mbar_arrive(consumed + si);
    if (NOUT > 0) mbar_arrive(full_out + so);
  }
  core.finish(prm, chain, valid);
}

template <class Core, int C, int K, int NSI, int NSO, int NLW, int NSW>
inline cudaError_t launch_chain_sweep2(const typename Core::Params& prm, int64_t nchains,
                                       cudaStream_t s) {
  using Cfg = Sweep2Cfg<Core, C, K, NSI, NSO, NLW, NSW>;
  auto kern = chain_sweep2_kernel<Core, C, K, NSI, NSO, NLW, NSW>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const unsigned grid = (unsigned)((nchains + C - 1) / C);
  kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, s>>>(prm);
  return cudaGetLastError();
}

}  // namespace mf
