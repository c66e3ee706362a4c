// Generic TMA-fed chain sweep: the streaming engine behind the sequential (per-chain) kernels.
//
// Every recurrence of the hot path has the same shape: a thread owns one chain, walks its T steps
// in order (forwards or backwards), reads a few small records per step (blocks / vectors of
// chain-contiguous [B,T,...] arrays) and optionally writes a few.  Reading them straight from
// global memory puts HBM latency on the sequential critical path and wastes sectors.  Here
//
//   CTA = NCW compute warps (lane == chain, recurrence state in registers)
//       + producer threads, one per (stream, chain) job, each issuing ONE 1-D bulk copy
//         (cp.async.bulk, the TMA engine; SASS UBLKCP) per tile of K steps.
//
// Shared memory keeps the global chain-contiguous layout: per stage, stream i of chain c owns a
// region of RS_i bytes (an odd multiple of 16, so lanes spread over banks) in which the tile's data
// sits at the chain's global misalignment a0 = addr & 15 -- every bulk copy is then 16-byte aligned
// on both sides for ANY pointer, T and D; the (< 16 B) ragged head/tail of a misaligned segment
// travels as element cp.async (loads) or plain stores (write-back).
//
//   ring: NSI input stages, NSO output stages
//   full_in[NSI]   loaders: arrive.expect_tx (bulk bytes) + cp.async arrive (elements)
//   consumed[NSI]  compute threads are done with an input stage
//   full_out[NSO]  compute threads filled an output stage
//   empty_out[NSO] storers' bulk stores have finished reading an output stage
//
// A Core policy supplies the streams and the per-tile arithmetic:
//   typename Core::T, Core::Params (POD, passed by value)
//   static constexpr int NIN, NOUT;  static constexpr int ein(i), eout(i)   elements per step
//   static constexpr bool BACKWARD                                          sweep direction
//   static __device__ int64_t num_chains(const Params&), max_steps(const Params&)
//   static __device__ StreamGeom in_geom(const Params&, int i, int64_t chain), out_geom(...)
//   __device__ void init(const Params&, int64_t chain)
//   __device__ void tile(const Params&, const T* const* in, T* const* out, int64_t j0, int ns)
//       in[i] / out[i] point at local step j0 of stream i (entries outside the stream's valid
//       range hold garbage and must not be used); process steps j0..j0+ns-1 (descending if BACKWARD)
//   __device__ void finish(const Params&, int64_t chain, bool valid)   called by ALL compute threads
#pragma once
#include "dispatch.cuh"
#include "pipe.cuh"

namespace mf {

// Where stream entries of one chain live in global memory: entry of local step j is at
// step0 + j * E * sizeof(T) and exists for first <= j < end (step0 itself may lie outside the array).
struct StreamGeom {
  char* step0;
  int64_t first, end;
};

template <class Core, int C, int K, int NSI, int NSO>
struct SweepCfg {
  using T = typename Core::T;
  static constexpr int ES = (int)sizeof(T);
  static constexpr int NIN = Core::NIN, NOUT = Core::NOUT;
  static constexpr int odd16(int bytes) {
    int q = (bytes + 15) / 16;
    if (q % 2 == 0) q += 1;
    return q * 16;
  }
  static constexpr int rs_in(int i) { return odd16(K * Core::ein(i) * ES + 16); }
  static constexpr int rs_out(int i) { return odd16(K * Core::eout(i) * ES + 16); }
  static constexpr int off_in(int i) {  // byte offset of stream i's regions inside an input stage
    int o = 0;
    for (int q = 0; q < i; ++q) o += C * rs_in(q);
    return o;
  }
  static constexpr int off_out(int i) {
    int o = 0;
    for (int q = 0; q < i; ++q) o += C * rs_out(q);
    return o;
  }
  static constexpr int STAGE_IN = off_in(NIN);
  static constexpr int STAGE_OUT = off_out(NOUT);
  static constexpr int NSO_EFF = NOUT > 0 ? NSO : 0;
  // NSO == 0 with outputs: DIRECT mode -- the compute threads store their results straight to
  // global memory (no output stages, no storer threads); one dump tile per stream absorbs the
  // writes of streams that have no destination (optional outputs)
  static constexpr bool DIRECT = NOUT > 0 && NSO == 0;
  static constexpr int dump_bytes() {
    int o = 0;
    for (int i = 0; i < NOUT; ++i) o += (K * Core::eout(i) * ES + 15) / 16 * 16;
    return DIRECT ? o : 0;
  }
  static constexpr int dump_off(int i) {
    int o = 0;
    for (int q = 0; q < i; ++q) o += (K * Core::eout(q) * ES + 15) / 16 * 16;
    return o;
  }
  static constexpr int NJ_IN = NIN * C, NJ_OUT = DIRECT ? 0 : NOUT * C;
  static constexpr int NCW = C / 32;                               // compute warps
  static constexpr int NPW = (NJ_IN + NJ_OUT + 31) / 32;           // producer warps
  static constexpr int THREADS = 32 * (NCW + NPW);
  static constexpr int NBAR = 2 * NSI + 2 * NSO_EFF;
  static constexpr size_t SMEM_BYTES =
      (size_t)STAGE_IN * NSI + (size_t)STAGE_OUT * NSO_EFF + dump_bytes() + sizeof(uint64_t) * NBAR + 16;
  static constexpr bool align_ok() {
    for (int i = 0; i < NIN; ++i)
      if ((K * Core::ein(i) * ES) % 16 != 0) return false;
    for (int i = 0; i < NOUT; ++i)
      if (!DIRECT && (K * Core::eout(i) * ES) % 16 != 0) return false;
    return true;
  }
  static constexpr bool FITS = align_ok() && SMEM_BYTES <= (size_t)232448 && THREADS <= 1024 &&
                               C % 32 == 0;
};

struct SweepSeg {
  char* g;     // virtual global address of the chain's local step 0 (nullptr: no such chain/stream)
  int a0;      // g & 15
  int64_t first, end;
};

__device__ __forceinline__ SweepSeg make_seg(const StreamGeom& sg, bool valid) {
  SweepSeg s;
  s.g = valid ? sg.step0 : nullptr;
  s.a0 = (int)(reinterpret_cast<uintptr_t>(s.g) & 15);
  s.first = sg.first;
  s.end = sg.end;
  return s;
}

// Byte range [lo, hi) (relative to the tile's first step) of the valid entries of tile [j0, j0+K),
// split into an unaligned head [lo, lo+head), a 16-byte aligned interior and a tail.
template <int ES, int K>
__device__ __forceinline__ uint32_t sweep_ranges(const SweepSeg& sg, int E, int64_t j0, int& lo,
                                                 int& hi, int& head) {
  int64_t a = sg.first > j0 ? sg.first - j0 : 0;
  int64_t b = sg.end - j0;
  if (b > K) b = K;
  if (b < a) b = a;
  lo = (int)a * E * ES;
  hi = (int)b * E * ES;
  head = (16 - ((sg.a0 + lo) & 15)) & 15;
  if (head > hi - lo) head = hi - lo;
  return (uint32_t)((hi - lo - head) & ~15);
}

template <class Core, int C, int K, int NSI, int NSO>
__global__ void __launch_bounds__(SweepCfg<Core, C, K, NSI, NSO>::THREADS)
chain_sweep_kernel(const typename Core::Params prm, const int cpb, const int elem_wait) {
  using Cfg = SweepCfg<Core, C, K, NSI, NSO>;
  using T = typename Core::T;
  constexpr int ES = Cfg::ES, NIN = Cfg::NIN, NOUT = Cfg::NOUT, NSOE = Cfg::NSO_EFF;
  static_assert(Cfg::FITS, "sweep configuration does not fit (alignment / shared memory / threads)");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  char* in_stages = reinterpret_cast<char*>(smem_raw);
  char* out_stages = in_stages + (size_t)Cfg::STAGE_IN * NSI;
  char* dump = out_stages + (size_t)Cfg::STAGE_OUT * NSOE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(dump + Cfg::dump_bytes());
  uint64_t* full_in = bars;
  uint64_t* consumed = bars + NSI;
  uint64_t* full_out = bars + 2 * NSI;
  uint64_t* empty_out = bars + 2 * NSI + NSOE;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t nchains = Core::num_chains(prm);
  // cpb <= C chains per CTA: few chains are spread over all SMs (fewer bulk copies per CTA and tile)
  const int64_t chain0 = (int64_t)blockIdx.x * cpb;
  const int64_t nsteps = Core::max_steps(prm);
  const int64_t ntiles = (nsteps + K - 1) / K;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NSI; ++i) {
      mbar_init(full_in + i, 2 * Cfg::NJ_IN);
      mbar_init(consumed + i, C);
    }
    for (int i = 0; i < NSOE; ++i) {
      mbar_init(full_out + i, C);
      mbar_init(empty_out + i, Cfg::NJ_OUT);
    }
    mbar_fence_init();
  }
  __syncthreads();

  auto tile_id = [&](int64_t t) { return Core::BACKWARD ? ntiles - 1 - t : t; };

  if (warp >= Cfg::NCW) {
    // ------------------------------- producer threads ---------------------------------------
    const int p = (warp - Cfg::NCW) * 32 + lane;
    if (p >= Cfg::NJ_IN + Cfg::NJ_OUT) return;
    if (p < Cfg::NJ_IN) {
      const int stream = p / C, c = p % C;
      const int64_t ch = chain0 + c;
      const int E = Core::ein(stream);
      int roff = 0;
#pragma unroll
      for (int q = 0; q < NIN; ++q)
        if (q == stream) roff = Cfg::off_in(q) + c * Cfg::rs_in(q);
      const bool valid = c < cpb && ch < nchains;
      const SweepSeg sg = make_seg(valid ? Core::in_geom(prm, stream, ch) : StreamGeom{nullptr, 0, 0}, valid);
      auto issue_load = [&](int64_t t) {
        const int si = (int)(t % NSI);
        uint64_t* bar = full_in + si;
        const int64_t j0 = tile_id(t) * K;
        uint32_t tx = 0;
        int lo = 0, hi = 0, head = 0;
        if (sg.g) tx = sweep_ranges<ES, K>(sg, E, j0, lo, hi, head);
        mbar_arrive_expect_tx(bar, tx);
        if (sg.g && hi > lo) {
          char* sd = in_stages + (size_t)si * Cfg::STAGE_IN + roff + sg.a0;
          const char* g0 = sg.g + j0 * (int64_t)(E * ES);
          if (tx) tma_load_1d(sd + lo + head, g0 + lo + head, tx, bar);
          for (int o = lo; o < lo + head; o += ES) cp_async_elem<ES>(sd + o, g0 + o);
          for (int o = lo + head + (int)tx; o < hi; o += ES) cp_async_elem<ES>(sd + o, g0 + o);
        }
        cp_async_arrive(bar, elem_wait);
      };
      for (int64_t t = 0; t < NSI && t < ntiles; ++t) issue_load(t);
      for (int64_t t = 0; t + NSI < ntiles; ++t) {
        mbar_wait(consumed + (int)(t % NSI), (uint32_t)((t / NSI) & 1));  // tile t consumed
        issue_load(t + NSI);
      }
    } else if (NOUT > 0 && !Cfg::DIRECT) {
      const int job = p - Cfg::NJ_IN;
      const int stream = job / C, c = job % C;
      const int64_t ch = chain0 + c;
      const int E = Core::eout(stream);
      int roff = 0;
#pragma unroll
      for (int q = 0; q < NOUT; ++q)
        if (q == stream) roff = Cfg::off_out(q) + c * Cfg::rs_out(q);
      const bool valid = c < cpb && ch < nchains;
      const SweepSeg sg = make_seg(valid ? Core::out_geom(prm, stream, ch) : StreamGeom{nullptr, 0, 0}, valid);
      for (int64_t t = 0; t < ntiles; ++t) {
        const int so = (int)(t % (NSO > 0 ? NSO : 1));
        mbar_wait(full_out + so, (uint32_t)((t / (NSO > 0 ? NSO : 1)) & 1));
        if (sg.g) {
          const int64_t j0 = tile_id(t) * K;
          int lo, hi, head;
          const uint32_t tx = sweep_ranges<ES, K>(sg, E, j0, lo, hi, head);
          if (hi > lo) {
            const char* sd = out_stages + (size_t)so * Cfg::STAGE_OUT + roff + sg.a0;
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
    }
    return;
  }

  // --------------------------------- compute threads ------------------------------------------
  const int c = warp * 32 + lane;
  const int64_t chain = chain0 + c;
  const bool valid = c < cpb && chain < nchains;
  int in_off[NIN > 0 ? NIN : 1], out_off[NOUT > 0 ? NOUT : 1];
#pragma unroll
  for (int i = 0; i < NIN; ++i) {
    const StreamGeom g = valid ? Core::in_geom(prm, i, chain) : StreamGeom{nullptr, 0, 0};
    in_off[i] = Cfg::off_in(i) + c * Cfg::rs_in(i) + (int)(reinterpret_cast<uintptr_t>(g.step0) & 15);
  }
  char* out_g[NOUT > 0 ? NOUT : 1];  // DIRECT: global address of the chain's local step 0
#pragma unroll
  for (int i = 0; i < NOUT; ++i) {
    const StreamGeom g = valid ? Core::out_geom(prm, i, chain) : StreamGeom{nullptr, 0, 0};
    out_off[i] = Cfg::off_out(i) + c * Cfg::rs_out(i) + (int)(reinterpret_cast<uintptr_t>(g.step0) & 15);
    out_g[i] = g.step0;
  }
  Core core;
  if (valid) core.init(prm, chain);
  for (int64_t t = 0; t < ntiles; ++t) {
    const int si = (int)(t % NSI);
    mbar_wait(full_in + si, (uint32_t)((t / NSI) & 1));
    const char* ist = in_stages + (size_t)si * Cfg::STAGE_IN;
    char* ost = nullptr;
    int so = 0;
    if (NOUT > 0 && !Cfg::DIRECT) {
      so = (int)(t % (NSO > 0 ? NSO : 1));
      mbar_wait(empty_out + so, (uint32_t)(((t / (NSO > 0 ? NSO : 1)) & 1) ^ 1));
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
      for (int i = 0; i < NOUT; ++i) {
        if (Cfg::DIRECT)
          out[i] = out_g[i] ? reinterpret_cast<T*>(out_g[i] + j0 * (int64_t)(Core::eout(i) * ES))
                            : reinterpret_cast<T*>(dump + Cfg::dump_off(i));
        else
          out[i] = reinterpret_cast<T*>(ost + out_off[i]);
      }
      core.tile(prm, in, out, j0, ns);
    }
    mbar_arrive(consumed + si);
    if (NOUT > 0 && !Cfg::DIRECT) {
      fence_proxy_async_smem();  // our shared-memory writes -> visible to the bulk stores
      mbar_arrive(full_out + so);
    }
  }
  core.finish(prm, chain, valid);  // all compute threads: cores may use warp collectives here
}

// Host-side launcher: configures dynamic shared memory once per instantiation.
template <class Core, int C, int K, int NSI, int NSO>
inline cudaError_t launch_chain_sweep(const typename Core::Params& prm, int64_t nchains,
                                      cudaStream_t s, bool spread = false) {
  using Cfg = SweepCfg<Core, C, K, NSI, NSO>;
  auto kern = chain_sweep_kernel<Core, C, K, NSI, NSO>;
  static SmemOnce once;  // per instantiation, per device
  {
    cudaError_t e = ensure_smem(once, kern, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
  }
  // few chains: spread them over all 148 SMs (a CTA then serves fewer than C chains)
  int64_t cpb = C;
  if (spread && nchains < (int64_t)148 * C) {
    cpb = (nchains + 147) / 148;
    if (cpb < 1) cpb = 1;
  }
  const unsigned grid = (unsigned)((nchains + cpb - 1) / cpb);
  kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, s>>>(prm, (int)cpb, tuning(12));
  return cudaGetLastError();
}

}  // namespace mf
