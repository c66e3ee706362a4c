// In-place variant of the tensor-map sweep (sweep_tm.cuh): ONE ring of NS stages; a stage is loaded, the compute
// threads overwrite every input record with the output record of the same step, and the same stage is stored.
//
// Why: with separate input and output rings (2 + 2 stages of K = 4 steps fill the shared memory two CTAs can
// have) the loader runs ONE tile ahead, ~1 us of compute against ~2 us of loaded-DRAM latency: the compute warps
// of the config-5 sweeps spent a quarter of their time on `full_in`.  Here the same bytes are 4 stages and the
// loader runs two tiles ahead (load t+2 | ready t+1 | compute t | store t-1).
//
// A Core opts in by pairing every input stream with an output stream of the same record size,
//     static constexpr int tm_alias(int i)      output stream that reuses input stream i's slots
// and by computing every step's outputs FROM that step's inputs (then a slot is always read before it is
// written: the store's value depends on the load).  Sub-diagonal ("incoming") streams keep their own x shift in
// each direction: the load box and the store box of a tile start at their own coordinates, so a step's input
// and output land on the same shared-memory bytes whatever the two shifts are.
//
// The pad of a row (sweep_tm.cuh) is refilled before the store from a copy of the previous tile's adjacent output
// bytes kept in REGISTERS (the previous stage may already be reloading).
#pragma once
#include "sweep_tm.cuh"

namespace mf {

template <class Core, int C, int K, int NS, int NX>
struct SweepTmiCfg {
  using Base = SweepTmCfg<Core, C, K, NS, 0, NX>;  // input-side layout of sweep_tm.cuh
  using T = typename Core::T;
  static constexpr int ES = (int)sizeof(T);
  static constexpr int NIN = Core::NIN, NOUT = Core::NOUT;
  static constexpr int STAGE = Base::STAGE_IN;
  static constexpr int NCW = C / 32;
  static constexpr int THREADS = 32 * (NCW + 2);
  static constexpr size_t SMEM_BYTES = (size_t)STAGE * NS + sizeof(uint64_t) * 3 * NS + 128;
  static constexpr bool alias_ok() {
    if (NIN != NOUT) return false;
    bool used[NOUT > 0 ? NOUT : 1] = {};
    for (int i = 0; i < NIN; ++i) {
      const int o = Core::tm_alias(i);
      if (o < 0 || o >= NOUT || used[o] || Core::ein(i) != Core::eout(o)) return false;
      used[o] = true;
    }
    return true;
  }
  static constexpr int alias_inv(int o) {
    for (int i = 0; i < NIN; ++i)
      if (Core::tm_alias(i) == o) return i;
    return 0;
  }
  static constexpr bool FITS = alias_ok() && Base::align_ok() && SMEM_BYTES <= (size_t)232448 && C % 32 == 0 &&
                               NIN * (1 + NX) <= 32 && NOUT * (1 + NX) <= 32 && NS >= 2;
};

template <class Core, int C, int K, int NS, int NX>
__global__ void __launch_bounds__(SweepTmiCfg<Core, C, K, NS, NX>::THREADS)
chain_sweep_tmi_kernel(const __grid_constant__ TmPack<Core::NIN + 2 * Core::NOUT> tm,
                       const typename Core::Params prm, const int cpb, const int64_t P, const int elem_wait) {
  using Cfg = SweepTmiCfg<Core, C, K, NS, NX>;
  using L = typename Cfg::Base;
  using T = typename Core::T;
  constexpr int ES = Cfg::ES, NIN = Cfg::NIN, NOUT = Cfg::NOUT;
  constexpr bool BWD = Core::BACKWARD;
  static_assert(Cfg::FITS, "in-place tensor-map sweep configuration does not fit");
  extern __shared__ __align__(128) unsigned char smem_raw_tmi[];
  char* stages = reinterpret_cast<char*>(smem_raw_tmi);
  stages += (128 - (smem_u32(stages) & 127)) & 127;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stages + (size_t)Cfg::STAGE * NS);
  uint64_t* full = bars;             // loads of a stage have landed
  uint64_t* computed = bars + NS;    // compute threads have turned the stage into outputs
  uint64_t* freed = bars + 2 * NS;   // the stores have finished reading the stage

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t nchains = Core::num_chains(prm);
  const int64_t v0 = (int64_t)blockIdx.x * cpb;
  const int64_t nsteps = Core::max_steps(prm);
  const int64_t ntiles = (nsteps + K - 1) / K;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) {
      mbar_init(full + i, NIN + 2 * NIN * NX);
      mbar_init(computed + i, C);
      mbar_init(freed + i, NOUT + NOUT * NX);
    }
    mbar_fence_init();
  }
  __syncthreads();

  auto tile_id = [&](int64_t t) { return BWD ? ntiles - 1 - t : t; };
  int y0, z0;
  if (P < 3) {
    y0 = (int)v0;
    z0 = 0;
  } else {
    y0 = (int)(v0 % P);
    z0 = (int)(v0 / P);
  }

  if (warp == Cfg::NCW) {
    // ------------------------------------ loader warp --------------------------------------------
    if (lane < NIN) {
      const int stream = lane;
      const int E = Core::ein(stream);
      const int padel = L::padb(E) / ES;
      const uint32_t bytes = tm.present[stream] ? (uint32_t)((K * E + padel) * ES) * (uint32_t)cpb : 0u;
      const int boff = L::off_in(stream);
      const int xs_in = tm.xshift[stream];
      int xs_out = 0;
#pragma unroll
      for (int i = 0; i < NIN; ++i)
        if (i == stream) xs_out = tm.xshift[NIN + Core::tm_alias(i)];
      for (int64_t t = 0; t < ntiles; ++t) {
        const int s = (int)(t % NS);
        if (t >= NS) mbar_wait(freed + s, (uint32_t)(((t / NS) & 1) ^ 1));
        uint64_t* bar = full + s;
        mbar_arrive_expect_tx(bar, bytes);
        if (bytes) {
          // a forward sweep's first tile sits at the row start (its store must not begin at a negative x)
          const int x0 = (!BWD && t == 0) ? xs_in - (xs_out < 0 ? xs_out : 0)
                                          : (int)(tile_id(t) * K) * E + xs_in - (BWD ? 0 : padel);
          tma_load_3d(stages + (size_t)s * Cfg::STAGE + boff, &tm.map[stream], x0, y0, z0, bar);
        }
      }
    } else if (lane < NIN + NIN * NX) {
      const int stream = (lane - NIN) / NX, q = (lane - NIN) % NX;
      const int E = Core::ein(stream);
      const int row = tm_special_row(q, v0, P, cpb);
      const bool valid = row >= 0 && v0 + row < nchains;
      const SweepSeg sg = make_seg(
          valid ? Core::in_geom(prm, stream, tm_row(row, v0, P, cpb).chain) : StreamGeom{nullptr, 0, 0}, valid);
      const int roff = L::off_in(stream) + L::box_bytes(E) + q * L::xreg(E);
      for (int64_t t = 0; t < ntiles; ++t) {
        const int s = (int)(t % NS);
        if (t >= NS) mbar_wait(freed + s, (uint32_t)(((t / NS) & 1) ^ 1));
        uint64_t* bar = full + s;
        const int64_t j0 = tile_id(t) * K;
        uint32_t tx = 0;
        int lo = 0, hi = 0, head = 0;
        if (sg.g) tx = sweep_ranges<ES, K>(sg, E, j0, lo, hi, head);
        mbar_arrive_expect_tx(bar, tx);
        if (sg.g && hi > lo) {
          char* sd = stages + (size_t)s * Cfg::STAGE + roff + sg.a0;
          const char* g0 = sg.g + j0 * (int64_t)(E * ES);
          if (tx) tma_load_1d(sd + lo + head, g0 + lo + head, tx, bar);
          for (int o = lo; o < lo + head; o += ES) cp_async_elem<ES>(sd + o, g0 + o);
          for (int o = lo + head + (int)tx; o < hi; o += ES) cp_async_elem<ES>(sd + o, g0 + o);
        }
        cp_async_arrive(bar, elem_wait);
      }
    }
    return;
  }
  if (warp == Cfg::NCW + 1) {
    // ------------------------------------ storer warp --------------------------------------------
    if (lane < NOUT) {
      const int stream = lane;
      const int E = Core::eout(stream);
      const int padel = L::padb(E) / ES;
      int boff = 0;
#pragma unroll
      for (int o = 0; o < NOUT; ++o)
        if (o == stream) boff = L::off_in(Cfg::alias_inv(o));
      const bool present = tm.present[NIN + stream] != 0;
      const int xs_out = tm.xshift[NIN + stream];
      for (int64_t t = 0; t < ntiles; ++t) {
        const int s = (int)(t % NS);
        mbar_wait(computed + s, (uint32_t)((t / NS) & 1));
        if (present) {
          const char* src = stages + (size_t)s * Cfg::STAGE + boff;
          if (!BWD && t == 0) {
            tma_store_3d(&tm.map[NIN + NOUT + stream], xs_out > 0 ? xs_out : 0, y0, z0, src);
          } else {
            const int x0 = (int)(tile_id(t) * K) * E + xs_out - (BWD ? 0 : padel);
            tma_store_3d(&tm.map[NIN + stream], x0, y0, z0, src);
          }
        }
        tma_store_commit();
        tma_store_wait_read<0>();
        mbar_arrive(freed + s);
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    } else if (lane < NOUT + NOUT * NX) {
      const int stream = (lane - NOUT) / NX, q = (lane - NOUT) % NX;
      const int E = Core::eout(stream);
      int istream = 0;
#pragma unroll
      for (int o = 0; o < NOUT; ++o)
        if (o == stream) istream = Cfg::alias_inv(o);
      const int row = tm_special_row(q, v0, P, cpb);
      const bool valid = row >= 0 && v0 + row < nchains;
      const int64_t chain = valid ? tm_row(row, v0, P, cpb).chain : 0;
      const SweepSeg sg = make_seg(valid ? Core::out_geom(prm, stream, chain) : StreamGeom{nullptr, 0, 0}, valid);
      // the region is laid out by the INPUT stream's misalignment; the bulk store needs the output's to match
      const int a_in = valid ? (int)(reinterpret_cast<uintptr_t>(Core::in_geom(prm, istream, chain).step0) & 15) : 0;
      const int roff = L::off_in(istream) + L::box_bytes(E) + q * L::xreg(E);
      for (int64_t t = 0; t < ntiles; ++t) {
        const int s = (int)(t % NS);
        mbar_wait(computed + s, (uint32_t)((t / NS) & 1));
        if (sg.g) {
          const int64_t j0 = tile_id(t) * K;
          int lo, hi, head;
          uint32_t tx = sweep_ranges<ES, K>(sg, E, j0, lo, hi, head);
          if (hi > lo) {
            const char* sd = stages + (size_t)s * Cfg::STAGE + roff + a_in;
            char* g0 = sg.g + j0 * (int64_t)(E * ES);
            if (a_in != sg.a0) { tx = 0; head = hi - lo; }  // element stores only
            if (tx) tma_store_1d(g0 + lo + head, sd + lo + head, tx);
            for (int o = lo; o < lo + head; o += ES)
              *reinterpret_cast<T*>(g0 + o) = *reinterpret_cast<const T*>(sd + o);
            for (int o = lo + head + (int)tx; o < hi; o += ES)
              *reinterpret_cast<T*>(g0 + o) = *reinterpret_cast<const T*>(sd + o);
          }
        }
        tma_store_commit();
        tma_store_wait_read<0>();
        mbar_arrive(freed + s);
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    return;
  }

  // --------------------------------- compute threads ------------------------------------------
  const int r = warp * 32 + lane;
  const bool valid = r < cpb && v0 + r < nchains;
  const TmRow row = tm_row(valid ? r : 0, v0, P, cpb);
  const int64_t chain = row.chain;
  const int sidx = valid ? row.special : -1;
  int off[NIN > 0 ? NIN : 1];    // row start of stream i (regular rows) / data start (special rows)
  int first0[NIN > 0 ? NIN : 1];  // forward sweeps: byte offset of local step 0 in the first tile
#pragma unroll
  for (int i = 0; i < NIN; ++i) {
    const int E = Core::ein(i);
    if (sidx >= 0) {
      const StreamGeom g = Core::in_geom(prm, i, chain);
      off[i] = L::off_in(i) + L::box_bytes(E) + sidx * L::xreg(E) + (int)(reinterpret_cast<uintptr_t>(g.step0) & 15);
      first0[i] = 0;
    } else {
      off[i] = L::off_in(i) + r * L::pitch(E);
      const int xo = tm.xshift[NIN + Core::tm_alias(i)];
      first0[i] = (xo < 0 ? xo : 0) * ES;
    }
  }
  Core core;
  if (valid) core.init(prm, chain);
  uint4 keep[NIN > 0 ? NIN : 1][2];  // the previous tile's output bytes next to this tile's pad
  for (int64_t t = 0; t < ntiles; ++t) {
    const int s = (int)(t % NS);
    mbar_wait(full + s, (uint32_t)((t / NS) & 1));
    char* st = stages + (size_t)s * Cfg::STAGE;
    const int64_t j0 = tile_id(t) * K;
    const int ns = (int)((nsteps - j0 < K) ? (nsteps - j0) : K);
    if (valid) {
      const T* in[NIN > 0 ? NIN : 1];
      T* out[NOUT > 0 ? NOUT : 1];
#pragma unroll
      for (int i = 0; i < NIN; ++i) {
        const int E = Core::ein(i);
        const int pb = L::padb(E), data = K * E * ES;
        int o = off[i];
        if (sidx < 0) {
          if (BWD) {
            if (t > 0) {  // trailing pad <- first bytes of the tile processed before (later in time)
              *reinterpret_cast<uint4*>(st + o + data) = keep[i][0];
              if (pb == 32) *reinterpret_cast<uint4*>(st + o + data + 16) = keep[i][1];
            }
          } else if (t == 0) {
            o += first0[i];
          } else {  // leading pad <- last bytes of the previous tile's outputs
            *reinterpret_cast<uint4*>(st + o) = keep[i][0];
            if (pb == 32) *reinterpret_cast<uint4*>(st + o + 16) = keep[i][1];
            o += pb;
          }
        }
        // every record of a mapped stream starts on a multiple of min(16, record size) bytes, in global memory
        // (tm_make_map checks bases and strides) and here: let the compiler use 128-bit shared-memory accesses
        T* slot = tm_assume_aligned<T>(st + o, E * ES);
        in[i] = slot;
        out[Core::tm_alias(i)] = slot;
      }
      core.tile(prm, in, out, j0, ns);
      if (sidx < 0) {
#pragma unroll
        for (int i = 0; i < NIN; ++i) {
          const int E = Core::ein(i);
          const int pb = L::padb(E), data = K * E * ES;
          // data start of this tile: row start (+ pad for forward tiles after the first; + shift for the first)
          const char* d0 = reinterpret_cast<const char*>(in[i]);
          const char* src = BWD ? d0 : d0 + data - pb;
          keep[i][0] = *reinterpret_cast<const uint4*>(src);
          if (pb == 32) keep[i][1] = *reinterpret_cast<const uint4*>(src + 16);
        }
      }
    }
    fence_proxy_async_smem();
    mbar_arrive(computed + s);
  }
  core.finish(prm, chain, valid);
}

template <class Core, int C, int K, int NS, int NX>
inline cudaError_t launch_chain_sweep_tmi(const typename Core::Params& prm, cudaStream_t s) {
  using Cfg = SweepTmiCfg<Core, C, K, NS, NX>;
  using L = typename Cfg::Base;
  using T = typename Core::T;
  constexpr int NIN = Core::NIN, NOUT = Core::NOUT;
  const int64_t P = Core::tm_segments(prm), Lseg = Core::tm_seg_len(prm), B = Core::tm_chains(prm);
  if (P >= 3 && Lseg < K) return cudaErrorNotSupported;
  int cpb = tm_rows_per_cta(P, C, NX);
  if (cpb <= 0) return cudaErrorNotSupported;
  const int64_t nrows = B * P;
  if (nrows < (int64_t)148 * C && tuning(13) != 2) return cudaErrorNotSupported;
  TmStream in[NIN], out[NOUT];
  Core::tm_describe(prm, in, out);
  TmPack<NIN + 2 * NOUT> pack;
  for (int i = 0; i < NIN; ++i) {
    const int E = Core::ein(i);
    pack.present[i] = in[i].base != nullptr;
    pack.xshift[i] = 0;
    if (!pack.present[i]) return cudaErrorNotSupported;  // an in-place stage needs every input
    if (!tm_make_map<T>(&pack.map[i], &pack.xshift[i], in[i], E, K, L::padb(E) / Cfg::ES, cpb, P, Lseg, B, false,
                        !Core::BACKWARD))
      return cudaErrorNotSupported;
  }
  for (int i = 0; i < NOUT; ++i) {
    const int E = Core::eout(i);
    const int padel = L::padb(E) / Cfg::ES;
    pack.present[NIN + i] = out[i].base != nullptr;
    pack.xshift[NIN + i] = 0;
    pack.present[NIN + NOUT + i] = 0;
    pack.xshift[NIN + NOUT + i] = 0;
    TmStream o = out[i];
    if (!o.base) {
      // an absent output still fixes where its slots sit relative to the input's: describe it without storing
      o.base = in[Cfg::alias_inv(i)].base;
    }
    if (!tm_make_map<T>(&pack.map[NIN + i], &pack.xshift[NIN + i], o, E, K, padel, cpb, P, Lseg, B, true,
                        !Core::BACKWARD))
      return cudaErrorNotSupported;
    if (!Core::BACKWARD) {
      pack.present[NIN + NOUT + i] = pack.present[NIN + i];
      if (!tm_make_map<T>(&pack.map[NIN + NOUT + i], &pack.xshift[NIN + NOUT + i], o, E, K, padel, cpb, P, Lseg, B,
                          true, true, true))
        return cudaErrorNotSupported;
    }
  }
  auto kern = chain_sweep_tmi_kernel<Core, C, K, NS, NX>;
  static SmemOnce once;
  {
    cudaError_t e = ensure_smem(once, kern, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
  }
  const unsigned grid = (unsigned)((nrows + cpb - 1) / cpb);
  kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, s>>>(pack, prm, cpb, P, tuning(12));
  tm_count_launch();
  return cudaGetLastError();
}

}  // namespace mf
