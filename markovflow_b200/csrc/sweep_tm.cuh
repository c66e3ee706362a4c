// Tensor-map variant of the chain sweep (sweep.cuh): ONE cp.async.bulk.tensor per (stream, tile) and CTA
// instead of one 1-D bulk copy per (stream, tile, chain).
//
// Why: the D <= 2 sweeps move 128-256 B per 1-D copy and are bound by the SM's bulk-copy issue rate
// (ncu on config 5: the compute warps sit on `full_in`, the producer warps on UBLKCP; 1.5e7 copies for 1e7
// state-steps).  The rows a CTA serves are regular in global memory -- consecutive segments of a chain are
// L*E elements apart, consecutive chains len*E -- so a rank-3 tensor map
//        dim0 = x   position inside the segment, in elements        (extent clips the ragged last tile)
//        dim1 = seg segment inside the chain   (stride L*E)          (extent clips the special rows, below)
//        dim2 = c   chain                      (stride len*E)        (extent clips the last CTA)
// describes them all, and a box [K*E + pad] x [rows of the CTA] moves a whole stage of a stream with one
// instruction in each direction.
//
// Shared-memory layout = the box, dense: row r of the CTA at r * PITCH, PITCH = K*E*sizeof(T) + PAD with
// PAD = 16 or 32 B chosen so that PITCH is an ODD multiple of 16 B (lanes spread over the banks exactly as
// in sweep.cuh; the cores' shared-memory indexing does not change).  The pad is part of the box: it sits on
// the side the sweep comes from (front for forward sweeps, back for backward ones).  Loads over-read the
// neighbouring tile into it; for stores the compute thread copies the neighbouring tile's adjacent PAD
// bytes (its own previous outputs, still in the previous output stage) into it, so the box rewrites those
// bytes with the values they already have.  Past either end of a segment the pad is out of the map's
// bounds and is dropped / zero-filled by the hardware.
//
// Special rows.  With P >= 3 segments per chain the first segment (streams that start one step early read
// the element BEFORE the chain there) and the last one (ragged length) are not in the map: the map's seg
// axis starts at segment 1 and has extent P - 2, so the box clips them.  They are served exactly as in
// sweep.cuh -- a producer lane per (stream, special row) with 1-D bulk copies, any alignment -- through NX
// extra row regions per stream and stage; their compute lanes point at those regions.  With P == 1 (rows
// are whole chains) every row is regular: shifted streams use a shifted x coordinate and the map's x
// extent clips the step that does not exist.
//
// Tensor STORES must not see a negative coordinate (measured on B200: UTMASTG raises "illegal instruction";
// loads zero-fill).  Two consequences:
//  * rows of a chain are served in the rotated order  q -> segment (q + 1) mod P : the map's rows are
//    q = 0 .. P-3 and the two special rows come LAST (q = P-2, P-1 <-> last and first segment), clipped by
//    the extent on the positive side;
//  * the first tile of a FORWARD sweep would start at x = -pad: it is stored through a second map per output
//    stream whose x extent ends with the tile, from data the compute threads place at the start of the row
//    (x0 = 0, the trailing pad is clipped instead of the leading one).
//
// A Core opts in with a host-side description of its streams (the device-side in_geom / out_geom stay the
// truth for the special rows):
//   static void tm_describe(const Params&, TmStream* in, TmStream* out);   // base, chain length, shift
//   static int64_t tm_segments(const Params&), tm_seg_len(const Params&), tm_chains(const Params&)
#pragma once
#include <cuda.h>

#include "sweep.cuh"

namespace mf {

// One stream as the host sees it: element (chain c, local step j of the segment starting at k0) is entry
// c * chain_len + k0 + j + shift of a [B, chain_len, E] array.
struct TmStream {
  const void* base;   // nullptr: stream absent (optional output)
  int64_t chain_len;  // steps per chain in this stream (T or T - 1)
  int shift;          // 0, or -1 for "incoming" transition streams
};

// maps [0, NIN): inputs; [NIN, NIN+NOUT): outputs; [NIN+NOUT, NIN+2*NOUT): first-tile maps of forward outputs
template <int NS>
struct alignas(64) TmPack {
  CUtensorMap map[NS > 0 ? NS : 1];
  int xshift[NS > 0 ? NS : 1];   // elements added to the tile's x coordinate
  int present[NS > 0 ? NS : 1];
};

// ---- device primitives ------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, int x, int y, int z,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(z),
      "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, int x, int y, int z, const void* smem_src) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(z), "r"(smem_u32(smem_src))
               : "memory");
}

// p as a T*, known to the compiler as aligned to min(16, largest power of two dividing the record size)
template <typename T>
__device__ __forceinline__ T* tm_assume_aligned(const void* p, int record_bytes) {
  void* q = const_cast<void*>(p);
  if (record_bytes % 16 == 0) return reinterpret_cast<T*>(__builtin_assume_aligned(q, 16));
  if (record_bytes % 8 == 0) return reinterpret_cast<T*>(__builtin_assume_aligned(q, 8));
  return reinterpret_cast<T*>(q);
}

template <class Core, int C, int K, int NSI, int NSO, int NX>
struct SweepTmCfg {
  using T = typename Core::T;
  static constexpr int ES = (int)sizeof(T);
  static constexpr int NIN = Core::NIN, NOUT = Core::NOUT;
  static constexpr int NSO_EFF = NOUT > 0 ? NSO : 0;
  static constexpr int padb(int E) { return ((K * E * ES / 16) % 2 == 0) ? 16 : 32; }
  static constexpr int pitch(int E) { return K * E * ES + padb(E); }
  static constexpr int xreg(int E) { return pitch(E) + 16; }  // special-row region: room for the a0 shift
  static constexpr int up128(int x) { return (x + 127) / 128 * 128; }
  static constexpr int box_bytes(int E) { return up128(C * pitch(E)); }
  static constexpr int sz(int E) { return box_bytes(E) + up128(NX * xreg(E)); }
  static constexpr int off_in(int i) {
    int o = 0;
    for (int q = 0; q < i; ++q) o += sz(Core::ein(q));
    return o;
  }
  static constexpr int off_out(int i) {
    int o = 0;
    for (int q = 0; q < i; ++q) o += sz(Core::eout(q));
    return o;
  }
  static constexpr int STAGE_IN = off_in(NIN), STAGE_OUT = off_out(NOUT);
  static constexpr int NCW = C / 32;
  static constexpr int THREADS = 32 * (NCW + 1 + (NOUT > 0 ? 1 : 0));
  static constexpr int NBAR = 2 * NSI + 2 * NSO_EFF;
  static constexpr size_t SMEM_BYTES =
      (size_t)STAGE_IN * NSI + (size_t)STAGE_OUT * NSO_EFF + sizeof(uint64_t) * NBAR + 128;
  static constexpr bool align_ok() {
    for (int i = 0; i < NIN; ++i)
      if ((K * Core::ein(i) * ES) % 16 != 0 || K * Core::ein(i) + padb(Core::ein(i)) / ES > 256) return false;
    for (int i = 0; i < NOUT; ++i)
      if ((K * Core::eout(i) * ES) % 16 != 0 || K * Core::eout(i) + padb(Core::eout(i)) / ES > 256) return false;
    return true;
  }
  static constexpr bool FITS = align_ok() && SMEM_BYTES <= (size_t)232448 && C % 32 == 0 &&
                               NIN * (1 + NX) <= 32 && NOUT * (1 + NX) <= 32;
};

// Row r of the CTA whose first row is v0 serves position q of its chain; q -> segment (q + 1) mod P.
struct TmRow {
  int64_t chain;  // logical virtual chain c * P + segment
  int special;    // extra region serving the row, -1: the row is in the box
};
__device__ __forceinline__ TmRow tm_row(int r, int64_t v0, int64_t P, int cpb) {
  TmRow o;
  if (P < 3) {
    o.chain = v0 + r;
    o.special = -1;
    return o;
  }
  int64_t c0, q;
  int ci = 0;
  if (P <= cpb) {
    ci = r / (int)P;
    q = r - ci * (int)P;
    c0 = v0 + (int64_t)ci * P;
  } else {
    const int64_t q0 = v0 % P;
    q = q0 + r;
    c0 = v0 - q0;
  }
  o.chain = c0 + (q + 1) % P;
  o.special = q >= P - 2 ? 2 * ci + (int)(q - (P - 2)) : -1;
  return o;
}
// inverse: the row served by extra region x (-1: none)
__device__ __forceinline__ int tm_special_row(int x, int64_t v0, int64_t P, int cpb) {
  if (P < 3) return -1;
  if (P <= cpb) {
    const int ci = x >> 1;
    if ((int64_t)(ci + 1) * P > cpb) return -1;
    return ci * (int)P + (int)P - 2 + (x & 1);
  }
  if (x > 1) return -1;
  const int64_t r = P - 2 + x - v0 % P;
  return (r >= 0 && r < cpb) ? (int)r : -1;
}

template <class Core, int C, int K, int NSI, int NSO, int NX>
__global__ void __launch_bounds__(SweepTmCfg<Core, C, K, NSI, NSO, NX>::THREADS)
chain_sweep_tm_kernel(const __grid_constant__ TmPack<Core::NIN + 2 * Core::NOUT> tm,
                      const typename Core::Params prm, const int cpb, const int64_t P, const int elem_wait) {
  using Cfg = SweepTmCfg<Core, C, K, NSI, NSO, NX>;
  using T = typename Core::T;
  constexpr int ES = Cfg::ES, NIN = Cfg::NIN, NOUT = Cfg::NOUT, NSOE = Cfg::NSO_EFF;
  constexpr bool BWD = Core::BACKWARD;
  static_assert(Cfg::FITS, "tensor-map sweep configuration does not fit");
  extern __shared__ __align__(128) unsigned char smem_raw_tm[];
  char* in_stages = reinterpret_cast<char*>(smem_raw_tm);
  in_stages += (128 - (smem_u32(in_stages) & 127)) & 127;
  char* out_stages = in_stages + (size_t)Cfg::STAGE_IN * NSI;
  uint64_t* bars = reinterpret_cast<uint64_t*>(out_stages + (size_t)Cfg::STAGE_OUT * NSOE);
  uint64_t* full_in = bars;
  uint64_t* consumed = bars + NSI;
  uint64_t* full_out = bars + 2 * NSI;
  uint64_t* empty_out = bars + 2 * NSI + NSOE;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t nchains = Core::num_chains(prm);
  const int64_t v0 = (int64_t)blockIdx.x * cpb;
  const int64_t nsteps = Core::max_steps(prm);
  const int64_t ntiles = (nsteps + K - 1) / K;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NSI; ++i) {
      mbar_init(full_in + i, NIN + 2 * NIN * NX);
      mbar_init(consumed + i, C);
    }
    for (int i = 0; i < NSOE; ++i) {
      mbar_init(full_out + i, C);
      mbar_init(empty_out + i, NOUT + NOUT * NX);
    }
    mbar_fence_init();
  }
  __syncthreads();

  auto tile_id = [&](int64_t t) { return BWD ? ntiles - 1 - t : t; };
  // box coordinates of the CTA on the (seg, chain) axes
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
      const int padel = Cfg::padb(E) / ES;
      const uint32_t bytes = tm.present[stream] ? (uint32_t)((K * E + padel) * ES) * (uint32_t)cpb : 0u;
      const int boff = Cfg::off_in(stream);
      auto issue = [&](int64_t t) {
        const int si = (int)(t % NSI);
        uint64_t* bar = full_in + si;
        mbar_arrive_expect_tx(bar, bytes);
        if (bytes) {
          const int x0 = (int)(tile_id(t) * K) * E + tm.xshift[stream] - (BWD ? 0 : padel);
          tma_load_3d(in_stages + (size_t)si * Cfg::STAGE_IN + boff, &tm.map[stream], x0, y0, z0, bar);
        }
      };
      for (int64_t t = 0; t < NSI && t < ntiles; ++t) issue(t);
      for (int64_t t = 0; t + NSI < ntiles; ++t) {
        mbar_wait(consumed + (int)(t % NSI), (uint32_t)((t / NSI) & 1));
        issue(t + NSI);
      }
    } else if (lane < NIN + NIN * NX) {
      const int stream = (lane - NIN) / NX, q = (lane - NIN) % NX;
      const int E = Core::ein(stream);
      const int row = tm_special_row(q, v0, P, cpb);
      const bool valid = row >= 0 && v0 + row < nchains;
      const SweepSeg sg = make_seg(
          valid ? Core::in_geom(prm, stream, tm_row(row, v0, P, cpb).chain) : StreamGeom{nullptr, 0, 0}, valid);
      const int roff = Cfg::off_in(stream) + Cfg::box_bytes(E) + q * Cfg::xreg(E);
      auto issue = [&](int64_t t) {
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
      for (int64_t t = 0; t < NSI && t < ntiles; ++t) issue(t);
      for (int64_t t = 0; t + NSI < ntiles; ++t) {
        mbar_wait(consumed + (int)(t % NSI), (uint32_t)((t / NSI) & 1));
        issue(t + NSI);
      }
    }
    return;
  }
  if (NOUT > 0 && warp == Cfg::NCW + 1) {
    // ------------------------------------ storer warp --------------------------------------------
    constexpr int NSOD = NSO > 0 ? NSO : 1;
    if (lane < NOUT) {
      const int stream = lane;
      const int E = Core::eout(stream);
      const int padel = Cfg::padb(E) / ES;
      const int boff = Cfg::off_out(stream);
      const bool present = tm.present[NIN + stream] != 0;
      for (int64_t t = 0; t < ntiles; ++t) {
        const int so = (int)(t % NSOD);
        mbar_wait(full_out + so, (uint32_t)((t / NSOD) & 1));
        if (present) {
          const char* src = out_stages + (size_t)so * Cfg::STAGE_OUT + boff;
          if (!BWD && t == 0) {
            tma_store_3d(&tm.map[NIN + NOUT + stream], 0, y0, z0, src);
          } else {
            const int x0 = (int)(tile_id(t) * K) * E + tm.xshift[NIN + stream] - (BWD ? 0 : padel);
            tma_store_3d(&tm.map[NIN + stream], x0, y0, z0, src);
          }
        }
        tma_store_commit();
        tma_store_wait_read<0>();
        mbar_arrive(empty_out + so);
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    } else if (lane < NOUT + NOUT * NX) {
      const int stream = (lane - NOUT) / NX, q = (lane - NOUT) % NX;
      const int E = Core::eout(stream);
      const int row = tm_special_row(q, v0, P, cpb);
      const bool valid = row >= 0 && v0 + row < nchains;
      const SweepSeg sg = make_seg(
          valid ? Core::out_geom(prm, stream, tm_row(row, v0, P, cpb).chain) : StreamGeom{nullptr, 0, 0}, valid);
      const int roff = Cfg::off_out(stream) + Cfg::box_bytes(E) + q * Cfg::xreg(E);
      for (int64_t t = 0; t < ntiles; ++t) {
        const int so = (int)(t % NSOD);
        mbar_wait(full_out + so, (uint32_t)((t / NSOD) & 1));
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
  const int r = warp * 32 + lane;
  const bool valid = r < cpb && v0 + r < nchains;
  const TmRow row = tm_row(valid ? r : 0, v0, P, cpb);
  const int64_t chain = row.chain;
  const int sidx = valid ? row.special : -1;
  int in_off[NIN > 0 ? NIN : 1], out_off[NOUT > 0 ? NOUT : 1];
  int out_x0[NOUT > 0 ? NOUT : 1];  // forward sweeps: byte shift of the first tile's data (<= 0) from the row start
#pragma unroll
  for (int i = 0; i < NIN; ++i) {
    const int E = Core::ein(i);
    if (sidx >= 0) {
      const StreamGeom g = Core::in_geom(prm, i, chain);
      in_off[i] = Cfg::off_in(i) + Cfg::box_bytes(E) + sidx * Cfg::xreg(E) +
                  (int)(reinterpret_cast<uintptr_t>(g.step0) & 15);
    } else {
      in_off[i] = Cfg::off_in(i) + r * Cfg::pitch(E) + (BWD ? 0 : Cfg::padb(E));
    }
  }
#pragma unroll
  for (int i = 0; i < NOUT; ++i) {
    const int E = Core::eout(i);
    if (sidx >= 0) {
      const StreamGeom g = Core::out_geom(prm, i, chain);
      out_off[i] = Cfg::off_out(i) + Cfg::box_bytes(E) + sidx * Cfg::xreg(E) +
                   (int)(reinterpret_cast<uintptr_t>(g.step0) & 15);
    } else {
      out_off[i] = Cfg::off_out(i) + r * Cfg::pitch(E) + (BWD ? 0 : Cfg::padb(E));
    }
    out_x0[i] = tm.xshift[NIN + i] * ES;
  }
  Core core;
  if (valid) core.init(prm, chain);
  constexpr int NSOD = NSO > 0 ? NSO : 1;
  for (int64_t t = 0; t < ntiles; ++t) {
    const int si = (int)(t % NSI);
    mbar_wait(full_in + si, (uint32_t)((t / NSI) & 1));
    const char* ist = in_stages + (size_t)si * Cfg::STAGE_IN;
    char* ost = nullptr;
    int so = 0;
    if (NOUT > 0) {
      so = (int)(t % NSOD);
      mbar_wait(empty_out + so, (uint32_t)(((t / NSOD) & 1) ^ 1));
      ost = out_stages + (size_t)so * Cfg::STAGE_OUT;
      if (t > 0 && valid && sidx < 0) {
        // the store box includes the pad: fill it with the adjacent bytes of the previous tile's outputs
        const char* pst = out_stages + (size_t)((t - 1) % NSOD) * Cfg::STAGE_OUT;
#pragma unroll
        for (int i = 0; i < NOUT; ++i) {
          const int E = Core::eout(i);
          const int pb = Cfg::padb(E), data = K * E * ES;
          // (the first tile of a forward sweep sits at the row start, shifted by the stream's x shift)
          const char* src = pst + out_off[i] + (BWD ? 0 : data - pb + (t == 1 ? out_x0[i] - pb : 0));
          char* dst = ost + out_off[i] + (BWD ? data : -pb);
          uint4 v[2];
          v[0] = *reinterpret_cast<const uint4*>(src);
          if (pb == 32) v[1] = *reinterpret_cast<const uint4*>(src + 16);
          *reinterpret_cast<uint4*>(dst) = v[0];
          if (pb == 32) *reinterpret_cast<uint4*>(dst + 16) = v[1];
        }
      }
    }
    const int64_t j0 = tile_id(t) * K;
    const int ns = (int)((nsteps - j0 < K) ? (nsteps - j0) : K);
    if (valid) {
      const T* in[NIN > 0 ? NIN : 1];
      T* out[NOUT > 0 ? NOUT : 1];
#pragma unroll
      for (int i = 0; i < NIN; ++i) {
        // records of a mapped stream start on multiples of min(16, record size) bytes (tm_make_map checks the
        // global side): 128-bit shared-memory loads
        in[i] = tm_assume_aligned<T>(ist + in_off[i], Core::ein(i) * ES);
      }
#pragma unroll
      for (int i = 0; i < NOUT; ++i) {
        int o = out_off[i];
        if (!BWD && t == 0 && sidx < 0) o += out_x0[i] - Cfg::padb(Core::eout(i));
        out[i] = reinterpret_cast<T*>(ost + o);
      }
      core.tile(prm, in, out, j0, ns);
    }
    mbar_arrive(consumed + si);
    if (NOUT > 0) {
      fence_proxy_async_smem();
      mbar_arrive(full_out + so);
    }
  }
  core.finish(prm, chain, valid);
}

// ---- host side ------------------------------------------------------------------------------------------

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
using TmEncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
TmEncodeFn tm_encode_fn();  // capi_tm.cu; nullptr when the driver does not provide it
void tm_count_launch();     // bumps the counter behind mf_tm_launch_count() (tests check which engine ran)

// Builds the map of one stream.  Returns false when the stream's geometry cannot be described (alignment).
//   rows: box height (rows of a CTA);  P, L: segmentation;  B: chains;  
//   store: an output (no negative coordinate may occur);  fwd: forward sweep;  first_tile: the map of a forward
//   output's first tile (x extent ends with the tile)
template <typename T>
inline bool tm_make_map(CUtensorMap* map, int* xshift, const TmStream& st, int E, int K, int padel, int cpb,
                        int64_t P, int64_t L, int64_t B, bool store, bool fwd, bool first_tile = false) {
  constexpr int ES = (int)sizeof(T);
  TmEncodeFn enc = tm_encode_fn();
  if (!enc) return false;
  const int64_t step_b = (int64_t)E * ES;
  const int64_t chain_stride = st.chain_len * step_b;
  char* base = const_cast<char*>(static_cast<const char*>(st.base));
  cuuint64_t gdim[3], gstride[2];
  cuuint32_t box[3], estr[3] = {1, 1, 1};
  box[0] = (cuuint32_t)(K * E + padel);
  if (P < 3) {
    if (P != 1) return false;
    // rows = chains; the shift lives in the x coordinate, the x extent clips the step that does not exist
    if (chain_stride % 16 != 0 || st.chain_len <= 0) return false;
    *xshift = st.shift * E;
    gdim[0] = (cuuint64_t)(st.chain_len * E);
    gdim[1] = (cuuint64_t)B;
    gdim[2] = 1;
    gstride[0] = (cuuint64_t)chain_stride;
    gstride[1] = (cuuint64_t)chain_stride;
    box[1] = (cuuint32_t)cpb;
    box[2] = 1;
  } else {
    // rows = segments 1 .. P-2 of every chain; x = 0 is the segment's first entry (one step early for
    // "incoming" streams)
    const int64_t seg_stride = L * step_b;
    if (seg_stride % 16 != 0 || chain_stride % 16 != 0) return false;
    if (L + st.shift < 0) return false;
    base += (L + st.shift) * step_b;
    *xshift = 0;
    gdim[0] = (cuuint64_t)(L * E);
    gdim[1] = (cuuint64_t)(P - 2);  // rows q = 0 .. P-3 <-> segments 1 .. P-2
    gdim[2] = (cuuint64_t)B;
    gstride[0] = (cuuint64_t)seg_stride;
    gstride[1] = (cuuint64_t)chain_stride;
    if (P <= cpb) {
      if (cpb % P != 0) return false;
      box[1] = (cuuint32_t)P;
      box[2] = (cuuint32_t)(cpb / P);
    } else {
      if (P % cpb != 0) return false;
      box[1] = (cuuint32_t)cpb;
      box[2] = 1;
    }
  }
  // the x coordinate of every box must be a multiple of 16 bytes (measured: UTMALDG / UTMASTG raise "illegal
  // instruction" otherwise); tiles and pads are, the shift has to be
  if ((*xshift * ES) % 16 != 0) return false;
  if (store) {
    // coordinates of a store: x0 = xshift (backward, last tile) or K*E + xshift - pad (forward, second tile)
    if (!fwd && *xshift < 0) return false;
    if (fwd && K * E + *xshift - padel < 0) return false;
    if (first_tile) {
      const int64_t n = (int64_t)K * E + *xshift;  // elements of the first tile that exist
      if (n < 1) return false;
      if ((cuuint64_t)n < gdim[0]) gdim[0] = (cuuint64_t)n;
    }
  }
  if (reinterpret_cast<uintptr_t>(base) % 16 != 0) return false;
  if (gdim[0] >= (1ull << 31) || gdim[1] >= (1ull << 32) || gdim[2] >= (1ull << 32)) return false;
  if (gstride[0] >= (1ull << 40) || gstride[1] >= (1ull << 40)) return false;
  if (box[0] > 256 || box[1] > 256 || box[2] > 256) return false;
  const CUtensorMapDataType dt = ES == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  // L2 promotion 256 B: a box row is 144-160 bytes and the sweep walks along x, so the promoted line is the
  // next tile's data (measured, config 5 float64: ssm_to_expectations 0.504 -> 0.476 ms,
  // naturals_to_ssm_params 0.581 -> 0.578 ms against 128 B; 64 B: 0.568 / 0.622 ms)
  const CUresult rc = enc(map, dt, 3, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return rc == CUDA_SUCCESS;
}

// Rows per CTA for P segments per chain: whole chains (P <= C) or a divisor of P (P > C); 0: no layout.
inline int tm_rows_per_cta(int64_t P, int C, int NX) {
  if (P == 1) return C;
  if (P < 3) return 0;
  if (P <= C) {
    int n = C / (int)P;
    if (2 * n > NX) n = NX / 2;  // two special rows (first / last segment) per chain in the CTA
    return n * (int)P >= C / 2 ? n * (int)P : 0;
  }
  for (int c = C; c >= C / 2; --c)
    if (P % c == 0) return c;
  return 0;
}

// Launches the tensor-map sweep; returns cudaErrorNotSupported when the geometry cannot be mapped (the
// caller then uses the 1-D engine of sweep.cuh).
template <class Core, int C, int K, int NSI, int NSO, int NX>
inline cudaError_t launch_chain_sweep_tm(const typename Core::Params& prm, cudaStream_t s) {
  using Cfg = SweepTmCfg<Core, C, K, NSI, NSO, NX>;
  using T = typename Core::T;
  constexpr int NIN = Core::NIN, NOUT = Core::NOUT;
  const int64_t P = Core::tm_segments(prm), L = Core::tm_seg_len(prm), B = Core::tm_chains(prm);
  if (P >= 3 && L < K) return cudaErrorNotSupported;
  int cpb = tm_rows_per_cta(P, C, NX);
  if (cpb <= 0) return cudaErrorNotSupported;
  const int64_t nrows = B * P;
  // The engine pays off where the 1-D engine is bound by bulk-copy ISSUE: every SM busy with full CTAs.  With
  // fewer rows than 148 CTAs' worth a sweep is bound by the latency of one warp's steps, and the 1-D engine's
  // longer tiles win (measured, 4096 chains x 1e4 steps, D = 2: 1.84 ms against 2.30 ms).  Knob 13 = 2 lifts
  // the threshold (tests).
  if (nrows < (int64_t)148 * C && tuning(13) != 2) return cudaErrorNotSupported;
  if (P == 1 && nrows < (int64_t)148 * C) {  // few chains: spread over the SMs
    cpb = (int)((nrows + 147) / 148);
    if (cpb < 1) cpb = 1;
  }
  TmStream in[NIN > 0 ? NIN : 1], out[NOUT > 0 ? NOUT : 1];
  Core::tm_describe(prm, in, out);
  TmPack<NIN + 2 * NOUT> pack;
  for (int i = 0; i < NIN; ++i) {
    const int E = Core::ein(i);
    pack.present[i] = in[i].base != nullptr;
    pack.xshift[i] = 0;
    if (pack.present[i] &&
        !tm_make_map<T>(&pack.map[i], &pack.xshift[i], in[i], E, K, Cfg::padb(E) / Cfg::ES, cpb, P, L, B, false,
                        !Core::BACKWARD))
      return cudaErrorNotSupported;
  }
  for (int i = 0; i < NOUT; ++i) {
    const int E = Core::eout(i);
    pack.present[NIN + i] = out[i].base != nullptr;
    pack.xshift[NIN + i] = 0;
    pack.present[NIN + NOUT + i] = 0;
    pack.xshift[NIN + NOUT + i] = 0;
    if (!pack.present[NIN + i]) continue;
    const int padel = Cfg::padb(E) / Cfg::ES;
    if (!tm_make_map<T>(&pack.map[NIN + i], &pack.xshift[NIN + i], out[i], E, K, padel, cpb, P, L, B, true,
                        !Core::BACKWARD))
      return cudaErrorNotSupported;
    if (!Core::BACKWARD) {
      pack.present[NIN + NOUT + i] = 1;
      if (!tm_make_map<T>(&pack.map[NIN + NOUT + i], &pack.xshift[NIN + NOUT + i], out[i], E, K, padel, cpb, P, L,
                          B, true, true, true))
        return cudaErrorNotSupported;
    }
  }
  auto kern = chain_sweep_tm_kernel<Core, C, K, NSI, NSO, NX>;
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
