// Shared-memory staging for chain-streaming sweeps.
//
// The public layout is chain-contiguous ([B, T, ...]); a thread that owns a chain therefore reads
// a private, strided stream -- the worst case for coalescing.  Here "copy warps" move each chain's
// next K steps global -> shared with cp.async (lanes along the contiguous time/element axis, so
// every warp instruction is one fully used 256-byte segment) into a TRANSPOSED tile
//     tile[f][c],  f = step_in_tile * E_total + element,  c = chain in CTA (row padded to C+1)
// which the compute warp reads bank-conflict free (lane == chain).  Results go the other way.
// mbarriers (cp.async.mbarrier.arrive.noinc) signal tile arrival; no thread ever blocks on its own
// load.  Works for any alignment (element-granular copies), any T (tail tiles are predicated).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace mf {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// all prior cp.async of this thread arrive on `bar` when they complete (count pre-provisioned)
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// Completion of this thread's element cp.async on `bar`.  Default: cp.async.mbarrier.arrive.noinc -- the arrival
// is triggered by the hardware when the copies land (the documented cp.async + mbarrier pattern; the thread does
// not wait).  wait_all != 0 (tuning knob 12, for compute-sanitizer racecheck, which does not model that
// completion path and reports the consumers' reads as hazards): the thread waits for its copies itself and then
// arrives with an ordinary release arrive.  Both make the copies visible to the waiting threads.
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar, int wait_all) {
  if (wait_all) {
    asm volatile("cp.async.wait_all;" ::: "memory");
    mbar_arrive(bar);
  } else {
    cp_async_arrive_noinc(bar);
  }
}

template <int BYTES>
__device__ __forceinline__ void cp_async_elem(void* smem, const void* gmem) {
  static_assert(BYTES == 4 || BYTES == 8, "element copies are 4 or 8 bytes");
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_u32(smem)), "l"(gmem),
               "n"(BYTES)
               : "memory");
}

// ---- TMA (bulk async copy) primitives: 1-D cp.async.bulk, no tensor map needed -----------------

// this thread arrives on `bar` and announces `bytes` of bulk-copy traffic that will complete on it
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}

// global -> shared bulk copy (16-byte aligned addresses, size multiple of 16), completes on `bar`
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// shared -> global bulk copy, tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void tma_store_1d(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void tma_store_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

// wait until at most N of this thread's bulk groups still have to READ their shared source
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

// make generic-proxy shared-memory writes visible to the async proxy (before a TMA store)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// One input or output stream of a sweep: E elements per step, `len` steps per chain.
template <typename T>
struct Stream {
  T* base;            // chain 0, step 0
  int64_t len;        // steps available in this stream (T or T-1)
  int64_t chain_mod;  // stream chain = global chain % chain_mod (broadcast of matrices over samples)
};

// Per-CTA table of element offsets of each chain's first record in a stream (computed once per
// kernel so that no 64-bit division/multiplication sits in the copy loops); -1 marks "no chain".
template <typename T, int E, int C>
__device__ __forceinline__ void fill_chain_offsets(int64_t* __restrict__ table,
                                                   const Stream<T>& st, int64_t chain0,
                                                   int64_t nchains_total, int tid, int nthreads) {
  for (int c = tid; c < C; c += nthreads) {
    const int64_t chain = chain0 + c;
    table[c] = (chain < nchains_total) ? (chain % st.chain_mod) * st.len * E : int64_t(-1);
  }
}

// Copy warps: global -> tile for `E` elements/step at element offset OFF inside the step record.
//   tile layout: tile[(s * ETOT + OFF + e) * CP + c]
// warp `w` of `nw` copy warps handles chains w, w+nw, ...; lanes run along i = s*E + e, so each
// warp instruction reads one contiguous run of a single chain.
template <typename T, int E, int OFF, int ETOT, int K, int C, int CP>
__device__ __forceinline__ void tile_load(T* __restrict__ tile, const T* __restrict__ base,
                                          const int64_t* __restrict__ chain_off, int64_t len,
                                          int64_t k0, int w, int nw, int lane) {
  constexpr int N = K * E;
  constexpr int ITERS = (N + 31) / 32;
  int f[ITERS];
  bool ok[ITERS];
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int i = it * 32 + lane;
    const int s = i / E, e = i - s * E;
    ok[it] = (i < N) && (k0 + s < len);
    f[it] = (s * ETOT + OFF + e) * CP;
  }
  const int64_t koff = k0 * E + lane;
#pragma unroll 4
  for (int c = w; c < C; c += nw) {
    const int64_t off = chain_off[c];
    const T* src = base + off + koff;
#pragma unroll
    for (int it = 0; it < ITERS; ++it)
      if (ok[it] && off >= 0) cp_async_elem<sizeof(T)>(tile + f[it] + c, src + it * 32);
  }
}

// Copy warps: tile -> global (coalesced along the chain's contiguous axis).
template <typename T, int E, int OFF, int ETOT, int K, int C, int CP>
__device__ __forceinline__ void tile_store(const T* __restrict__ tile, T* __restrict__ base,
                                           const int64_t* __restrict__ chain_off, int64_t len,
                                           int64_t k0, int w, int nw, int lane) {
  constexpr int N = K * E;
  constexpr int ITERS = (N + 31) / 32;
  int f[ITERS];
  bool ok[ITERS];
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int i = it * 32 + lane;
    const int s = i / E, e = i - s * E;
    ok[it] = (i < N) && (k0 + s < len);
    f[it] = (s * ETOT + OFF + e) * CP;
  }
  const int64_t koff = k0 * E + lane;
#pragma unroll 4
  for (int c = w; c < C; c += nw) {
    const int64_t off = chain_off[c];
    T* dst = base + off + koff;
    T v[ITERS];
#pragma unroll
    for (int it = 0; it < ITERS; ++it)
      if (ok[it]) v[it] = tile[f[it] + c];
#pragma unroll
    for (int it = 0; it < ITERS; ++it)
      if (ok[it] && off >= 0) __stcs(dst + it * 32, v[it]);
  }
}

}  // namespace mf
