// Host side of the half-warp-per-chain large-block path (btd_big2.cuh): plan, workspace, the three passes.
// Called from big_cholesky (capi_big.cu) for 9 <= D <= 17.
#include "btd_big2.cuh"
#include "dispatch.cuh"

namespace mf {

namespace {

template <typename T>
__global__ void big2_sum_parts_kernel(const T* __restrict__ parts, T* __restrict__ out, int64_t B, int64_t P) {
  const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (b >= B) return;
  T s = T(0);
  for (int64_t p = 0; p < P; ++p) s += parts[b * P + p];  // in segment order: reproducible
  out[b] = s;
}

template <typename F>
int dispatch_big2(int dtype, int64_t D, F&& f) {
  if (dtype != MF_F32 && dtype != MF_F64) return MF_ERR_BAD_ARG;
#define MF_B2_CASE(n)                                               \
  case n:                                                           \
    if (dtype == MF_F64) return f(TypeTag<double>{}, IntTag<n>{});  \
    return f(TypeTag<float>{}, IntTag<n>{});
  switch (D) {
    MF_B2_CASE(9) MF_B2_CASE(10) MF_B2_CASE(11) MF_B2_CASE(12) MF_B2_CASE(13) MF_B2_CASE(14)
    MF_B2_CASE(15) MF_B2_CASE(16) MF_B2_CASE(17)
    default: return MF_ERR_UNSUPPORTED;
  }
#undef MF_B2_CASE
}

// Segments per chain.  The half-warp kernels keep ~22 chains resident per SM; with fewer chains than that the
// sweep is bound by the latency of a chain's step, so chains are cut in time until the GPU is full (every
// segment but the first and the last pays the ~2.7x element pass, so at least 4 segments to come out ahead).
// tuning knob 7: 1 = never cut, n >= 3 = that many segments.
Big2Plan make_plan(int64_t B, int64_t T) {
  Big2Plan pl;
  pl.P = 1;
  pl.L = T;
  // the element pass holds 148 SMs x 7 warps x 2 segments at once: aim at two full waves of it
  // (a partial last wave costs as much as a full one), plus the first and the last segment
  const int64_t resident = 148 * 7 * 2;
  int64_t P = 2 * resident / (B > 0 ? B : 1) + 2;
  if (P < 4) P = 1;  // enough chains to fill the GPU as they are
  if (tuning(7) == 1) P = 1;
  if (tuning(7) >= 3) P = tuning(7);
  if (P > T / 32) P = T / 32;  // segments of at least 32 steps
  if (P < 4 && tuning(7) < 3) return pl;
  if (P < 3) return pl;
  pl.L = (T + P - 1) / P;
  pl.P = (T + pl.L - 1) / pl.L;
  if (pl.P < 3) { pl.P = 1; pl.L = T; }
  return pl;
}

// Side stream on which segment 0 is factorised while the element pass runs (fork / join with events; the
// pattern is legal under stream capture).  One per device, created on first use.
struct SideStream {
  cudaStream_t s = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};
SideStream g_side[64];

SideStream* side_stream() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  SideStream& ss = g_side[dev];
  if (!ss.s) {
    // keep the stream-ordered pool's memory between calls: the workspace is re-allocated on every call and
    // a pool that returns it to the driver at each synchronisation pays the mapping again (tens of ms)
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      uint64_t keep = ~uint64_t(0);
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    if (cudaStreamCreateWithFlags(&ss.s, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&ss.fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ss.join, cudaEventDisableTiming) != cudaSuccess)
      return nullptr;
  }
  return &ss;
}

}  // namespace

int big2_cholesky(int dtype, const void* diag, const void* sub, const void* rhs, void* out_diag, void* out_sub,
                  void* out_x, void* out_logdet, int32_t* info, int64_t B, int64_t T, int64_t D, cudaStream_t s) {
  return dispatch_big2(dtype, D, [&](auto tt, auto dd) -> int {
    using Tp = typename decltype(tt)::type;
    constexpr int kD = decltype(dd)::value;
    using C = Big2<Tp, kD>;
    using E = Big2E<Tp, kD>;
    Big2Plan pl = make_plan(B, sub ? T : 1);
    if (!sub) { pl.P = 1; pl.L = T; }
    const int64_t Bpad = (B + 1) & ~int64_t(1);
    auto factor = out_logdet ? big2_factor_kernel<Tp, kD, true> : big2_factor_kernel<Tp, kD, false>;
    auto element = big2_element_kernel<Tp, kD>;
    auto fold = big2_fold_kernel<Tp, kD>;
    constexpr size_t smem_f = sizeof(Tp) * (size_t)C::PER_CHAIN * 4;  // 2 warps = 4 chains per CTA
    constexpr size_t smem_e = sizeof(Tp) * (size_t)E::PER_CHAIN * 2;  // 1 warp  = 2 chains per CTA
    constexpr size_t smem_fold = sizeof(Tp) * (size_t)C::COLS * 4;
    if (cudaFuncSetAttribute(factor, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f) != cudaSuccess ||
        cudaFuncSetAttribute(element, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_e) != cudaSuccess)
      return check_launch();
    if (info && cudaMemsetAsync(info, 0, sizeof(int32_t) * B, s) != cudaSuccess) return check_launch();
    // workspace: elements | seeds | log-det parts
    const size_t n_el = (size_t)B * pl.P * C::ELEM, n_sd = (size_t)B * pl.P * C::SEED, n_ld = (size_t)B * pl.P;
    Tp* ws = nullptr;
    const bool need_ws = pl.P > 1 || out_logdet;
    if (need_ws && cudaMallocAsync((void**)&ws, sizeof(Tp) * (n_el + n_sd + n_ld), s) != cudaSuccess)
      return check_launch();
    Tp* elems = ws;
    Tp* seeds = ws ? ws + n_el : nullptr;
    Tp* ldp = (ws && out_logdet) ? ws + n_el + n_sd : nullptr;
    // write_seed: only segment 0's sweep of pass 1 hands a seed on (pass 3 must neither overwrite the fold's
    // seeds nor read blocks that a later segment is overwriting in place)
    auto launch_factor = [&](int64_t p_lo, int64_t p_hi, bool write_seed, cudaStream_t st) {
      const int64_t nv = (p_hi - p_lo) * Bpad;  // virtual chains, p-major
      const unsigned grid = (unsigned)((nv + 3) / 4);
      factor<<<grid, 64, smem_f, st>>>((const Tp*)diag, (const Tp*)sub, (const Tp*)rhs, (Tp*)out_diag, (Tp*)out_sub,
                                      (Tp*)out_x, ldp, info, seeds, write_seed ? seeds : nullptr, B, T, pl, p_lo, p_hi);
    };
    int rc = MF_OK;
    if (pl.P == 1) {
      launch_factor(0, 1, false, s);
      rc = check_launch();
    } else {
      // pass 1: segment 0 for good (it also seeds segment 1) + the elements of segments 1..P-2
      // (the two are independent: segment 0's latency-bound sweep runs on a side stream under the element pass)
      SideStream* ss = pl.P > 2 ? side_stream() : nullptr;
      if (ss && cudaEventRecord(ss->fork, s) == cudaSuccess && cudaStreamWaitEvent(ss->s, ss->fork, 0) == cudaSuccess) {
        launch_factor(0, 1, true, ss->s);
        cudaEventRecord(ss->join, ss->s);
      } else {
        ss = nullptr;
        launch_factor(0, 1, true, s);
      }
      if (pl.P > 2) {
        const int64_t nv = (pl.P - 2) * Bpad;
        element<<<(unsigned)((nv + 1) / 2), 32, smem_e, s>>>((const Tp*)diag, (const Tp*)sub, (const Tp*)rhs, elems,
                                                             info, B, T, pl);
        if (ss) cudaStreamWaitEvent(s, ss->join, 0);
        fold<<<(unsigned)((Bpad + 3) / 4), 64, smem_fold, s>>>(elems, seeds, info, B, pl);
      }
      rc = check_launch();
      if (rc == MF_OK) {
        launch_factor(1, pl.P, false, s);  // pass 3
        rc = check_launch();
      }
    }
    if (rc == MF_OK && out_logdet) {
      big2_sum_parts_kernel<Tp><<<grid_for(B, 128), 128, 0, s>>>(ldp, (Tp*)out_logdet, B, pl.P);
      rc = check_launch();
    }
    if (ws) cudaFreeAsync(ws, s);
    return rc;
  });
}

}  // namespace mf
