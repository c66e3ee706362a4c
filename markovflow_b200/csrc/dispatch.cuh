// Host-side helpers shared by the C-ABI translation units: dtype / block-size dispatch, launch
// error capture.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>

#include "../../include/markovflow_b200.h"

namespace mf {

template <typename T> struct TypeTag { using type = T; };
template <int N> struct IntTag { static constexpr int value = N; };

void set_last_error(const char* msg);
int check_launch();  // returns MF_OK or MF_ERR_CUDA after cudaGetLastError()
int tuning(int knob);  // value set through mf_set_tuning (0 = default)

// Calls f(TypeTag<T>{}, IntTag<D>{}) for dtype in {f32,f64} and 1 <= D <= MF_SMALL_D_MAX.
template <typename F>
int dispatch_small(int dtype, int64_t D, F&& f) {
#define MF_CASE_D(n)                                                        \
  case n:                                                                   \
    if (dtype == MF_F64) return f(TypeTag<double>{}, IntTag<n>{});          \
    return f(TypeTag<float>{}, IntTag<n>{});
  if (dtype != MF_F32 && dtype != MF_F64) return MF_ERR_BAD_ARG;
  switch (D) {
    MF_CASE_D(1) MF_CASE_D(2) MF_CASE_D(3) MF_CASE_D(4)
    MF_CASE_D(5) MF_CASE_D(6) MF_CASE_D(7) MF_CASE_D(8)
    default: return MF_ERR_UNSUPPORTED;
  }
#undef MF_CASE_D
}

template <typename F>
int dispatch_dtype(int dtype, F&& f) {
  if (dtype == MF_F64) return f(TypeTag<double>{});
  if (dtype == MF_F32) return f(TypeTag<float>{});
  return MF_ERR_BAD_ARG;
}

// large-block implementations (capi_big.cu)
int big_cholesky(int dtype, const void* diag, const void* sub, const void* rhs, void* out_diag,
                 void* out_sub, void* out_x, void* out_logdet, int32_t* info, int64_t B, int64_t T,
                 int64_t D, cudaStream_t s);
int big_solve(int dtype, const void* ld, const void* ls, const void* rhs, void* out, int64_t n_rhs,
              int64_t Bm, int64_t T, int64_t D, int transpose, cudaStream_t s);

// Seed folds of the parallel-in-time paths: one warp per chain (shuffle scan over the segment
// elements, log2 depth) from this many segments per chain on, else one thread per chain (sequential
// fold).  Measured on config 5 (B=1024, 28 segments): nat_seed_kernel 99 us thread-per-chain.
// tuning knob 10 overrides the threshold.
inline bool warp_fold(int64_t P) { return P >= (tuning(10) > 0 ? tuning(10) : 8); }

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: it is set once per (kernel
// instantiation, device), tracked in a bit mask that each launcher keeps as a function-local static.
struct SmemOnce {
  std::atomic<uint64_t> mask{0};
};
template <class K>
inline cudaError_t ensure_smem(SmemOnce& once, K kern, size_t bytes) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const uint64_t bit = 1ull << (dev & 63);
  if (once.mask.load(std::memory_order_acquire) & bit) return cudaSuccess;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return e;
  once.mask.fetch_or(bit, std::memory_order_release);
  return cudaSuccess;
}

inline unsigned grid_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }

}  // namespace mf
