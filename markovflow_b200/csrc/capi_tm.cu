// cuTensorMapEncodeTiled, resolved once through the CUDA runtime (the library does not link libcuda).
#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <mutex>

#include "sweep_tm.cuh"

namespace mf {

TmEncodeFn tm_encode_fn() {
  static TmEncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<TmEncodeFn>(p);
    (void)cudaGetLastError();
  });
  return fn;
}

static std::atomic<int64_t> g_tm_launches{0};
void tm_count_launch() { g_tm_launches.fetch_add(1, std::memory_order_relaxed); }

}  // namespace mf

extern "C" int64_t mf_tm_launch_count(void) { return mf::g_tm_launches.load(std::memory_order_relaxed); }
