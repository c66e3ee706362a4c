// Probe of cp.async.bulk.tensor (rank 3, tiled, no swizzle) legality rules on the GPU at hand: one load or store
// of one box at given coordinates through a map with given extents / strides.  Prints OK or the CUDA error.
//   tma_probe <load|store> <elem bytes 4|8> <g0> <g1> <g2> <stride0 B> <stride1 B> <b0> <b1> <b2> <x> <y> <z> [base offset B]
// Build: nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o tools/tma_probe tools/tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap map, int store, int x, int y, int z, uint32_t bytes,
                      unsigned long long* out) {
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm);
  unsigned char* buf = sm + 128;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (uint32_t i = threadIdx.x; i < bytes / 8; i += blockDim.x) reinterpret_cast<unsigned long long*>(buf)[i] = 0x1111111111111111ull * (1 + i % 7);
  __syncthreads();
  if (threadIdx.x == 0) {
    if (store) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(
                       reinterpret_cast<uint64_t>(&map)),
                   "r"(x), "r"(y), "r"(z), "r"(s32(buf))
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    } else {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
      asm volatile(
          "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
              s32(buf)),
          "l"(reinterpret_cast<uint64_t>(&map)), "r"(x), "r"(y), "r"(z), "r"(s32(bar))
          : "memory");
      uint32_t done = 0;
      long spins = 0;
      while (!done && spins < 2000000) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done)
                     : "r"(s32(bar))
                     : "memory");
        ++spins;
      }
      out[0] = done ? 1 : 0;
      unsigned long long acc = 0;
      for (uint32_t i = 0; i < bytes / 8; ++i) acc += reinterpret_cast<unsigned long long*>(buf)[i] != 0;
      out[1] = acc;
    }
  }
}

int main(int argc, char** argv) {
  if (argc < 14) return 2;
  const int store = !strcmp(argv[1], "store");
  const int es = atoi(argv[2]);
  cuuint64_t g[3] = {(cuuint64_t)atoll(argv[3]), (cuuint64_t)atoll(argv[4]), (cuuint64_t)atoll(argv[5])};
  cuuint64_t st[2] = {(cuuint64_t)atoll(argv[6]), (cuuint64_t)atoll(argv[7])};
  cuuint32_t b[3] = {(cuuint32_t)atoi(argv[8]), (cuuint32_t)atoi(argv[9]), (cuuint32_t)atoi(argv[10])};
  const int x = atoi(argv[11]), y = atoi(argv[12]), z = atoi(argv[13]);
  const size_t off = argc > 14 ? (size_t)atoll(argv[14]) : 0;
  cuuint32_t e[3] = {1, 1, 1};
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
    printf("no encode fn\n");
    return 1;
  }
  using Fn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  const size_t total = (size_t)st[1] * g[2] + (size_t)st[0] * g[1] + g[0] * es + (1 << 20);
  char* d = nullptr;
  cudaMalloc(&d, total + off + 4096);
  cudaMemset(d, 0x3c, total + off + 4096);
  CUtensorMap m;
  CUresult rc = reinterpret_cast<Fn>(fn)(&m, es == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                                         d + 2048 + off, g, st, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    printf("ENCODE_FAIL %d\n", (int)rc);
    return 0;
  }
  unsigned long long* out;
  cudaMalloc(&out, 16);
  cudaMemset(out, 0, 16);
  const uint32_t bytes = b[0] * b[1] * b[2] * es;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  probe<<<1, 128, 128 + bytes + 128>>>(m, store, x, y, z, bytes, out);
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) {
    printf("FAIL %s\n", cudaGetErrorString(err));
    return 0;
  }
  unsigned long long h[2];
  cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
  printf("OK done=%llu nonzero_words=%llu\n", h[0], h[1]);
  return 0;
}
