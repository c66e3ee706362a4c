// Producer warp (stage loop: STS, syncwarp, arrive, 13 LDS, 18 FP64) with consumer warps that wait
// for every arrival on per-column mbarriers: what does signalling cost the PRODUCER?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void arrive(uint32_t a) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory"); }
__device__ __forceinline__ void waitp(uint32_t a, uint32_t par) {
  asm volatile("{\n.reg .pred P1;\nLW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra LD;\nbra LW;\nLD:\n}" ::"r"(a), "r"(par) : "memory");
}
constexpr int NB = 16;
// MODE 0: no arrive; 1: arrive, no waiters; 2: arrive + waiters (consumers wait then read 4 LDS)
template <int MODE>
__global__ void k(double* out, long long* cyc, int n, double seed) {
  __shared__ double sm[1024];
  __shared__ uint64_t bar[NB];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) for (int i = 0; i < NB; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar + i)), "r"(1));
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = seed;
  __syncthreads();
  uint32_t b0;
  asm volatile("mov.u32 %0, %1;" : "=r"(b0) : "r"(smem_u32(bar)));
  double x[6];
  for (int s = 0; s < 6; ++s) x[s] = seed + lane + s;
  if (warp == 0) {
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        sm[j * 32 + lane] = x[0];
        __syncwarp();
        if (MODE >= 1 && lane == 0) arrive(b0 + 8 * j);
        const double p = sm[j * 32 + (j & 15)];
        double ur[6], uc[6];
#pragma unroll
        for (int s = 0; s < 6; ++s) { ur[s] = sm[j * 32 + ((lane + s) & 31)]; uc[s] = sm[j * 32 + ((s * 7 + (lane >> 3)) & 31)]; }
#pragma unroll
        for (int s = 0; s < 6; ++s) x[s] = fma(-ur[s], uc[s], p * x[s]) * 0.5;
      }
    }
    long long t1 = clock64();
    if (lane == 0) cyc[0] = t1 - t0;
  } else if (MODE == 2) {
    for (int i = 0; i < n; ++i) {
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        waitp(b0 + 8 * j, i & 1);
        double a = sm[j * 32 + lane], b = sm[j * 32 + ((lane + warp) & 31)];
        x[0] = fma(a, b, x[0]) * 0.5;
      }
    }
  }
  out[threadIdx.x] = x[0] + x[1] + x[2] + x[3] + x[4] + x[5];
}
template <int M>
void run(const char* name, int threads) {
  double* out; long long* cyc;
  cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 8);
  const int n = 2000;
  k<M><<<1, threads>>>(out, cyc, n, 1.5);
  k<M><<<1, threads>>>(out, cyc, n, 1.5);
  long long h = 0;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%-70s %8.1f cycles/stage %s\n", name, (double)h / n / NB, e == cudaSuccess ? "" : cudaGetErrorString(e));
}
int main() {
  run<0>("producer stage loop, no signalling", 32);
  run<1>("producer + mbarrier.arrive per stage, no waiters", 32);
  run<2>("producer + arrive, 1 consumer warp waiting per stage", 64);
  run<2>("producer + arrive, 3 consumer warps waiting per stage", 128);
  run<2>("producer + arrive, 4 consumer warps (one shares the SMSP)", 160);
  return 0;
}
