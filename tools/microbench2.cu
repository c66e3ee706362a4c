// How much do warps parked in mbarrier.try_wait (or polling) slow down a latency-bound warp?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// warp 0: n iterations of STS -> syncwarp -> 12 LDS -> 18 FP64 (3-deep chain) ; other warps: wait
template <int WAITMODE>
__global__ void k(double* out, long long* cyc, int n, double seed) {
  __shared__ double sm[128];
  __shared__ uint64_t bar;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
  sm[threadIdx.x & 127] = seed;
  __syncthreads();
  double x[6];
  for (int s = 0; s < 6; ++s) x[s] = seed + lane + s;
  if (warp == 0) {
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
      sm[lane] = x[0];
      __syncwarp();
      const double p = sm[(i & 15)];
      double ur[6], uc[6];
#pragma unroll
      for (int s = 0; s < 6; ++s) { ur[s] = sm[(lane + s) & 63]; uc[s] = sm[(s * 7 + (lane >> 3)) & 63]; }
#pragma unroll
      for (int s = 0; s < 6; ++s) x[s] = fma(-ur[s], uc[s], p * x[s]) * 0.5;
      __syncwarp();
    }
    long long t1 = clock64();
    if (lane == 0) cyc[0] = t1 - t0;
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory");
  } else {
    if (WAITMODE == 0) {
      asm volatile("{\n.reg .pred P1;\nW0:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D0;\nbra W0;\nD0:\n}" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
    } else if (WAITMODE == 1) {
      asm volatile("{\n.reg .pred P1;\nW1:\nmbarrier.test_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D1;\nbra W1;\nD1:\n}" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
    } else if (WAITMODE == 2) {
      asm volatile("{\n.reg .pred P1;\nW2:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D2;\nnanosleep.u32 64;\nbra W2;\nD2:\n}" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
    }
  }
  out[threadIdx.x] = x[0] + x[1] + x[2] + x[3] + x[4] + x[5];
}

template <int WM>
void run(const char* name, int threads) {
  double* out; long long* cyc;
  cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 8);
  const int n = 20000;
  k<WM><<<1, threads>>>(out, cyc, n, 1.5);
  k<WM><<<1, threads>>>(out, cyc, n, 1.5);
  long long h = 0;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%-78s %8.1f cycles/stage %s\n", name, (double)h / n, e == cudaSuccess ? "" : cudaGetErrorString(e));
}
int main() {
  run<0>("stage loop alone (1 warp)", 32);
  run<0>("stage loop + 3 warps in try_wait (other SMSPs)", 128);
  run<0>("stage loop + 7 warps in try_wait (1 shares the SMSP)", 256);
  run<0>("stage loop + 11 warps in try_wait", 384);
  run<1>("stage loop + 7 warps polling test_wait", 256);
  run<2>("stage loop + 7 warps try_wait + nanosleep(64)", 256);
  return 0;
}
