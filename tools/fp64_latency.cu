// Micro-benchmark: dependent-chain latency (cycles) of the FP64 operations on the Cholesky
// critical path.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void chain(double* out, long long* cyc, double seed, int n) {
  double x = seed + threadIdx.x * 1e-3, y = 1.000001;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
    if (OP == 0) x = fma(x, y, 1e-9);
    if (OP == 1) x = rsqrt(x) + 1.5;
    if (OP == 2) x = sqrt(x) + 1.5;
    if (OP == 3) x = 1.0 / x + 1.5;
    if (OP == 4) x = log(x) + 2.5;
    if (OP == 5) x = __shfl_xor_sync(0xffffffffu, x, 1) + 1e-9;
    if (OP == 6) { float f = rsqrtf((float)x); x = (double)f + 1.5; }
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 1024 * 8); cudaMallocManaged(&cyc, 8);
  const char* names[] = {"dfma", "rsqrt+add", "sqrt+add", "div+add", "log+add", "shfl+add", "f32 rsqrt cvt+add"};
  const int n = 4096;
  for (int threads : {1, 32}) {
    for (int op = 0; op < 7; ++op) {
      for (int rep = 0; rep < 2; ++rep) {
        switch (op) {
          case 0: chain<0><<<1, threads>>>(out, cyc, 1.1, n); break;
          case 1: chain<1><<<1, threads>>>(out, cyc, 1.1, n); break;
          case 2: chain<2><<<1, threads>>>(out, cyc, 1.1, n); break;
          case 3: chain<3><<<1, threads>>>(out, cyc, 1.1, n); break;
          case 4: chain<4><<<1, threads>>>(out, cyc, 1.1, n); break;
          case 5: chain<5><<<1, threads>>>(out, cyc, 1.1, n); break;
          case 6: chain<6><<<1, threads>>>(out, cyc, 1.1, n); break;
        }
        cudaDeviceSynchronize();
      }
      printf("threads=%2d %-20s %.1f cycles/iter\n", threads, names[op], (double)*cyc / n);
    }
  }
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("clock rate attr %d kHz\n", clk);
  return 0;
}
