// FP64 throughput / latency microbenchmarks on B200 (sm_100a): DFMA against the FP64 tensor-core
// instruction mma.sync.aligned.m8n8k4.row.col.f64 (SASS DMMA).  Answers, with numbers, whether the
// block products of the D >= 16 path (SYRK / TRSM-as-GEMM, BASELINE.json north_star) should run on DMMA.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/dmma_bench tools/dmma_bench.cu && tools/dmma_bench
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int ILP>
__global__ void k_dmma(double* out, int iters, long long* cycles) {
  double c[ILP][2];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x * 1e-3; c[i][1] = i; }
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) dmma(c[i][0], c[i][1], a, b);
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int ILP>
__global__ void k_dfma(double* out, int iters, long long* cycles) {
  double c[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) c[i] = threadIdx.x * 1e-3 + i;
  const double a = 1.0 + threadIdx.x * 1e-12, b = 1e-9;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i] = fma(c[i], a, b);
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <class K>
void run(const char* name, K kern, int ilp, int warps, int blocks, double flops_per_instr, int sms) {
  double* out;
  long long* cyc;
  cudaMalloc(&out, sizeof(double) * 1024 * 1024);
  cudaMalloc(&cyc, 8);
  const int iters = 4096;
  kern<<<blocks, 32 * warps>>>(out, 16, cyc);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  kern<<<blocks, 32 * warps>>>(out, iters, cyc);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double instr = (double)iters * ilp;
  const double total_flops = instr * flops_per_instr * warps * blocks;
  printf("{\"bench\": \"%s\", \"ilp\": %d, \"warps_per_cta\": %d, \"ctas\": %d, \"cycles_per_instr_per_warp\": %.2f, "
         "\"tflops\": %.2f, \"ms\": %.4f}\n", name, ilp, warps, blocks, (double)h / instr, total_flops / (ms * 1e-3) / 1e12, ms);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_mhz\": %d}\n", p.name, sms, p.clockRate / 1000);
  // latency: one warp, one dependent chain
  run("dmma_m8n8k4_dependent_chain", k_dmma<1>, 1, 1, 1, 512.0, sms);
  run("dfma_dependent_chain", k_dfma<1>, 1, 1, 1, 64.0, sms);
  // issue rate of ONE warp with independent accumulators
  run("dmma_one_warp_ilp8", k_dmma<8>, 8, 1, 1, 512.0, sms);
  run("dfma_one_warp_ilp8", k_dfma<8>, 8, 1, 1, 64.0, sms);
  // one warp per scheduler (4 per SM), all SMs
  run("dmma_4warps_per_sm_ilp8", k_dmma<8>, 8, 4, sms, 512.0, sms);
  run("dfma_4warps_per_sm_ilp8", k_dfma<8>, 8, 4, sms, 64.0, sms);
  // saturated: 16 warps per SM
  run("dmma_16warps_per_sm_ilp8", k_dmma<8>, 8, 16, sms, 512.0, sms);
  run("dfma_16warps_per_sm_ilp8", k_dfma<8>, 8, 16, sms, 64.0, sms);
  run("dmma_32warps_per_sm_ilp4", k_dmma<4>, 4, 32, sms, 512.0, sms);
  run("dfma_32warps_per_sm_ilp4", k_dfma<4>, 4, 32, sms, 64.0, sms);
  return 0;
}
