// Dependent-issue latencies and single-warp issue costs on B200 that shape the chain kernels.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void k(double* out, long long* cyc, int n, double seed) {
  __shared__ double sm[64];
  __shared__ uint64_t bar[2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar + 1)), "r"(1));
  }
  sm[threadIdx.x & 63] = seed;
  __syncthreads();
  double a0 = seed + lane, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  long long t0 = clock64();
  if (MODE == 0) {  // dependent DFMA
    for (int i = 0; i < n; ++i) a0 = fma(a0, m, c);
  } else if (MODE == 1) {  // 8 independent DFMA chains: issue cost
    for (int i = 0; i < n; ++i) {
      a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
      a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
  } else if (MODE == 2) {  // STS -> syncwarp -> LDS round trip (dependent)
    for (int i = 0; i < n; ++i) {
      sm[lane] = a0;
      __syncwarp();
      a0 = sm[(lane + 1) & 31] + c;
      __syncwarp();
    }
  } else if (MODE == 3) {  // mbarrier arrive by lane 0 (issue cost seen by the warp), no waiting
    for (int i = 0; i < n; ++i) {
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
      a0 = fma(a0, m, c);
    }
  } else if (MODE == 4) {  // ping-pong between two warps through mbarriers: 2 hops per iteration
    uint64_t* mine = bar + warp;
    uint64_t* other = bar + (warp ^ 1);
    for (int i = 0; i < n; ++i) {
      if (warp == 0) {
        sm[lane] = a0;
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(other)) : "memory");
        asm volatile("{\n.reg .pred P1;\nW0:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D0;\nbra W0;\nD0:\n}" ::"r"(smem_u32(mine)), "r"(i & 1) : "memory");
        a0 = sm[32 + lane] + c;
      } else {
        asm volatile("{\n.reg .pred P1;\nW1:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D1;\nbra W1;\nD1:\n}" ::"r"(smem_u32(mine)), "r"(i & 1) : "memory");
        a0 = sm[lane] + c;
        sm[32 + lane] = a0;
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(other)) : "memory");
      }
    }
  } else if (MODE == 5) {  // ping-pong through a shared-memory flag (volatile polling), 2 hops / iteration
    volatile int* flag = reinterpret_cast<volatile int*>(bar);
    volatile double* vs = sm;
    for (int i = 0; i < n; ++i) {
      if (warp == 0) {
        vs[lane] = a0;
        __syncwarp();
        if (lane == 0) { __threadfence_block(); flag[0] = i + 1; }
        while (flag[2] != i + 1) {}
        a0 = vs[32 + lane] + c;
      } else {
        while (flag[0] != i + 1) {}
        a0 = vs[lane] + c;
        vs[32 + lane] = a0;
        __syncwarp();
        if (lane == 0) { __threadfence_block(); flag[2] = i + 1; }
      }
    }
  } else if (MODE == 6) {  // dependent rsqrt
    for (int i = 0; i < n; ++i) a0 = rsqrt(a0) + 2.0;
  } else if (MODE == 7) {  // shfl dependent
    for (int i = 0; i < n; ++i) a0 = __shfl_sync(0xffffffffu, a0, (lane + 1) & 31) + c;
  } else if (MODE == 8) {  // named barrier ping-pong between two warps (bar.sync 1, 64): 1 sync / iteration
    for (int i = 0; i < n; ++i) {
      sm[warp * 32 + lane] = a0;
      asm volatile("bar.sync 1, 64;" ::: "memory");
      a0 = sm[(warp ^ 1) * 32 + lane] + c;
      asm volatile("bar.sync 2, 64;" ::: "memory");
    }
  } else if (MODE == 9) {  // 8 independent DMUL+DFMA+DMUL triples
    for (int i = 0; i < n; ++i) {
      a0 = fma(a0, m, c); a1 = a1 * m; a2 = fma(a2, m, c); a3 = a3 * m;
      a4 = fma(a4, m, c); a5 = a5 * m; a6 = fma(a6, m, c); a7 = a7 * m;
    }
  }
  long long t1 = clock64();
  out[threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int MODE>
void run(const char* name, int threads, int per_iter) {
  double* out;
  long long* cyc;
  cudaMalloc(&out, 1024 * 8);
  cudaMalloc(&cyc, 8);
  const int n = 20000;
  k<MODE><<<1, threads>>>(out, cyc, n, 1.5);
  k<MODE><<<1, threads>>>(out, cyc, n, 1.5);
  long long h = 0;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%-70s %8.1f cycles%s\n", name, (double)h / n / per_iter, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run<0>("DFMA dependent latency", 32, 1);
  run<1>("DFMA issue cost, one warp, 8 independent chains (per DFMA)", 32, 8);
  run<1>("DFMA issue cost, 4 warps on 4 SMSPs (per DFMA per warp)", 128, 8);
  run<1>("DFMA issue cost, 8 warps = 2 per SMSP (per DFMA per warp)", 256, 8);
  run<9>("DMUL/DFMA mix issue cost, one warp (per op)", 32, 8);
  run<2>("STS -> syncwarp -> LDS -> syncwarp round trip", 32, 1);
  run<3>("mbarrier.arrive (lane 0) + 1 dependent DFMA per iteration", 32, 1);
  run<4>("mbarrier ping-pong between two warps (per hop)", 64, 2);
  run<5>("flag-polling ping-pong between two warps (per hop)", 64, 2);
  run<6>("rsqrt(double) dependent (+1 DADD)", 32, 1);
  run<7>("shfl dependent (+1 DADD)", 32, 1);
  run<8>("named-barrier exchange between two warps (per iteration: STS, bar, LDS, bar)", 64, 1);
  return 0;
}
