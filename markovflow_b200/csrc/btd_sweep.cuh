// Block-tridiagonal Cholesky(+solve) as a Core of the generic chain sweeps (sweep.cuh / sweep2.cuh):
// the per-chain arithmetic is CholCore (chol_core.cuh), exactly as in btd_tma.cuh.
#pragma once
#include "btd_tma.cuh"
#include "ssm_sweep.cuh"

namespace mf {

template <typename T>
struct CholSweepParams {
  const T *diag, *sub, *rhs;
  T *od, *os, *ox, *logdet;
  int32_t* info;
  int64_t B, Tn;
};

template <typename T_, int D, bool RHS>
struct CholSweepCore {
  using T = T_;
  using Params = CholSweepParams<T>;
  static constexpr int DD = D * D;
  static constexpr int NIN = RHS ? 3 : 2, NOUT = RHS ? 3 : 2;
  static constexpr bool BACKWARD = false;
  static constexpr int ein(int i) { return i < 2 ? DD : D; }
  static constexpr int eout(int i) { return i < 2 ? DD : D; }
  static __device__ __forceinline__ int64_t num_chains(const Params& p) { return p.B; }
  static __device__ __forceinline__ int64_t max_steps(const Params& p) { return p.Tn; }
  static __device__ __forceinline__ StreamGeom in_geom(const Params& p, int i, int64_t c) {
    if (i == 1) return geom_outgoing<T>(p.sub, c, p.Tn, DD);
    return geom_states<T>(i == 0 ? p.diag : p.rhs, c, p.Tn, ein(i));
  }
  static __device__ __forceinline__ StreamGeom out_geom(const Params& p, int i, int64_t c) {
    if (i == 1) return geom_outgoing<T>(p.os, c, p.Tn, DD);
    return geom_states<T>(i == 0 ? p.od : p.ox, c, p.Tn, eout(i));
  }
  CholCore<T, D, RHS, ContiguousLayout<D>> core;
  __device__ __forceinline__ void init(const Params&, int64_t) { core.init(); }
  __device__ __forceinline__ void tile(const Params& p, const T* const* in, T* const* out,
                                       int64_t j0, int ns) {
    core.tile(in[0], in[1], RHS ? in[2] : nullptr, out[0], out[1], RHS ? out[2] : nullptr, ns, j0,
              p.Tn, p.logdet != nullptr);
  }
  __device__ __forceinline__ void finish(const Params& p, int64_t c, bool valid) {
    if (!valid) return;
    if (p.logdet) p.logdet[c] = core.log_det();
    if (p.info) p.info[c] = core.fail;
  }
};

}  // namespace mf
