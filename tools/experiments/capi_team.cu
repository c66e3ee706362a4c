// Team (one CTA per chain) implementation of mf_btd_cholesky for large blocks and few chains
// (btd_team.cuh).  Called from capi_big.cu.
#include "btd_team.cuh"
#include "dispatch.cuh"

namespace mf {

namespace {

template <typename Tp, int kD>
int launch_team(const void* diag, const void* sub, const void* rhs, void* out_diag, void* out_sub,
                void* out_x, void* out_logdet, int32_t* info, int64_t B, int64_t T, cudaStream_t s) {
  using Cfg = TeamCfg<Tp, kD>;
  auto kern = btd_chol_team_kernel<Tp, kD>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)Cfg::SMEM_BYTES) != cudaSuccess)
      return check_launch();
    configured = true;
  }
  kern<<<(unsigned)B, Cfg::THREADS, Cfg::SMEM_BYTES, s>>>(
      (const Tp*)diag, (const Tp*)sub, (const Tp*)rhs, (Tp*)out_diag, (Tp*)out_sub, (Tp*)out_x,
      (Tp*)out_logdet, info, B, T);
  return check_launch();
}

}  // namespace


int team_cholesky(int dtype, const void* diag, const void* sub, const void* rhs, void* out_diag,
                  void* out_sub, void* out_x, void* out_logdet, int32_t* info, int64_t B, int64_t T,
                  int64_t D, cudaStream_t s) {
  if (dtype != MF_F64 && dtype != MF_F32) return MF_ERR_BAD_ARG;
#define MF_TEAM_CASE(n)                                                                           \
  case n:                                                                                         \
    if (dtype == MF_F64)                                                                          \
      return launch_team<double, n>(diag, sub, rhs, out_diag, out_sub, out_x, out_logdet, info, B, T, s); \
    return launch_team<float, n>(diag, sub, rhs, out_diag, out_sub, out_x, out_logdet, info, B, T, s);
  switch (D) {
    MF_TEAM_CASE(17)
    default: return MF_ERR_UNSUPPORTED;
  }
#undef MF_TEAM_CASE
}

}  // namespace mf

#ifdef MF_TEAM_DEBUG
extern "C" int mf_debug_team_timeline(long long* out) {
  return (int)cudaMemcpyFromSymbol(out, mf::g_team_dbg, sizeof(long long) * 160);
}
#endif
