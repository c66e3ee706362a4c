// Large-block (MF_SMALL_D_MAX < D <= MF_BIG_D_MAX) implementations behind mf_btd_cholesky and
// mf_btd_solve: one warp per chain (btd_big.cuh).  Called from capi_btd.cu.
#include "btd_big.cuh"
#include "dispatch.cuh"

namespace mf {

int big2_cholesky(int dtype, const void* diag, const void* sub, const void* rhs, void* out_diag, void* out_sub,
                  void* out_x, void* out_logdet, int32_t* info, int64_t B, int64_t T, int64_t D,
                  cudaStream_t s);  // capi_big2.cu

namespace {

template <typename F>
int dispatch_big(int dtype, int64_t D, F&& f) {
  if (dtype != MF_F32 && dtype != MF_F64) return MF_ERR_BAD_ARG;
#define MF_BIG_CASE(n)                                                \
  case n:                                                             \
    if (dtype == MF_F64) return f(TypeTag<double>{}, IntTag<n>{});    \
    return f(TypeTag<float>{}, IntTag<n>{});
  switch (D) {
    MF_BIG_CASE(9) MF_BIG_CASE(10) MF_BIG_CASE(11) MF_BIG_CASE(12) MF_BIG_CASE(13) MF_BIG_CASE(14)
    MF_BIG_CASE(15) MF_BIG_CASE(16) MF_BIG_CASE(17) MF_BIG_CASE(18) MF_BIG_CASE(19) MF_BIG_CASE(20)
    MF_BIG_CASE(21) MF_BIG_CASE(22) MF_BIG_CASE(23) MF_BIG_CASE(24) MF_BIG_CASE(25) MF_BIG_CASE(26)
    MF_BIG_CASE(27) MF_BIG_CASE(28) MF_BIG_CASE(29) MF_BIG_CASE(30) MF_BIG_CASE(31) MF_BIG_CASE(32)
    default: return MF_ERR_UNSUPPORTED;
  }
#undef MF_BIG_CASE
}

template <typename K>
int set_smem(K kern, size_t bytes) {  // set on every launch: per-device attribute, cheap
  if (bytes > 48 * 1024 &&
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess)
    return check_launch();
  return MF_OK;
}

}  // namespace

int big_cholesky(int dtype, const void* diag, const void* sub, const void* rhs, void* out_diag,
                 void* out_sub, void* out_x, void* out_logdet, int32_t* info, int64_t B, int64_t T,
                 int64_t D, cudaStream_t s) {
  // 9 <= D <= 17: half a warp per chain, parallel in time for few long chains (btd_big2.cuh);
  // tuning knob 7 = 2 keeps the one-warp-per-chain kernel below (A/B measurements)
  if (D <= 17 && tuning(7) != 2)
    return big2_cholesky(dtype, diag, sub, rhs, out_diag, out_sub, out_x, out_logdet, info, B, T, D, s);
  return dispatch_big(dtype, D, [&](auto tt, auto dd) {
    using Tp = typename decltype(tt)::type;
    constexpr int kD = decltype(dd)::value;
    auto kern = btd_chol_big_kernel<Tp, kD>;
    constexpr size_t smem = BigCfg<Tp, kD>::SMEM_BYTES;
    int rc = set_smem(kern, smem);
    if (rc != MF_OK) return rc;
    kern<<<grid_for(B, kBigWarps), 32 * kBigWarps, smem, s>>>(
        (const Tp*)diag, (const Tp*)sub, (const Tp*)rhs, (Tp*)out_diag, (Tp*)out_sub, (Tp*)out_x,
        (Tp*)out_logdet, info, B, T);
    return check_launch();
  });
}

int big_solve(int dtype, const void* ld, const void* ls, const void* rhs, void* out, int64_t n_rhs,
              int64_t Bm, int64_t T, int64_t D, int transpose, cudaStream_t s) {
  return dispatch_big(dtype, D, [&](auto tt, auto dd) {
    using Tp = typename decltype(tt)::type;
    constexpr int kD = decltype(dd)::value;
    auto kern = btd_solve_big_kernel<Tp, kD>;
    constexpr size_t smem = BigCfg<Tp, kD>::SMEM_BYTES;
    int rc = set_smem(kern, smem);
    if (rc != MF_OK) return rc;
    kern<<<grid_for(n_rhs, kBigWarps), 32 * kBigWarps, smem, s>>>(
        (const Tp*)ld, (const Tp*)ls, (const Tp*)rhs, (Tp*)out, n_rhs, Bm, T, transpose);
    return check_launch();
  });
}

}  // namespace mf
