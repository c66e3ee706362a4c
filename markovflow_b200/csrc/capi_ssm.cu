// C-ABI entry points for the StateSpaceModel operators (include/markovflow_b200.h).
#include "dispatch.cuh"
#include "mid_api.h"
#include "ssm_kernels.cuh"
#include "ssm_sweep_api.h"

using namespace mf;

extern "C" {

int mf_ssm_build_precision(int dtype, const void* chol_p0, const void* a, const void* chol_q,
                           const void* h, const void* r_inv, void* out_diag, void* out_sub,
                           int64_t B, int64_t T, int64_t D, int64_t m, int64_t h_batch,
                           int64_t r_steps, void* stream) {
  if (B < 0 || T < 1 || D < 1) return MF_ERR_BAD_ARG;
  if (B == 0) return MF_OK;
  if (!chol_p0 || !out_diag) return MF_ERR_BAD_ARG;
  if (T > 1 && (!a || !chol_q || !out_sub)) return MF_ERR_BAD_ARG;
  if (h && (!r_inv || m < 1 || (h_batch != 1 && h_batch != B) || (r_steps != 1 && r_steps != T)))
    return MF_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  if (mid_dim(D))
    return mid_build_precision(dtype, chol_p0, a, chol_q, h, r_inv, out_diag, out_sub, B, T, D, m, h_batch, r_steps, s);
  return dispatch_small(dtype, D, [&](auto tt, auto dd) {
    using Tp = typename decltype(tt)::type;
    constexpr int kD = decltype(dd)::value;
    ssm_build_precision_kernel<Tp, kD><<<grid_for(B * T, 128), 128, 0, s>>>(
        (const Tp*)chol_p0, (const Tp*)a, (const Tp*)chol_q, (const Tp*)h, (const Tp*)r_inv,
        (Tp*)out_diag, (Tp*)out_sub, B, T, (int)m, h_batch, r_steps);
    return check_launch();
  });
}

int mf_ssm_affine_scan(int dtype, const void* mu0, const void* chol_p0, const void* a,
                       const void* b, const void* chol_q, const void* eps, void* out, int64_t n,
                       int64_t Bm, int64_t T, int64_t D, void* stream) {
  if (n < 0 || Bm < 1 || T < 1 || D < 1) return MF_ERR_BAD_ARG;
  if (n == 0) return MF_OK;
  if (!mu0 || !out || (T > 1 && (!a || !b))) return MF_ERR_BAD_ARG;
  if (eps && (!chol_p0 || (T > 1 && !chol_q))) return MF_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  if (mid_dim(D)) return mid_affine_scan(dtype, mu0, chol_p0, a, b, chol_q, eps, out, n, Bm, T, D, s);
  if (D <= kSsmSweepMaxD && T > 1 && tuning(4) != 1) {
    const int rc = ssm_sweep_affine(dtype, D, mu0, chol_p0, a, b, chol_q, eps, out, n, Bm, T, s);
    if (rc != MF_ERR_UNSUPPORTED) return rc;
  }
  return dispatch_small(dtype, D, [&](auto tt, auto dd) {
    using Tp = typename decltype(tt)::type;
    constexpr int kD = decltype(dd)::value;
    ssm_affine_scan_kernel<Tp, kD><<<grid_for(n, 32), 32, 0, s>>>(
        (const Tp*)mu0, (const Tp*)chol_p0, (const Tp*)a, (const Tp*)b, (const Tp*)chol_q,
        (const Tp*)eps, (Tp*)out, n, Bm, T, 0, 0ull);
    return check_launch();
  });
}

int mf_ssm_sample(int dtype, const void* mu0, const void* chol_p0, const void* a, const void* b,
                  const void* chol_q, uint64_t seed, void* out, int64_t n, int64_t Bm, int64_t T, int64_t D,
                  void* stream) {
  if (n < 0 || Bm < 1 || T < 1 || D < 1) return MF_ERR_BAD_ARG;
  if (n == 0) return MF_OK;
  if (!mu0 || !chol_p0 || !out || (T > 1 && (!a || !b || !chol_q))) return MF_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  if (mid_dim(D)) {
    // blocks in shared memory: the Philox stream is written to a stream-ordered scratch array first
    const size_t es = dtype == MF_F64 ? 8 : 4;
    void* eps = nullptr;
    if (cudaMallocAsync(&eps, (size_t)n * T * D * es, s) != cudaSuccess) return check_launch();
    int rc = mf_philox_normal(dtype, seed, eps, n, T, D, stream);
    if (rc == MF_OK) rc = mid_affine_scan(dtype, mu0, chol_p0, a, b, chol_q, eps, out, n, Bm, T, D, s);
    cudaFreeAsync(eps, s);
    return rc;
  }
  if (D <= kSsmSweepMaxD && T > 1 && tuning(4) != 1) {
    const int rc = ssm_sweep_affine(dtype, D, mu0, chol_p0, a, b, chol_q, nullptr, out, n, Bm, T, s, 1, seed);
    if (rc != MF_ERR_UNSUPPORTED) return rc;
  }
  return dispatch_small(dtype, D, [&](auto tt, auto dd) {
    using Tp = typename decltype(tt)::type;
    constexpr int kD = decltype(dd)::value;
    ssm_affine_scan_kernel<Tp, kD><<<grid_for(n, 32), 32, 0, s>>>(
        (const Tp*)mu0, (const Tp*)chol_p0, (const Tp*)a, (const Tp*)b, (const Tp*)chol_q, nullptr, (Tp*)out, n,
        Bm, T, 1, (unsigned long long)seed);
    return check_launch();
  });
}

int mf_philox_normal(int dtype, uint64_t seed, void* out, int64_t n, int64_t T, int64_t D, void* stream) {
  if (n < 0 || T < 1 || D < 1) return MF_ERR_BAD_ARG;
  if (n == 0) return MF_OK;
  if (!out) return MF_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  if (mid_dim(D))
    return dispatch_dtype(dtype, [&](auto tt) {
      using Tp = typename decltype(tt)::type;
      philox_normal_dyn_kernel<Tp><<<grid_for(n * T, 128), 128, 0, s>>>((Tp*)out, n, T, (int)D,
                                                                        (unsigned long long)seed);
      return check_launch();
    });
  return dispatch_small(dtype, D, [&](auto tt, auto dd) {
    using Tp = typename decltype(tt)::type;
    constexpr int kD = decltype(dd)::value;
    philox_normal_kernel<Tp, kD><<<grid_for(n * T, 128), 128, 0, s>>>((Tp*)out, n, T, (unsigned long long)seed);
    return check_launch();
  });
}

int mf_ssm_marginals(int dtype, const void* mu0, const void* chol_p0, const void* a, const void* b,
                     const void* chol_q, void* out_mean, void* out_cov, void* out_sub, int64_t B,
                     int64_t T, int64_t D, void* stream) {
  if (B < 0 || T < 1 || D < 1) return MF_ERR_BAD_ARG;
  if (B == 0) return MF_OK;
  if (!mu0 || !chol_p0 || (T > 1 && (!a || !b || !chol_q))) return MF_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  if (mid_dim(D)) return mid_ssm_moments(dtype, 0, mu0, chol_p0, a, b, chol_q, out_mean, out_cov, out_sub, B, T, D, s);
  if (D <= kSsmSweepMaxD && T > 1 && tuning(4) != 1) {
    const int rc = ssm_sweep_moments(dtype, D, 0, mu0, chol_p0, a, b, chol_q, out_mean, out_cov,
                                     out_sub, B, T, s);
    if (rc != MF_ERR_UNSUPPORTED) return rc;
  }
  return dispatch_small(dtype, D, [&](auto tt, auto dd) {
    using Tp = typename decltype(tt)::type;
    constexpr int kD = decltype(dd)::value;
    ssm_marginals_kernel<Tp, kD><<<grid_for(B, 32), 32, 0, s>>>(
        (const Tp*)mu0, (const Tp*)chol_p0, (const Tp*)a, (const Tp*)b, (const Tp*)chol_q,
        (Tp*)out_mean, (Tp*)out_cov, (Tp*)out_sub, B, T);
    return check_launch();
  });
}

int mf_ssm_log_pdf(int dtype, const void* mu0, const void* chol_p0, const void* a, const void* b,
                   const void* chol_q, const void* states, void* out, int64_t n, int64_t Bm,
                   int64_t T, int64_t D, void* stream) {
  if (n < 0 || Bm < 1 || T < 1 || D < 1) return MF_ERR_BAD_ARG;
  if (n == 0) return MF_OK;
  if (!mu0 || !chol_p0 || !states || !out || (T > 1 && (!a || !b || !chol_q)))
    return MF_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  if (mid_dim(D)) return mid_log_pdf(dtype, mu0, chol_p0, a, b, chol_q, states, out, n, Bm, T, D, s);
  return dispatch_small(dtype, D, [&](auto tt, auto dd) {
    using Tp = typename decltype(tt)::type;
    constexpr int kD = decltype(dd)::value;
    const int64_t seg_len = 4096;
    const int64_t nseg = (T + seg_len - 1) / seg_len;
    if (nseg > 1 && cudaMemsetAsync(out, 0, sizeof(Tp) * n, s) != cudaSuccess)
      return check_launch();
    for (int64_t c0 = 0; c0 < n; c0 += 65535) {
      const int64_t nc = (n - c0 < 65535) ? n - c0 : 65535;
      dim3 grid((unsigned)nseg, (unsigned)nc);
      ssm_log_pdf_kernel<Tp, kD><<<grid, 128, 0, s>>>(
          (const Tp*)mu0, (const Tp*)chol_p0, (const Tp*)a, (const Tp*)b, (const Tp*)chol_q,
          (const Tp*)states, (Tp*)out, Bm, T, seg_len, nseg > 1 ? 1 : 0, c0);
    }
    return check_launch();
  });
}

int mf_ssm_kl_divergence(int dtype, const void* q_mu0, const void* q_chol_p0, const void* q_a,
                         const void* q_b, const void* q_chol_q, const void* p_mu0,
                         const void* p_chol_p0, const void* p_a, const void* p_b,
                         const void* p_chol_q, void* out, int64_t B, int64_t T, int64_t D,
                         void* stream) {
  if (B < 0 || T < 1 || D < 1) return MF_ERR_BAD_ARG;
  if (B == 0) return MF_OK;
  if (!q_mu0 || !q_chol_p0 || !p_mu0 || !p_chol_p0 || !out) return MF_ERR_BAD_ARG;
  if (T > 1 && (!q_a || !q_b || !q_chol_q || !p_a || !p_b || !p_chol_q)) return MF_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  if (mid_dim(D))
    return mid_kl_divergence(dtype, q_mu0, q_chol_p0, q_a, q_b, q_chol_q, p_mu0, p_chol_p0, p_a, p_b, p_chol_q, out, B,
                             T, D, s);
  if (D <= kSsmSweepMaxD && T > 1 && tuning(4) != 1) {
    const int rc = ssm_sweep_kl(dtype, D, q_mu0, q_chol_p0, q_a, q_b, q_chol_q, p_mu0, p_chol_p0,
                                p_a, p_b, p_chol_q, out, B, T, s);
    if (rc != MF_ERR_UNSUPPORTED) return rc;
  }
  return dispatch_small(dtype, D, [&](auto tt, auto dd) {
    using Tp = typename decltype(tt)::type;
    constexpr int kD = decltype(dd)::value;
    ssm_kl_kernel<Tp, kD><<<grid_for(B, 32), 32, 0, s>>>(
        (const Tp*)q_mu0, (const Tp*)q_chol_p0, (const Tp*)q_a, (const Tp*)q_b,
        (const Tp*)q_chol_q, (const Tp*)p_mu0, (const Tp*)p_chol_p0, (const Tp*)p_a,
        (const Tp*)p_b, (const Tp*)p_chol_q, (Tp*)out, B, T);
    return check_launch();
  });
}

int mf_block_cholesky_or_zero(int dtype, const void* cov, void* out, int32_t* info, int64_t n,
                              int64_t D, void* stream) {
  if (n < 0 || D < 1) return MF_ERR_BAD_ARG;
  if (n == 0) return MF_OK;
  if (!cov || !out) return MF_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  if (info && cudaMemsetAsync(info, 0, sizeof(int32_t), s) != cudaSuccess) return check_launch();
  if (mid_dim(D)) return mid_block_cholesky_or_zero(dtype, cov, out, info, n, D, s);
  return dispatch_small(dtype, D, [&](auto tt, auto dd) {
    using Tp = typename decltype(tt)::type;
    constexpr int kD = decltype(dd)::value;
    block_cholesky_or_zero_kernel<Tp, kD><<<grid_for(n, 128), 128, 0, s>>>((const Tp*)cov, (Tp*)out,
                                                                          info, n);
    return check_launch();
  });
}

int mf_block_chol_of_inverse(int dtype, const void* chol, void* out, int64_t n, int64_t D,
                             void* stream) {
  if (n < 0 || D < 1) return MF_ERR_BAD_ARG;
  if (n == 0) return MF_OK;
  if (!chol || !out) return MF_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  if (mid_dim(D)) return mid_block_chol_of_inverse(dtype, chol, out, n, D, s);
  return dispatch_small(dtype, D, [&](auto tt, auto dd) {
    using Tp = typename decltype(tt)::type;
    constexpr int kD = decltype(dd)::value;
    block_chol_of_inverse_kernel<Tp, kD><<<grid_for(n, 128), 128, 0, s>>>((const Tp*)chol, (Tp*)out, n);
    return check_launch();
  });
}

}  // extern "C"
