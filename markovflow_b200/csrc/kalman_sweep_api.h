// Internal (C++) interface between capi_kalman.cu and capi_kalman_sweep.cu: the TMA-sweep
// implementation of the m = 1 Kalman log-likelihood for D <= kKalmanSweepMaxD.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace mf {

constexpr int kKalmanSweepMaxD = 4;

struct KalmanRawArgs {
  int dtype;
  const void *mu0, *chol_p0, *a, *b, *chol_q, *h, *obs, *chol_r;
  int64_t B, T, D, h_batch, r_steps;
  int first_is_initial;
};

// virtual chains resident per wave for state dim D (chains per CTA x 148 SMs)
int kalman_sweep_chains_per_cta(int64_t D);
// scan-block width used for state dim D
int kalman_sweep_scan_threads(int64_t D);

// mode 0: plain filter, P == 1 (out [B])
// mode 1: summaries     (out = elems [B,P,N]; with reduce_warp != 0 and P a multiple of 32:
//                        out = [B,P/32,N], the joins of 32 consecutive segments)
// mode 2: seeded filter (out = partial [B,P]; needs local/block prefix)
int kalman_sweep_launch(int mode, const KalmanRawArgs& g, int64_t P, int64_t L,
                        const void* local_prefix, const void* block_prefix, int64_t nblk,
                        int have_prefix, void* out, cudaStream_t s, int reduce_warp = 0);

// in-place block scan of elems [B,P,N] -> local exclusive prefixes, block_agg [B,nblk,N]
int kalman_sweep_block_scan(int dtype, int64_t D, void* elems, void* block_agg, int64_t B,
                            int64_t P, int64_t nblk, cudaStream_t s);
// block_prefix [B,nblk,N] (or NULL) from block_agg (+ prefix_in [B,N] or NULL); total_out [B,N] or
// NULL; ell_out [B] or NULL (log-likelihood of the joined series)
int kalman_sweep_top_scan(int dtype, int64_t D, const void* block_agg, const void* prefix_in,
                          void* block_prefix, void* total_out, void* ell_out, int64_t B,
                          int64_t nblk, cudaStream_t s);

// peer-mapped exchange regions of a time-sharded evaluation (kalman_sweep.cuh: PeerExchange)
struct KalmanPeerArgs {
  void* region[8];
  unsigned long long epoch;
  int rank, world;
};

// ordered reduction of elems [B,P,N]: total_out [B,N] or NULL, ell_out [B] or NULL; with `peers` the chain
// totals are exchanged with the other ranks and joined in rank order in the same launch
int kalman_sweep_reduce(int dtype, int64_t D, const void* elems, void* total_out, void* ell_out,
                        int64_t B, int64_t P, cudaStream_t s, const KalmanPeerArgs* peers = nullptr);

}  // namespace mf
