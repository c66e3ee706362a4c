// Host-buffer entry points (mf_host_*, include/markovflow_b200.h): what a host-resident framework --
// the reference runs TensorFlow on CPU tensors -- binds.  The arrays live in HOST memory; the library
// cuts the work into chunks and pipelines host->device copies, the CUDA sweeps and device->host
// copies on three internal streams (PCIe is full duplex), through device staging slots that are kept
// per device between calls.  Calls return when the results are in host memory.
//
//   mf_host_btd_cholesky        chunks of chains                                (config 2 end to end)
//   mf_host_kalman_log_likelihood   many series: chunks of series; few long series: chunks of TIME,
//                               each reduced on the device to one scan element (mf_kalman_segment_summary)
//                               and joined at the end (mf_kalman_fold_elements) -- 104 B per step go in,
//                               one scalar per series comes out                 (config 3 end to end)
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <mutex>

#include "dispatch.cuh"

using namespace mf;

namespace {

constexpr int kSlots = 3;
constexpr int kMaxDev = 64;

struct HostCtx {
  bool ready = false;
  cudaStream_t h2d = nullptr, comp = nullptr, d2h = nullptr;
  cudaEvent_t in_done[kSlots], comp_done[kSlots], out_done[kSlots];
  bool out_pending[kSlots];
  char* slot[kSlots] = {nullptr, nullptr, nullptr};
  size_t cap = 0;
  std::mutex mu;
};

HostCtx g_ctx[kMaxDev];

#define MF_CU(call)                          \
  do {                                       \
    cudaError_t e_ = (call);                 \
    if (e_ != cudaSuccess) {                 \
      set_last_error(cudaGetErrorString(e_)); \
      return MF_ERR_CUDA;                    \
    }                                        \
  } while (0)

int ctx_prepare(HostCtx& c, size_t slot_bytes) {
  if (!c.ready) {
    MF_CU(cudaStreamCreateWithFlags(&c.h2d, cudaStreamNonBlocking));
    MF_CU(cudaStreamCreateWithFlags(&c.comp, cudaStreamNonBlocking));
    MF_CU(cudaStreamCreateWithFlags(&c.d2h, cudaStreamNonBlocking));
    for (int i = 0; i < kSlots; ++i) {
      MF_CU(cudaEventCreateWithFlags(&c.in_done[i], cudaEventDisableTiming));
      MF_CU(cudaEventCreateWithFlags(&c.comp_done[i], cudaEventDisableTiming));
      MF_CU(cudaEventCreateWithFlags(&c.out_done[i], cudaEventDisableTiming));
    }
    c.ready = true;
  }
  if (slot_bytes > c.cap) {
    for (int i = 0; i < kSlots; ++i) {
      if (c.slot[i]) MF_CU(cudaFree(c.slot[i]));
      c.slot[i] = nullptr;
    }
    c.cap = 0;
    for (int i = 0; i < kSlots; ++i) MF_CU(cudaMalloc((void**)&c.slot[i], slot_bytes));
    c.cap = slot_bytes;
  }
  for (int i = 0; i < kSlots; ++i) c.out_pending[i] = false;
  return MF_OK;
}

size_t up256(size_t v) { return (v + 255) & ~size_t(255); }

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (dev >= 0) {
      ok = cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess;
    }
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

}  // namespace

extern "C" {

int mf_host_pin(void* ptr, size_t bytes) {
  if (!ptr || !bytes) return MF_ERR_BAD_ARG;
  MF_CU(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
  return MF_OK;
}

int mf_host_unpin(void* ptr) {
  if (!ptr) return MF_ERR_BAD_ARG;
  MF_CU(cudaHostUnregister(ptr));
  return MF_OK;
}

int mf_host_release(int device) {
  DeviceGuard guard(device);
  if (!guard.ok) return MF_ERR_CUDA;
  int dev = 0;
  MF_CU(cudaGetDevice(&dev));
  if (dev >= kMaxDev) return MF_ERR_BAD_ARG;
  HostCtx& c = g_ctx[dev];
  std::lock_guard<std::mutex> lock(c.mu);
  for (int i = 0; i < kSlots; ++i) {
    if (c.slot[i]) cudaFree(c.slot[i]);
    c.slot[i] = nullptr;
  }
  c.cap = 0;
  return MF_OK;
}

int mf_host_btd_cholesky(int dtype, const void* diag, const void* sub, const void* rhs, void* out_diag,
                         void* out_sub, void* out_x, void* out_logdet, int32_t* info, int64_t B,
                         int64_t T, int64_t D, int64_t chunk, int device, int64_t* bytes_moved) {
  if (B < 0 || T < 1 || D < 1) return MF_ERR_BAD_ARG;
  if (dtype != MF_F32 && dtype != MF_F64) return MF_ERR_BAD_ARG;
  if (bytes_moved) bytes_moved[0] = bytes_moved[1] = 0;
  if (B == 0) return MF_OK;
  if (!diag || !out_diag || !info) return MF_ERR_BAD_ARG;
  if (T == 1) { sub = nullptr; out_sub = nullptr; }
  if ((sub != nullptr) != (out_sub != nullptr) || (rhs != nullptr) != (out_x != nullptr)) return MF_ERR_BAD_ARG;
  DeviceGuard guard(device);
  if (!guard.ok) return MF_ERR_CUDA;
  int dev = 0;
  MF_CU(cudaGetDevice(&dev));
  if (dev >= kMaxDev) return MF_ERR_BAD_ARG;
  const size_t es = dtype == MF_F64 ? 8 : 4;
  if (chunk <= 0) chunk = 128;
  if (chunk > B) chunk = B;
  const size_t nd = (size_t)T * D * D * es, ns = sub ? (size_t)(T - 1) * D * D * es : 0,
               nr = rhs ? (size_t)T * D * es : 0;
  // slot layout: diag | sub | rhs | ld | ls | x | logdet | info
  const size_t o_d = 0, o_s = up256(o_d + chunk * nd), o_r = up256(o_s + chunk * ns),
               o_ld = up256(o_r + chunk * nr), o_ls = up256(o_ld + chunk * nd),
               o_x = up256(o_ls + chunk * ns), o_lg = up256(o_x + chunk * nr),
               o_info = up256(o_lg + chunk * es), total = up256(o_info + chunk * 4);
  HostCtx& c = g_ctx[dev];
  std::lock_guard<std::mutex> lock(c.mu);
  int rc = ctx_prepare(c, total);
  if (rc != MF_OK) return rc;
  int64_t h2d_bytes = 0, d2h_bytes = 0;
  const int64_t nchunks = (B + chunk - 1) / chunk;
  for (int64_t q = 0; q < nchunks; ++q) {
    const int64_t b0 = q * chunk, nb = (B - b0 < chunk) ? B - b0 : chunk;
    const int k = (int)(q % kSlots);
    char* s = c.slot[k];
    if (c.out_pending[k]) MF_CU(cudaStreamWaitEvent(c.h2d, c.out_done[k], 0));  // slot drained
    MF_CU(cudaMemcpyAsync(s + o_d, (const char*)diag + b0 * nd, nb * nd, cudaMemcpyHostToDevice, c.h2d));
    if (sub) MF_CU(cudaMemcpyAsync(s + o_s, (const char*)sub + b0 * ns, nb * ns, cudaMemcpyHostToDevice, c.h2d));
    if (rhs) MF_CU(cudaMemcpyAsync(s + o_r, (const char*)rhs + b0 * nr, nb * nr, cudaMemcpyHostToDevice, c.h2d));
    h2d_bytes += nb * (int64_t)(nd + ns + nr);
    MF_CU(cudaEventRecord(c.in_done[k], c.h2d));
    MF_CU(cudaStreamWaitEvent(c.comp, c.in_done[k], 0));
    rc = mf_btd_cholesky(dtype, s + o_d, sub ? s + o_s : nullptr, rhs ? s + o_r : nullptr, s + o_ld,
                         sub ? s + o_ls : nullptr, rhs ? s + o_x : nullptr, out_logdet ? s + o_lg : nullptr,
                         (int32_t*)(s + o_info), nb, T, D, (void*)c.comp);
    if (rc != MF_OK) return rc;
    MF_CU(cudaEventRecord(c.comp_done[k], c.comp));
    MF_CU(cudaStreamWaitEvent(c.d2h, c.comp_done[k], 0));
    MF_CU(cudaMemcpyAsync((char*)out_diag + b0 * nd, s + o_ld, nb * nd, cudaMemcpyDeviceToHost, c.d2h));
    if (sub) MF_CU(cudaMemcpyAsync((char*)out_sub + b0 * ns, s + o_ls, nb * ns, cudaMemcpyDeviceToHost, c.d2h));
    if (rhs) MF_CU(cudaMemcpyAsync((char*)out_x + b0 * nr, s + o_x, nb * nr, cudaMemcpyDeviceToHost, c.d2h));
    if (out_logdet)
      MF_CU(cudaMemcpyAsync((char*)out_logdet + b0 * es, s + o_lg, nb * es, cudaMemcpyDeviceToHost, c.d2h));
    MF_CU(cudaMemcpyAsync(info + b0, s + o_info, nb * 4, cudaMemcpyDeviceToHost, c.d2h));
    d2h_bytes += nb * (int64_t)(nd + ns + nr + 4 + (out_logdet ? es : 0));
    MF_CU(cudaEventRecord(c.out_done[k], c.d2h));
    c.out_pending[k] = true;
  }
  MF_CU(cudaStreamSynchronize(c.d2h));
  if (bytes_moved) { bytes_moved[0] = h2d_bytes; bytes_moved[1] = d2h_bytes; }
  return MF_OK;
}

int mf_host_kalman_log_likelihood(int dtype, const void* mu0, const void* chol_p0, const void* a,
                                  const void* b, const void* chol_q, const void* h, const void* obs,
                                  const void* chol_r, void* out, int64_t B, int64_t T, int64_t D, int64_t m,
                                  int64_t h_batch, int64_t r_steps, int64_t chunk_steps, int device,
                                  int64_t* bytes_moved) {
  if (B < 0 || T < 1 || D < 1 || m < 1) return MF_ERR_BAD_ARG;
  if (dtype != MF_F32 && dtype != MF_F64) return MF_ERR_BAD_ARG;
  if (bytes_moved) bytes_moved[0] = bytes_moved[1] = 0;
  if (B == 0) return MF_OK;
  if (!mu0 || !chol_p0 || !h || !obs || !chol_r || !out || (T > 1 && (!a || !b || !chol_q))) return MF_ERR_BAD_ARG;
  if ((h_batch != 1 && h_batch != B) || (r_steps != 1 && r_steps != T)) return MF_ERR_BAD_ARG;
  if (B > 65535) return MF_ERR_UNSUPPORTED;
  DeviceGuard guard(device);
  if (!guard.ok) return MF_ERR_CUDA;
  int dev = 0;
  MF_CU(cudaGetDevice(&dev));
  if (dev >= kMaxDev) return MF_ERR_BAD_ARG;
  const size_t es = dtype == MF_F64 ? 8 : 4;
  // chunks of time: every chunk of every series becomes one scan element on the device
  if (chunk_steps <= 0) {
    // ~48 MB of parameters per chunk: long enough to saturate the sweep, short enough to pipeline
    chunk_steps = (int64_t)(48u << 20) / (int64_t)((2 * D * D + D + m * D + m) * es * B);
    if (chunk_steps < 4096) chunk_steps = 4096;
  }
  if (chunk_steps > T) chunk_steps = T;
  const int64_t nchunks = (T + chunk_steps - 1) / chunk_steps;
  const int64_t NE = 3 * D * D + 2 * D + 1;
  const int64_t L = chunk_steps;
  size_t ws_bytes = mf_kalman_workspace_bytes(dtype, B, L, D);
  {
    const int64_t last = T - (nchunks - 1) * L;  // the ragged last chunk plans its own segments
    const size_t w2 = mf_kalman_workspace_bytes(dtype, B, last, D);
    if (w2 > ws_bytes) ws_bytes = w2;
  }
  // slot layout (per chunk of L steps, all B series): a | b | chol_q | h | obs | chol_r | workspace
  const size_t o_a = 0, o_b = up256(o_a + (size_t)B * L * D * D * es), o_q = up256(o_b + (size_t)B * L * D * es),
               o_h = up256(o_q + (size_t)B * L * D * D * es), o_y = up256(o_h + (size_t)h_batch * L * m * D * es),
               o_r = up256(o_y + (size_t)B * L * m * es), o_ws = up256(o_r + (size_t)(r_steps == 1 ? 1 : L) * m * m * es),
               total = up256(o_ws + ws_bytes);
  HostCtx& c = g_ctx[dev];
  std::lock_guard<std::mutex> lock(c.mu);
  int rc = ctx_prepare(c, total);
  if (rc != MF_OK) return rc;
  // small persistent device arrays: prior, elements, result
  char* small = nullptr;
  const size_t o_mu = 0, o_p0 = up256((size_t)B * D * es), o_el = up256(o_p0 + (size_t)B * D * D * es),
               o_out = up256(o_el + (size_t)nchunks * B * NE * es), small_total = up256(o_out + (size_t)B * NE * es);
  MF_CU(cudaMallocAsync((void**)&small, small_total, c.h2d));
  MF_CU(cudaMemcpyAsync(small + o_mu, mu0, (size_t)B * D * es, cudaMemcpyHostToDevice, c.h2d));
  MF_CU(cudaMemcpyAsync(small + o_p0, chol_p0, (size_t)B * D * D * es, cudaMemcpyHostToDevice, c.h2d));
  int64_t h2d_bytes = (int64_t)((size_t)B * (D + D * D) * es);
  auto copy2d = [&](char* dst, const void* src, int64_t rows, size_t row_elems, int64_t k0, int64_t nk,
                    size_t rec) -> cudaError_t {
    // rows series, each `row_elems` records of `rec` bytes; take records [k0, k0+nk) of every series
    h2d_bytes += rows * nk * (int64_t)rec;
    return cudaMemcpy2DAsync(dst, (size_t)nk * rec, (const char*)src + (size_t)k0 * rec, row_elems * rec,
                             (size_t)nk * rec, (size_t)rows, cudaMemcpyHostToDevice, c.h2d);
  };
  for (int64_t q = 0; q < nchunks; ++q) {
    const int64_t lo = q * L, hi = (lo + L < T) ? lo + L : T, nl = hi - lo;
    const int first = q == 0;
    const int64_t tlo = first ? 0 : lo - 1, nt = first ? nl - 1 : nl;  // transitions leading into the steps
    const int k = (int)(q % kSlots);
    char* s = c.slot[k];
    if (c.out_pending[k]) MF_CU(cudaStreamWaitEvent(c.h2d, c.out_done[k], 0));
    if (nt > 0) {
      MF_CU(copy2d(s + o_a, a, B, (size_t)(T - 1), tlo, nt, (size_t)D * D * es));
      MF_CU(copy2d(s + o_b, b, B, (size_t)(T - 1), tlo, nt, (size_t)D * es));
      MF_CU(copy2d(s + o_q, chol_q, B, (size_t)(T - 1), tlo, nt, (size_t)D * D * es));
    }
    MF_CU(copy2d(s + o_h, h, h_batch, (size_t)T, lo, nl, (size_t)m * D * es));
    MF_CU(copy2d(s + o_y, obs, B, (size_t)T, lo, nl, (size_t)m * es));
    if (r_steps == 1) {
      if (q < kSlots) {
        MF_CU(cudaMemcpyAsync(s + o_r, chol_r, (size_t)m * m * es, cudaMemcpyHostToDevice, c.h2d));
        h2d_bytes += (int64_t)(m * m * es);
      }
    } else {
      MF_CU(copy2d(s + o_r, chol_r, 1, (size_t)T, lo, nl, (size_t)m * m * es));
    }
    MF_CU(cudaEventRecord(c.in_done[k], c.h2d));
    MF_CU(cudaStreamWaitEvent(c.comp, c.in_done[k], 0));
    rc = mf_kalman_segment_summary(dtype, first ? small + o_mu : nullptr, first ? small + o_p0 : nullptr, s + o_a,
                                   s + o_b, s + o_q, s + o_h, s + o_y, s + o_r,
                                   small + o_el + (size_t)q * B * NE * es, B, nl, D, m, h_batch,
                                   r_steps == 1 ? 1 : nl, first, s + o_ws, ws_bytes, (void*)c.comp);
    if (rc != MF_OK) { cudaFreeAsync(small, c.comp); return rc; }
    MF_CU(cudaEventRecord(c.out_done[k], c.comp));  // slot free once the summary has run
    c.out_pending[k] = true;
  }
  rc = mf_kalman_fold_elements(dtype, small + o_el, small + o_out, nchunks, B, D, (void*)c.comp);
  if (rc != MF_OK) { cudaFreeAsync(small, c.comp); return rc; }
  // the log-likelihood is the last component of the joined element
  MF_CU(cudaMemcpy2DAsync(out, es, small + o_out + (size_t)(NE - 1) * es, (size_t)NE * es, es, (size_t)B,
                          cudaMemcpyDeviceToHost, c.comp));
  MF_CU(cudaFreeAsync(small, c.comp));
  MF_CU(cudaStreamSynchronize(c.comp));
  if (bytes_moved) { bytes_moved[0] = h2d_bytes; bytes_moved[1] = (int64_t)((size_t)B * es); }
  return MF_OK;
}

}  // extern "C"
