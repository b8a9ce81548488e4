// Library context: one CUDA device, options, error string, device scratch arena.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/svscope_b200.h"

struct svs_ctx {
  int device = 0;
  std::string error;
  // options
  int poa_threads = 384;   // 384 threads x 8 columns: one resident window per SM, 12 warps on one alignment
  int prune = 1;       // exact score-bound pruning of DP cells (persistent kernel)
  int poa_cols = 8;    // read columns per thread (16 only with 256 threads)
  int ring_rows = 8;
  int dp_kernel = 2;   // 2: warp-pipelined DP (no CTA barrier per row), 1: barrier-per-row DP
  int workers = 4;
  int64_t arena_mb = 0;
  int lane_jobs = 0;   // alignments per round of a lane (0 = derived)
  int inflight = 0;    // (unused by the round scheduler)
  int streams = 0;     // concurrent round streams (0 = 2)
  // device arena shared by the batched calls (allocated lazily, reused)
  int* slot_flags = nullptr;   // busy flags of the per-SM scratch slots (two resident CTAs per SM)
  void* arena = nullptr;
  size_t arena_bytes = 0;
  int sm_count = 0;
  int n_smid = 0;      // upper bound of %smid
  std::mutex mu;
  std::mutex mu_ed;   // edit-distance calls (they may run on a side thread next to the other stages)
};

struct svs_reads {
  svs_ctx* ctx = nullptr;
  std::vector<uint8_t> host;     // concatenated sequences
  std::vector<int64_t> off;      // n+1
  uint8_t* dev = nullptr;        // device copy of `host`
  int64_t* dev_off = nullptr;    // device copy of `off`
  int64_t n = 0;
};

namespace svs {

inline int fail(svs_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->error = msg;
  return code;
}

#define SVS_CUDA(ctx, expr)                                                                  \
  do {                                                                                       \
    cudaError_t err__ = (expr);                                                              \
    if (err__ != cudaSuccess) {                                                              \
      char buf__[512];                                                                       \
      std::snprintf(buf__, sizeof(buf__), "%s failed: %s (%s:%d)", #expr,                    \
                    cudaGetErrorString(err__), __FILE__, __LINE__);                          \
      return ::svs::fail((ctx), SVS_ERR_CUDA, buf__);                                        \
    }                                                                                        \
  } while (0)

int ensure_arena(svs_ctx* ctx);

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Copies on the calling thread's own stream, complete on return.  The per-thread stream does not
// synchronise with the streams of the window kernel, and neither cudaFree nor a device-wide
// synchronise is used by the auxiliary stages: both would wait for a running window kernel.
inline cudaError_t svs_memcpy_pt(void* dst, const void* src, size_t n, cudaMemcpyKind kind) {
  cudaError_t e = cudaMemcpyAsync(dst, src, n, kind, cudaStreamPerThread);
  return e != cudaSuccess ? e : cudaStreamSynchronize(cudaStreamPerThread);
}

}  // namespace svs
