// MSA post-processing on the device (svs_msa_features, include/svscope_b200.h):
//   column statistics + feature mask  = FindNonSameSite  (src/DataScanner.py:167-179, 217-219)
//   ZeroParamNum                      = EMCluster        (src/ReadsCluster.py:226-234)
//   read-by-read identity counts      = pariwiseDistance (src/ReadsCluster.py:44-59)
// Byte work, bound by HBM/L2 reads of the encoded MSA; thread per column (coalesced over
// columns) for the statistics, warp per read pair (lanes stride the columns) for identities.
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

#include "context.h"

namespace svs {
namespace {

struct FeatTask {
  const int8_t* enc;   // n_rows x n_cols
  const uint8_t* drop; // n_cols
  uint8_t* keep;       // n_cols
  int32_t* ident;      // n_rows x n_rows
  int32_t* counters;   // [2] nf, zero_params
  int32_t n_rows, n_cols;
  double cutoff;
};

__global__ void column_stats_kernel(const FeatTask* __restrict__ tasks) {
  const FeatTask t = tasks[blockIdx.y];
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= t.n_cols) return;
  int cnt[5] = {0, 0, 0, 0, 0};
  for (int r = 0; r < t.n_rows; ++r) {
    const int v = t.enc[static_cast<size_t>(r) * t.n_cols + col];
#pragma unroll
    for (int a = 0; a < 5; ++a) cnt[a] += (v == a);
  }
  // second largest of the five counts
  int hi = cnt[0], second = INT32_MIN;
#pragma unroll
  for (int a = 1; a < 5; ++a) {
    if (cnt[a] > hi) { second = hi; hi = cnt[a]; }
    else if (cnt[a] > second) second = cnt[a];
  }
  const bool kept = !t.drop[col] && static_cast<double>(second) >= t.cutoff;
  t.keep[col] = kept ? 1 : 0;
  if (kept) {
    int zeros = 0;
#pragma unroll
    for (int a = 0; a < 5; ++a) zeros += (cnt[a] == 0);
    atomicAdd(&t.counters[0], 1);
    if (zeros) atomicAdd(&t.counters[1], zeros);
  }
}

// grid (max_rows, n_windows); warp w of the block handles pairs (i, j) with j = w, w+nw, ... < i
__global__ void identity_kernel(const FeatTask* __restrict__ tasks) {
  const FeatTask t = tasks[blockIdx.y];
  const int i = blockIdx.x;
  if (i >= t.n_rows) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int8_t* ri = t.enc + static_cast<size_t>(i) * t.n_cols;
  for (int j = warp; j < i; j += nw) {
    const int8_t* rj = t.enc + static_cast<size_t>(j) * t.n_cols;
    int same = 0;
    for (int c = lane; c < t.n_cols; c += 32) same += (t.keep[c] && ri[c] == rj[c]);
    same = __reduce_add_sync(0xffffffffu, same);
    if (lane == 0) {
      t.ident[static_cast<size_t>(i) * t.n_rows + j] = same;
      t.ident[static_cast<size_t>(j) * t.n_rows + i] = same;
    }
  }
  if (threadIdx.x == 0) t.ident[static_cast<size_t>(i) * t.n_rows + i] = -1;  // diagonal: caller sets 1.0
}

}  // namespace
}  // namespace svs

using namespace svs;

extern "C" int svs_msa_features(svs_ctx* ctx, int64_t n_windows, const int8_t* enc, const int64_t* enc_off,
                                const int32_t* n_rows, const int32_t* n_cols, const uint8_t* drop,
                                const int64_t* col_off, const double* cutoff, uint8_t* keep, int32_t* nf,
                                int32_t* zero_params, int32_t* ident, const int64_t* ident_off) {
  if (!ctx || n_windows < 0) return fail(ctx, SVS_ERR_ARG, "null argument");
  if (n_windows == 0) return SVS_OK;
  std::lock_guard<std::mutex> lock(ctx->mu);
  SVS_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t enc_bytes = static_cast<size_t>(enc_off[n_windows]);
  const size_t col_total = static_cast<size_t>(col_off[n_windows]);
  const size_t ident_total = static_cast<size_t>(ident_off[n_windows]);
  int8_t* d_enc = nullptr; uint8_t *d_drop = nullptr, *d_keep = nullptr;
  int32_t *d_ident = nullptr, *d_cnt = nullptr; FeatTask* d_tasks = nullptr;
  auto cleanup = [&]() {
    ((d_enc) ? cudaFreeAsync(d_enc, cudaStreamPerThread) : cudaSuccess); ((d_drop) ? cudaFreeAsync(d_drop, cudaStreamPerThread) : cudaSuccess); ((d_keep) ? cudaFreeAsync(d_keep, cudaStreamPerThread) : cudaSuccess); ((d_ident) ? cudaFreeAsync(d_ident, cudaStreamPerThread) : cudaSuccess); ((d_cnt) ? cudaFreeAsync(d_cnt, cudaStreamPerThread) : cudaSuccess); ((d_tasks) ? cudaFreeAsync(d_tasks, cudaStreamPerThread) : cudaSuccess);
  };
#define SVS_CU(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { cleanup(); \
    return fail(ctx, SVS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); } } while (0)
  SVS_CU(cudaMallocAsync(reinterpret_cast<void**>(&d_enc), enc_bytes + 16, cudaStreamPerThread));
  SVS_CU(cudaMallocAsync(reinterpret_cast<void**>(&d_drop), col_total + 16, cudaStreamPerThread));
  SVS_CU(cudaMallocAsync(reinterpret_cast<void**>(&d_keep), col_total + 16, cudaStreamPerThread));
  SVS_CU(cudaMallocAsync(reinterpret_cast<void**>(&d_ident), (ident_total + 4) * sizeof(int32_t), cudaStreamPerThread));
  SVS_CU(cudaMallocAsync(reinterpret_cast<void**>(&d_cnt), 2 * n_windows * sizeof(int32_t), cudaStreamPerThread));
  SVS_CU(cudaMallocAsync(reinterpret_cast<void**>(&d_tasks), n_windows * sizeof(FeatTask), cudaStreamPerThread));
  SVS_CU(svs_memcpy_pt(d_enc, enc, enc_bytes, cudaMemcpyHostToDevice));
  SVS_CU(svs_memcpy_pt(d_drop, drop, col_total, cudaMemcpyHostToDevice));
  SVS_CU(cudaMemsetAsync(d_cnt, 0, 2 * n_windows * sizeof(int32_t), cudaStreamPerThread));
  std::vector<FeatTask> tasks(n_windows);
  int max_rows = 0, max_cols = 0;
  for (int64_t w = 0; w < n_windows; ++w) {
    FeatTask& t = tasks[w];
    t.enc = d_enc + enc_off[w];
    t.drop = d_drop + col_off[w];
    t.keep = d_keep + col_off[w];
    t.ident = d_ident + ident_off[w];
    t.counters = d_cnt + 2 * w;
    t.n_rows = n_rows[w];
    t.n_cols = n_cols[w];
    t.cutoff = cutoff[w];
    max_rows = std::max(max_rows, t.n_rows);
    max_cols = std::max(max_cols, t.n_cols);
  }
  SVS_CU(svs_memcpy_pt(d_tasks, tasks.data(), n_windows * sizeof(FeatTask), cudaMemcpyHostToDevice));
  if (max_cols > 0 && max_rows > 0) {
    const dim3 g1((max_cols + 255) / 256, static_cast<unsigned>(n_windows));
    column_stats_kernel<<<g1, 256, 0, cudaStreamPerThread>>>(d_tasks);
    SVS_CU(cudaGetLastError());
    const dim3 g2(max_rows, static_cast<unsigned>(n_windows));
    identity_kernel<<<g2, 256, 0, cudaStreamPerThread>>>(d_tasks);
    SVS_CU(cudaGetLastError());
  } else {
    SVS_CU(cudaMemsetAsync(d_keep, 0, col_total + 16, cudaStreamPerThread));
  }
  SVS_CU(cudaStreamSynchronize(cudaStreamPerThread));
  if (col_total) SVS_CU(svs_memcpy_pt(keep, d_keep, col_total, cudaMemcpyDeviceToHost));
  if (ident_total) SVS_CU(svs_memcpy_pt(ident, d_ident, ident_total * sizeof(int32_t), cudaMemcpyDeviceToHost));
  std::vector<int32_t> cnt(2 * n_windows);
  SVS_CU(svs_memcpy_pt(cnt.data(), d_cnt, cnt.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
  for (int64_t w = 0; w < n_windows; ++w) {
    nf[w] = cnt[2 * w];
    zero_params[w] = cnt[2 * w + 1];
  }
#undef SVS_CU
  cleanup();
  return SVS_OK;
}
