// Batched MisScore alignments (svs_misscore_pairs): one CTA per consensus pair.
//
// Contract: the counts of the first alignment of Bio.pairwise2.align.globalms(a, b, match,
// mismatch, g, g) that the reference's AligmentScore takes its MisScore from
// (src/PairwiseCompare.py:19-30; MisScore = alignment columns - '|' columns), see
// misscore_tb.h for the traversal rule.  Scores are integers, open == extend == g <= 0.
//
// Layout: seqA runs down the rows, seqB along the columns.  A CTA sweeps the rows of a strip of
// at most kThreads*kC columns; every thread owns kC consecutive columns in registers.  The
// left-neighbour dependence is a max-plus prefix scan (misscore_cell.h): warp shuffles, one
// shared-memory hop between warps, one barrier per row.  Per cell three trace bits go to HBM
// as a nibble (one 64-bit store per thread and row, rows contiguous): 0.5 B per cell is the
// traffic of the kernel.  Wider pairs take several strips, the last column of a strip parked
// in global memory.  Thread 0 then walks the trace (one load per alignment column).
// Integer add/max/compare work, ALU-bound; no tensor-core shape in it.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <string>
#include <vector>

#include "context.h"
#include "misscore_cell.h"

namespace svs {
namespace {

constexpr int kThreads = 256;
constexpr int kC = 16;
constexpr int kWarps = kThreads / 32;

struct MisTask {
  const uint8_t* a;
  const uint8_t* b;
  int32_t la, lb;
  uint8_t* trace;     // [la][pitch] nibbles
  int64_t pitch;
  int32_t* bnd;       // [2][la+1] last column of the previous strip (only if more than one strip)
  uint8_t* line;      // [la+lb] match line, or null
  int32_t* result;    // [4] score, columns, matches, status
};

__global__ void __launch_bounds__(kThreads, 2)
misscore_kernel(const MisTask* __restrict__ tasks, const int32_t* __restrict__ order, int s_match, int s_mis,
                int gap) {
  __shared__ int wtot[2][kWarps];
  const MisTask t = tasks[order[blockIdx.x]];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int strip = mis_strip_cols(t.lb, kThreads, kC);
  const int nstrips = (t.lb + strip - 1) / strip;
  int score = 0;
  for (int s = 0; s < nstrips; ++s) {
    const int col0 = s * strip;
    const int j0 = col0 + tid * kC + 1;           // first DP column of this thread
    const bool stores = tid * kC < strip && j0 <= t.lb;
    const bool last_thread = (tid + 1) * kC == strip;
    const int32_t* bin = t.bnd ? t.bnd + static_cast<int64_t>((s & 1) ^ 1) * (t.la + 1) : nullptr;
    int32_t* bout = t.bnd ? t.bnd + static_cast<int64_t>(s & 1) * (t.la + 1) : nullptr;
    uint32_t bw[kC / 4];  // symbols of seqB of this thread's columns, four per word
    int up[kC], loc[kC];
#pragma unroll
    for (int q = 0; q < kC / 4; ++q) bw[q] = 0;
#pragma unroll
    for (int k = 0; k < kC; ++k) {
      const uint32_t ch = (j0 + k <= t.lb) ? t.b[j0 + k - 1] : 0u;  // 0 never equals a sequence symbol
      bw[k / 4] |= ch << (8 * (k % 4));
      up[k] = -gap * (j0 + k);
    }
    int upleft = -gap * (j0 - 1);
    uint8_t* out_row = t.trace + (j0 - 1) / 2;
    const int jl = t.lb - j0;  // index of the last column of seqB inside this thread, if 0 <= jl < kC
    uint8_t a_next = t.a[0];
    int lin0_next = (tid == 0) ? (s == 0 ? -gap : bin[1]) : 0;
    for (int r = 1; r <= t.la; ++r) {
      const uint8_t ar = a_next;
      const int lin0 = lin0_next;
      if (r < t.la) {
        a_next = t.a[r];
        if (tid == 0) lin0_next = s == 0 ? -gap * (r + 1) : bin[r + 1];
      }
      // bit k of eq: seqA symbol of the row equals seqB symbol of column k (byte compare, then the
      // four 0/1 bytes gathered into a nibble by one multiply)
      const uint32_t arw = static_cast<uint32_t>(ar) * 0x01010101u;
      uint32_t eq = 0;
#pragma unroll
      for (int q = 0; q < kC / 4; ++q)
        eq |= ((((__vcmpeq4(bw[q], arw) & 0x01010101u) * 0x01020408u) >> 24) & 0xfu) << (4 * q);
      const int p = mis_pass1<kC>(up, upleft, eq, s_match, s_mis, gap, tid == 0 ? lin0 : kMisNeg, loc);
      // exclusive prefix maximum of x over the threads of the CTA
      const int x = p + gap * kC * (tid + 1);
      int v = x;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v = max(v, n);
      }
      if (lane == 31) wtot[r & 1][warp] = v;
      int excl = __shfl_up_sync(0xffffffffu, v, 1);
      if (lane == 0) excl = kMisNeg;
      __syncthreads();
      const int w = (lane < warp) ? wtot[r & 1][lane] : kMisNeg;
      excl = max(excl, __reduce_max_sync(0xffffffffu, w));
      const int lin = tid == 0 ? lin0 : excl - gap * kC * tid;
      const uint64_t nibs = mis_pass2<kC>(up, upleft, eq, s_match, s_mis, gap, lin, loc);
      if (stores) *reinterpret_cast<uint64_t*>(out_row + static_cast<int64_t>(r - 1) * t.pitch) = nibs;
      if (last_thread && bout) bout[r] = up[kC - 1];
    }
    if (jl >= 0 && jl < kC && tid * kC < strip) {
#pragma unroll
      for (int k = 0; k < kC; ++k)
        if (k == jl) score = up[k];
      t.result[0] = score;
    }
    __syncthreads();  // strip boundary and trace rows visible to the whole CTA
  }
  if (tid == 0) {
    MisTrace T{t.trace, t.pitch};
    int32_t res[2] = {0, 0};
    const int rc = misscore_traceback(T, t.a, t.la, t.b, t.lb, t.line, res);
    t.result[1] = res[0];
    t.result[2] = res[1];
    t.result[3] = rc;
  }
}

}  // namespace
}  // namespace svs

using namespace svs;

extern "C" int svs_misscore_pairs(svs_ctx* ctx, const svs_reads* reads, const int64_t* a, const int64_t* b,
                                  int64_t n_pairs, int match, int mismatch, int open, int extend, int32_t* out,
                                  uint8_t* lines, const int64_t* line_off, double* stats, int n_stats) {
  if (!ctx || !reads || n_pairs < 0 || (n_pairs > 0 && (!a || !b || !out))) return fail(ctx, SVS_ERR_ARG, "null argument");
  if (lines && !line_off) return fail(ctx, SVS_ERR_ARG, "lines without line_off");
  if (open != extend)
    return fail(ctx, SVS_ERR_UNSUPPORTED, "misscore: only equal open and extend gap scores (the reference uses -1, -1)");
  if (open > 0) return fail(ctx, SVS_ERR_ARG, "misscore: gap scores must be non-positive");
  if (std::abs(match) > 1000 || std::abs(mismatch) > 1000 || open < -1000)
    return fail(ctx, SVS_ERR_UNSUPPORTED, "misscore: |score| <= 1000");
  std::lock_guard<std::mutex> lock(ctx->mu);
  if (n_pairs == 0) return SVS_OK;
  SVS_CUDA(ctx, cudaSetDevice(ctx->device));
  const int rc_arena = ensure_arena(ctx);
  if (rc_arena) return rc_arena;

  const int64_t n = n_pairs;
  std::vector<MisTask> tasks(n);
  std::vector<size_t> need(n);
  std::vector<double> cost(n);
  double cells = 0;
  for (int64_t k = 0; k < n; ++k) {
    if (a[k] < 0 || a[k] >= reads->n || b[k] < 0 || b[k] >= reads->n) return fail(ctx, SVS_ERR_ARG, "read index out of range");
    const int64_t la = reads->off[a[k] + 1] - reads->off[a[k]], lb = reads->off[b[k] + 1] - reads->off[b[k]];
    if (la == 0 || lb == 0) return fail(ctx, SVS_ERR_ARG, "misscore: empty sequence (globalms returns no alignment)");
    const int64_t smax = std::max<int64_t>(std::max(std::abs(match), std::abs(mismatch)), std::max(-open, 1));
    if ((la + lb) * smax >= (1 << 27)) return fail(ctx, SVS_ERR_CAPACITY, "misscore: scores of this pair leave the 32-bit working range");
    MisTask& t = tasks[k];
    t.a = reads->dev + reads->off[a[k]];
    t.b = reads->dev + reads->off[b[k]];
    t.la = static_cast<int32_t>(la);
    t.lb = static_cast<int32_t>(lb);
    t.pitch = mis_trace_pitch(t.lb);
    const bool multi = mis_strip_cols(t.lb, kThreads, kC) < t.lb;
    need[k] = align_up(static_cast<size_t>(t.pitch) * la, 256) + (multi ? align_up(2 * (la + 1) * sizeof(int32_t), 256) : 0) +
              (lines ? align_up(static_cast<size_t>(la + lb), 256) : 0);
    cost[k] = static_cast<double>(la) * static_cast<double>(lb);
    cells += cost[k];
  }
  std::vector<int32_t> order(n);
  for (int64_t k = 0; k < n; ++k) order[k] = static_cast<int32_t>(k);
  std::sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { return cost[x] > cost[y]; });

  // fixed part of the arena: task table and launch order
  uint8_t* base = static_cast<uint8_t*>(ctx->arena);
  const size_t fixed = align_up(n * sizeof(MisTask), 256) + align_up(n * sizeof(int32_t), 256) +
                       align_up(n * 4 * sizeof(int32_t), 256);
  if (fixed >= ctx->arena_bytes) return fail(ctx, SVS_ERR_CAPACITY, "misscore: too many pairs for the arena");
  MisTask* d_tasks = reinterpret_cast<MisTask*>(base);
  int32_t* d_order = reinterpret_cast<int32_t*>(base + align_up(n * sizeof(MisTask), 256));
  int32_t* d_results = d_order + align_up(n * sizeof(int32_t), 256) / sizeof(int32_t);
  const size_t room = ctx->arena_bytes - fixed;
  for (int64_t k = 0; k < n; ++k)
    if (need[k] > room) return fail(ctx, SVS_ERR_CAPACITY, "misscore: one pair needs " + std::to_string(need[k] >> 20) + " MiB of trace, more than the arena holds");

  cudaEvent_t e0 = nullptr, e1 = nullptr;
  SVS_CUDA(ctx, cudaEventCreate(&e0));
  SVS_CUDA(ctx, cudaEventCreate(&e1));
  auto done = [&](int rc) { cudaEventDestroy(e0); cudaEventDestroy(e1); return rc; };
#define SVS_CU(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) \
    return done(fail(ctx, SVS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__))); } while (0)
  double kernel_ms = 0;
  int launches = 0;
  std::vector<int32_t> results(4 * n);
  int64_t next = 0;
  while (next < n) {
    // one round: as many pairs (most expensive first) as the arena holds
    size_t used = 0;
    int64_t end = next;
    while (end < n && used + need[order[end]] <= room) {
      const int32_t k = order[end];
      MisTask& t = tasks[k];
      uint8_t* p = base + fixed + used;
      t.trace = p;
      p += align_up(static_cast<size_t>(t.pitch) * t.la, 256);
      const bool multi = mis_strip_cols(t.lb, kThreads, kC) < t.lb;
      t.bnd = multi ? reinterpret_cast<int32_t*>(p) : nullptr;
      if (multi) p += align_up(2 * (static_cast<size_t>(t.la) + 1) * sizeof(int32_t), 256);
      t.line = lines ? p : nullptr;
      if (lines) p += align_up(static_cast<size_t>(t.la) + t.lb, 256);
      t.result = d_results + 4 * static_cast<int64_t>(k);
      used += need[k];
      ++end;
    }
    const int cnt = static_cast<int>(end - next);
    SVS_CU(cudaMemcpy(d_tasks, tasks.data(), n * sizeof(MisTask), cudaMemcpyHostToDevice));
    SVS_CU(cudaMemcpy(d_order, order.data() + next, cnt * sizeof(int32_t), cudaMemcpyHostToDevice));
    SVS_CU(cudaEventRecord(e0));
    misscore_kernel<<<cnt, kThreads>>>(d_tasks, d_order, match, mismatch, -open);
    SVS_CU(cudaGetLastError());
    SVS_CU(cudaEventRecord(e1));
    SVS_CU(cudaDeviceSynchronize());
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    kernel_ms += ms;
    ++launches;
    SVS_CU(cudaMemcpy(results.data(), d_results, 4 * n * sizeof(int32_t), cudaMemcpyDeviceToHost));
    for (int64_t i = next; i < end; ++i) {
      const int32_t k = order[i];
      const int32_t* res = results.data() + 4 * static_cast<int64_t>(k);
      if (res[3] != kMisOk)
        return done(fail(ctx, SVS_ERR_INTERNAL, "misscore: traceback of pair " + std::to_string(k) + " ended with status " + std::to_string(res[3])));
      out[4 * k + 0] = res[0];
      out[4 * k + 1] = res[1];
      out[4 * k + 2] = res[2];
      out[4 * k + 3] = res[1] - res[2];
      if (lines) {
        uint8_t* dst = lines + line_off[k];
        SVS_CU(cudaMemcpy(dst, tasks[k].line, res[1], cudaMemcpyDeviceToHost));
        std::reverse(dst, dst + res[1]);
      }
    }
    next = end;
  }
#undef SVS_CU
  if (stats) {
    const double v[4] = {cells, kernel_ms, static_cast<double>(launches), 0.5 * cells};
    for (int k = 0; k < n_stats && k < 4; ++k) stats[k] = v[k];
  }
  return done(SVS_OK);
}
