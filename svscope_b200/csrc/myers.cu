// Batched unit-cost edit distance, Myers/Hyyro bit-parallel, one warp per read pair
// (svs_edit_distance_matrix / svs_edit_distance_pairs).
//
// Contract: Levenshtein.distance(a, b) as used by the read-by-read matrix of the commented
// FindSomClust (src/DecisionMaker.py:76-84; dependency README.md:23).
//
// The pattern (one of the two reads) is laid out over the 32 lanes of a warp, W consecutive
// 32-bit words per lane (32*32*W pattern rows per strip).  The text is consumed as an
// anti-diagonal wavefront over the lanes: at step s lane l advances its W words by text
// character s-l, taking the horizontal delta (+1/0/-1) of the word block above it from lane
// l-1 by one warp shuffle that also carries the text symbol.  Patterns longer than one strip
// are processed strip by strip with the boundary deltas parked in global memory.
// Integer/logic work only, bound by the ALU pipes; reads stream once from HBM.
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "context.h"

namespace svs {
namespace {

struct PairTask {
  const uint8_t* a;  // pattern
  const uint8_t* b;  // text
  int32_t la, lb;
  int8_t* bnd;       // lb bytes of strip boundary deltas (only if la > one strip)
};

constexpr int kMaxW = 32;

__device__ __forceinline__ int sym_of(uint8_t ch) { return (ch >> 1) & 3; }  // A,C,T,G -> 0,1,2,3

// One strip of at most 1024*W pattern rows starting at row `row0`.  Returns (in the lane that
// owns the last pattern row of the strip) the sum of horizontal deltas at that row.
template <int W>
__device__ int strip_pass(const PairTask& t, int row0, int rows, bool first_strip, bool last_strip, int lane) {
  uint32_t peq[W][4], pv[W], mv[W];
#pragma unroll
  for (int w = 0; w < W; ++w) {
    pv[w] = 0xffffffffu; mv[w] = 0;
#pragma unroll
    for (int s = 0; s < 4; ++s) peq[w][s] = 0;
    const int base = (lane * W + w) * 32;
    for (int bit = 0; bit < 32; ++bit) {
      const int r = base + bit;
      if (r < rows) {
        const int s = sym_of(t.a[row0 + r]);
#pragma unroll
        for (int q = 0; q < 4; ++q) peq[w][q] |= (s == q) ? (1u << bit) : 0u;
      }
    }
  }
  const int last = rows - 1;
  const int last_lane = last / (32 * W), last_word = (last / 32) % W;
  const uint32_t last_bit = 1u << (last % 32);
  int score = 0;
  int carry = 0;  // packed: (hout + 1) | sym << 2 | valid << 4 handed to the next lane
  const int n = t.lb;
  for (int step = 0; step < n + 31; ++step) {
    const int from_up = __shfl_up_sync(0xffffffffu, carry, 1);
    int hin, sym, valid;
    if (lane == 0) {
      valid = step < n;
      sym = valid ? sym_of(t.b[step]) : 0;
      hin = first_strip ? 1 : (valid ? static_cast<int>(t.bnd[step]) : 0);
    } else {
      hin = (from_up & 3) - 1;
      sym = (from_up >> 2) & 3;
      valid = (from_up >> 4) & 1;
    }
    int hout = hin;
    if (valid) {
#pragma unroll
      for (int w = 0; w < W; ++w) {
        uint32_t eq = peq[w][0];
        eq = (sym == 1) ? peq[w][1] : eq;
        eq = (sym == 2) ? peq[w][2] : eq;
        eq = (sym == 3) ? peq[w][3] : eq;
        const uint32_t Pv = pv[w], Mv = mv[w];
        const uint32_t xv = eq | Mv;
        if (hout < 0) eq |= 1u;
        const uint32_t xh = (((eq & Pv) + Pv) ^ Pv) | eq;
        uint32_t ph = Mv | ~(xh | Pv);
        uint32_t mh = Pv & xh;
        if (lane == last_lane && w == last_word) score += ((ph & last_bit) ? 1 : 0) - ((mh & last_bit) ? 1 : 0);
        const int ho = static_cast<int>(ph >> 31) - static_cast<int>(mh >> 31);
        ph <<= 1; mh <<= 1;
        if (hout < 0) mh |= 1u; else if (hout > 0) ph |= 1u;
        pv[w] = mh | ~(xv | ph);
        mv[w] = ph & xv;
        hout = ho;
      }
      if (!last_strip && lane == 31) t.bnd[step - 31] = static_cast<int8_t>(hout);
    }
    carry = (hout + 1) | (sym << 2) | (valid << 4);
  }
  return score;
}

template <int W>
__device__ int pair_distance(const PairTask& t, int lane) {
  const int cap = 32 * 32 * W;
  int total = t.la;
  for (int row0 = 0; row0 < t.la; row0 += cap) {
    const int rows = min(cap, t.la - row0);
    const bool last_strip = row0 + cap >= t.la;
    const int sc = strip_pass<W>(t, row0, rows, row0 == 0, last_strip, lane);
    // only the last strip's bottom row is the distance row
    const int last = rows - 1;
    const int owner = last / (32 * W);
    const int v = __shfl_sync(0xffffffffu, sc, owner);
    if (last_strip) total += v;
    __syncwarp();
  }
  return total;
}

__global__ void __launch_bounds__(128) myers_kernel(const PairTask* __restrict__ tasks, int32_t* __restrict__ dist,
                                                    int n_pairs) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= n_pairs) return;
  const PairTask t = tasks[warp];
  int d;
  if (t.la == 0) d = t.lb;
  else if (t.lb == 0) d = t.la;
  else {
    const int words = (t.la + 1023) / 1024;  // words per lane needed for a single strip
    if (words <= 1) d = pair_distance<1>(t, lane);
    else if (words <= 2) d = pair_distance<2>(t, lane);
    else if (words <= 4) d = pair_distance<4>(t, lane);
    else if (words <= 8) d = pair_distance<8>(t, lane);
    else if (words <= 12) d = pair_distance<12>(t, lane);
    else if (words <= 16) d = pair_distance<16>(t, lane);
    else if (words <= 24) d = pair_distance<24>(t, lane);
    else d = pair_distance<kMaxW>(t, lane);
  }
  if (lane == 0) dist[warp] = d;
}

int run_pairs(svs_ctx* ctx, const svs_reads* reads, const std::vector<int64_t>& a, const std::vector<int64_t>& b,
              int32_t* out, double* stats, int n_stats) {
  const int64_t n = static_cast<int64_t>(a.size());
  if (n == 0) return SVS_OK;
  SVS_CUDA(ctx, cudaSetDevice(ctx->device));
  std::vector<PairTask> tasks(n);
  size_t bnd_total = 0;
  double cells = 0, bytes = 0;
  const int strip_cap = 32 * 32 * kMaxW;
  for (int64_t k = 0; k < n; ++k) {
    int64_t ia = a[k], ib = b[k];
    if (ia < 0 || ia >= reads->n || ib < 0 || ib >= reads->n) return fail(ctx, SVS_ERR_ARG, "read index out of range");
    int64_t la = reads->off[ia + 1] - reads->off[ia], lb = reads->off[ib + 1] - reads->off[ib];
    if (la > lb) { std::swap(ia, ib); std::swap(la, lb); }  // shorter read is the pattern
    tasks[k].a = reads->dev + reads->off[ia];
    tasks[k].b = reads->dev + reads->off[ib];
    tasks[k].la = static_cast<int32_t>(la);
    tasks[k].lb = static_cast<int32_t>(lb);
    tasks[k].bnd = nullptr;
    if (la > strip_cap) bnd_total += align_up(static_cast<size_t>(lb), 16);
    cells += static_cast<double>(la) * static_cast<double>(lb);
    bytes += static_cast<double>(la + lb) + 4;
  }
  PairTask* d_tasks = nullptr; int32_t* d_dist = nullptr; int8_t* d_bnd = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  auto cleanup = [&]() {
    cudaFree(d_tasks); cudaFree(d_dist); cudaFree(d_bnd);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
  };
#define SVS_CU(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { cleanup(); \
    return fail(ctx, SVS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); } } while (0)
  if (bnd_total) {
    SVS_CU(cudaMalloc(&d_bnd, bnd_total));
    size_t off = 0;
    for (int64_t k = 0; k < n; ++k) {
      if (tasks[k].la > strip_cap) {
        tasks[k].bnd = d_bnd + off;
        off += align_up(static_cast<size_t>(tasks[k].lb), 16);
      }
    }
  }
  SVS_CU(cudaMalloc(&d_tasks, n * sizeof(PairTask)));
  SVS_CU(cudaMalloc(&d_dist, n * sizeof(int32_t)));
  SVS_CU(cudaMemcpy(d_tasks, tasks.data(), n * sizeof(PairTask), cudaMemcpyHostToDevice));
  SVS_CU(cudaEventCreate(&e0));
  SVS_CU(cudaEventCreate(&e1));
  const int block = 128;
  const int64_t blocks = (n * 32 + block - 1) / block;
  SVS_CU(cudaEventRecord(e0));
  myers_kernel<<<static_cast<unsigned>(blocks), block>>>(d_tasks, d_dist, static_cast<int>(n));
  SVS_CU(cudaGetLastError());
  SVS_CU(cudaEventRecord(e1));
  SVS_CU(cudaDeviceSynchronize());
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  SVS_CU(cudaMemcpy(out, d_dist, n * sizeof(int32_t), cudaMemcpyDeviceToHost));
#undef SVS_CU
  cleanup();
  if (stats) {
    const double v[4] = {cells, static_cast<double>(ms), bytes, static_cast<double>(n)};
    for (int k = 0; k < n_stats && k < 4; ++k) stats[k] = v[k];
  }
  return SVS_OK;
}

bool only_acgt(const svs_reads* reads) {
  for (uint8_t ch : reads->host)
    if (ch != 'A' && ch != 'C' && ch != 'G' && ch != 'T') return false;
  return true;
}

}  // namespace
}  // namespace svs

using namespace svs;

extern "C" {

int svs_edit_distance_pairs(svs_ctx* ctx, const svs_reads* reads, const int64_t* a, const int64_t* b,
                            int64_t n_pairs, int32_t* dist, double* stats, int n_stats) {
  if (!ctx || !reads || n_pairs < 0) return fail(ctx, SVS_ERR_ARG, "null argument");
  std::lock_guard<std::mutex> lock(ctx->mu);
  if (!only_acgt(reads)) return fail(ctx, SVS_ERR_UNSUPPORTED, "edit distance kernel expects upper-case A,C,G,T");
  std::vector<int64_t> va(a, a + n_pairs), vb(b, b + n_pairs);
  return run_pairs(ctx, reads, va, vb, dist, stats, n_stats);
}

int svs_edit_distance_matrix(svs_ctx* ctx, const svs_reads* reads, const int64_t* members, const int64_t* group_off,
                             int64_t n_groups, int32_t* dist, const int64_t* dist_off, double* stats, int n_stats) {
  if (!ctx || !reads || n_groups < 0) return fail(ctx, SVS_ERR_ARG, "null argument");
  std::lock_guard<std::mutex> lock(ctx->mu);
  if (!only_acgt(reads)) return fail(ctx, SVS_ERR_UNSUPPORTED, "edit distance kernel expects upper-case A,C,G,T");
  std::vector<int64_t> va, vb;
  std::vector<int64_t> where;  // output index of dist[i][j]
  for (int64_t g = 0; g < n_groups; ++g) {
    const int64_t n = group_off[g + 1] - group_off[g];
    for (int64_t i = 0; i < n; ++i) {
      dist[dist_off[g] + i * n + i] = 0;
      for (int64_t j = i + 1; j < n; ++j) {
        va.push_back(members[group_off[g] + i]);
        vb.push_back(members[group_off[g] + j]);
        where.push_back(dist_off[g] + i * n + j);
        where.push_back(dist_off[g] + j * n + i);
      }
    }
  }
  std::vector<int32_t> flat(va.size());
  const int rc = run_pairs(ctx, reads, va, vb, flat.data(), stats, n_stats);
  if (rc) return rc;
  for (size_t k = 0; k < flat.size(); ++k) {
    dist[where[2 * k]] = flat[k];
    dist[where[2 * k + 1]] = flat[k];
  }
  return SVS_OK;
}

}  // extern "C"
