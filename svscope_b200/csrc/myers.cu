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

#include <algorithm>
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

// One strip of at most 1024*W pattern rows starting at row `row0`.  Returns this lane's share of
// D[row0+rows][n] - D[row0][n] (sum of the vertical deltas of its rows in the last column).
template <int W>
__device__ int strip_pass(const PairTask& t, int row0, int rows, bool first_strip, bool last_strip, int lane) {
  // pattern as two bit planes of the 2-bit symbol (half the registers of four match masks);
  // rows past the end of the pattern hold arbitrary bits: carries only travel upwards, so they
  // never reach the score row
  uint32_t b0[W], b1[W], pv[W], mv[W];
#pragma unroll
  for (int w = 0; w < W; ++w) {
    pv[w] = 0xffffffffu; mv[w] = 0;
    b0[w] = 0; b1[w] = 0;
    const int base = (lane * W + w) * 32;
    for (int bit = 0; bit < 32; ++bit) {
      const int r = base + bit;
      if (r < rows) {
        const uint32_t s = static_cast<uint32_t>(sym_of(t.a[row0 + r]));
        b0[w] |= (s & 1u) << bit;
        b1[w] |= (s >> 1) << bit;
      }
    }
  }
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
      const uint32_t m0 = 0u - static_cast<uint32_t>(sym & 1), m1 = 0u - static_cast<uint32_t>(sym >> 1);
#pragma unroll
      for (int w = 0; w < W; ++w) {
        uint32_t eq = ~((b0[w] ^ m0) | (b1[w] ^ m1));
        const uint32_t Pv = pv[w], Mv = mv[w];
        const uint32_t xv = eq | Mv;
        if (hout < 0) eq |= 1u;
        const uint32_t xh = (((eq & Pv) + Pv) ^ Pv) | eq;
        uint32_t ph = Mv | ~(xh | Pv);
        uint32_t mh = Pv & xh;
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
  // After the last text character Pv / Mv hold the vertical deltas D[i][n] - D[i-1][n] of the
  // strip's rows, so the strip adds popcount(Pv) - popcount(Mv) over its valid rows to D[.][n].
  int delta = 0;
#pragma unroll
  for (int w = 0; w < W; ++w) {
    const int base = (lane * W + w) * 32;
    const int nvalid = rows - base;
    const uint32_t mask = nvalid >= 32 ? 0xffffffffu : (nvalid <= 0 ? 0u : ((1u << nvalid) - 1u));
    delta += __popc(pv[w] & mask) - __popc(mv[w] & mask);
  }
  return delta;
}

template <int W>
__device__ int pair_distance(const PairTask& t, int lane) {
  const int cap = 32 * 32 * W;
  int total = t.lb;   // D[0][n] = n
  for (int row0 = 0; row0 < t.la; row0 += cap) {
    const int rows = min(cap, t.la - row0);
    const bool last_strip = row0 + cap >= t.la;
    const int part = strip_pass<W>(t, row0, rows, row0 == 0, last_strip, lane);
    total += __reduce_add_sync(0xffffffffu, part);
    __syncwarp();
  }
  return total;
}

// One kernel per words-per-lane class, so that short patterns are not compiled with (and do not
// pay the occupancy of) the register footprint of the longest ones.
template <int W>
__global__ void __launch_bounds__(128, (W <= 16 ? 4 : (W <= 24 ? 3 : 2))) myers_kernel(const PairTask* __restrict__ tasks, const int32_t* __restrict__ ids,
                                                    int32_t* __restrict__ dist, int n_pairs) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= n_pairs) return;
  const int id = ids[warp];
  const PairTask t = tasks[id];
  int d;
  if (t.la == 0) d = t.lb;
  else if (t.lb == 0) d = t.la;
  else d = pair_distance<W>(t, lane);
  if (lane == 0) dist[id] = d;
}

constexpr int kClasses[] = {1, 2, 4, 6, 8, 10, 12, 16, 24, 32};
constexpr int kNumClasses = 10;

int class_of(int32_t la) {
  const int words = (la + 1023) / 1024;
  for (int c = 0; c < kNumClasses; ++c)
    if (words <= kClasses[c]) return c;
  return kNumClasses - 1;  // longer patterns run strip by strip with 32 words per lane
}

cudaError_t launch_class(int c, const PairTask* d_tasks, const int32_t* d_ids, int32_t* d_dist, int n) {
  if (n <= 0) return cudaSuccess;
  const int block = 128;
  const unsigned blocks = static_cast<unsigned>((static_cast<int64_t>(n) * 32 + block - 1) / block);
  switch (kClasses[c]) {
    case 1: myers_kernel<1><<<blocks, block, 0, cudaStreamPerThread>>>(d_tasks, d_ids, d_dist, n); break;
    case 2: myers_kernel<2><<<blocks, block, 0, cudaStreamPerThread>>>(d_tasks, d_ids, d_dist, n); break;
    case 4: myers_kernel<4><<<blocks, block, 0, cudaStreamPerThread>>>(d_tasks, d_ids, d_dist, n); break;
    case 6: myers_kernel<6><<<blocks, block, 0, cudaStreamPerThread>>>(d_tasks, d_ids, d_dist, n); break;
    case 8: myers_kernel<8><<<blocks, block, 0, cudaStreamPerThread>>>(d_tasks, d_ids, d_dist, n); break;
    case 10: myers_kernel<10><<<blocks, block, 0, cudaStreamPerThread>>>(d_tasks, d_ids, d_dist, n); break;
    case 12: myers_kernel<12><<<blocks, block, 0, cudaStreamPerThread>>>(d_tasks, d_ids, d_dist, n); break;
    case 16: myers_kernel<16><<<blocks, block, 0, cudaStreamPerThread>>>(d_tasks, d_ids, d_dist, n); break;
    case 24: myers_kernel<24><<<blocks, block, 0, cudaStreamPerThread>>>(d_tasks, d_ids, d_dist, n); break;
    default: myers_kernel<32><<<blocks, block, 0, cudaStreamPerThread>>>(d_tasks, d_ids, d_dist, n); break;
  }
  return cudaGetLastError();
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
  PairTask* d_tasks = nullptr; int32_t* d_dist = nullptr; int8_t* d_bnd = nullptr; int32_t* d_ids = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  auto cleanup = [&]() {
    ((d_tasks) ? cudaFreeAsync(d_tasks, cudaStreamPerThread) : cudaSuccess); ((d_dist) ? cudaFreeAsync(d_dist, cudaStreamPerThread) : cudaSuccess); ((d_bnd) ? cudaFreeAsync(d_bnd, cudaStreamPerThread) : cudaSuccess); ((d_ids) ? cudaFreeAsync(d_ids, cudaStreamPerThread) : cudaSuccess);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
  };
#define SVS_CU(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { cleanup(); \
    return fail(ctx, SVS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); } } while (0)
  if (bnd_total) {
    SVS_CU(cudaMallocAsync(reinterpret_cast<void**>(&d_bnd), bnd_total, cudaStreamPerThread));
    size_t off = 0;
    for (int64_t k = 0; k < n; ++k) {
      if (tasks[k].la > strip_cap) {
        tasks[k].bnd = d_bnd + off;
        off += align_up(static_cast<size_t>(tasks[k].lb), 16);
      }
    }
  }
  SVS_CU(cudaMallocAsync(reinterpret_cast<void**>(&d_tasks), n * sizeof(PairTask), cudaStreamPerThread));
  SVS_CU(cudaMallocAsync(reinterpret_cast<void**>(&d_dist), n * sizeof(int32_t), cudaStreamPerThread));
  SVS_CU(svs_memcpy_pt(d_tasks, tasks.data(), n * sizeof(PairTask), cudaMemcpyHostToDevice));
  SVS_CU(cudaEventCreate(&e0));
  SVS_CU(cudaEventCreate(&e1));
  // pairs grouped by words-per-lane class; inside a class longest text first
  std::vector<std::vector<int32_t>> by_class(kNumClasses);
  for (int64_t k = 0; k < n; ++k) by_class[class_of(tasks[k].la)].push_back(static_cast<int32_t>(k));
  std::vector<int32_t> ids;
  std::vector<size_t> coff(kNumClasses + 1, 0);
  for (int c = 0; c < kNumClasses; ++c) {
    auto& v = by_class[c];
    std::sort(v.begin(), v.end(), [&](int32_t x, int32_t y) { return tasks[x].lb > tasks[y].lb; });
    coff[c] = ids.size();
    ids.insert(ids.end(), v.begin(), v.end());
  }
  coff[kNumClasses] = ids.size();
  SVS_CU(cudaMallocAsync(reinterpret_cast<void**>(&d_ids), n * sizeof(int32_t), cudaStreamPerThread));
  SVS_CU(svs_memcpy_pt(d_ids, ids.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice));
  SVS_CU(cudaEventRecord(e0, cudaStreamPerThread));
  int launches = 0;
  for (int c = kNumClasses - 1; c >= 0; --c) {
    const int nc = static_cast<int>(by_class[c].size());
    if (!nc) continue;
    SVS_CU(launch_class(c, d_tasks, d_ids + coff[c], d_dist, nc));
    ++launches;
  }
  SVS_CU(cudaEventRecord(e1, cudaStreamPerThread));
  SVS_CU(cudaStreamSynchronize(cudaStreamPerThread));
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  SVS_CU(svs_memcpy_pt(out, d_dist, n * sizeof(int32_t), cudaMemcpyDeviceToHost));
#undef SVS_CU
  cleanup();
  if (stats) {
    const double v[5] = {cells, static_cast<double>(ms), bytes, static_cast<double>(n), static_cast<double>(launches)};
    for (int k = 0; k < n_stats && k < 5; ++k) stats[k] = v[k];
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
  std::lock_guard<std::mutex> lock(ctx->mu_ed);
  if (!only_acgt(reads)) return fail(ctx, SVS_ERR_UNSUPPORTED, "edit distance kernel expects upper-case A,C,G,T");
  std::vector<int64_t> va(a, a + n_pairs), vb(b, b + n_pairs);
  return run_pairs(ctx, reads, va, vb, dist, stats, n_stats);
}

int svs_edit_distance_matrix(svs_ctx* ctx, const svs_reads* reads, const int64_t* members, const int64_t* group_off,
                             int64_t n_groups, int32_t* dist, const int64_t* dist_off, double* stats, int n_stats) {
  if (!ctx || !reads || n_groups < 0) return fail(ctx, SVS_ERR_ARG, "null argument");
  std::lock_guard<std::mutex> lock(ctx->mu_ed);
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
