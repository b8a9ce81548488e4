// Context, options and read-set entry points of the C ABI (include/svscope_b200.h).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>

#include "context.h"

namespace svs {

namespace {
// Integer-ALU issue-rate probe: 8 independent dependent chains per thread so that the pipe,
// not the 4-cycle latency, limits.  kind 0: add (IADD3), 1: max (VIMNMX), 2: xor (LOP3),
// 3: fused add+max as the DP uses it.
__global__ void nsmid_kernel(unsigned* out) {
  unsigned n;
  asm volatile("mov.u32 %0, %%nsmid;" : "=r"(n));
  out[0] = n;
}

template <int KIND>
__global__ void alu_probe_kernel(int32_t* out, int iters, int32_t seed) {
  int32_t a[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = seed + threadIdx.x * 8 + k;
  const int32_t b = seed ^ 0x5bd1e995, c = seed | 3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (KIND == 0) a[k] = a[k] + c;
      else if (KIND == 1) a[k] = max(a[k], b ^ it);
      else if (KIND == 2) a[k] = a[k] ^ (c + it);
      else a[k] = max(a[k] + c, b - it);
    }
  }
  int32_t r = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) r ^= a[k];
  if (r == 0x7fffffff) out[0] = r;
}
}  // namespace

int ensure_arena(svs_ctx* ctx) {
  if (ctx->arena) return SVS_OK;
  SVS_CUDA(ctx, cudaSetDevice(ctx->device));
  size_t free_b = 0, total_b = 0;
  SVS_CUDA(ctx, cudaMemGetInfo(&free_b, &total_b));
  size_t want = ctx->arena_mb > 0 ? static_cast<size_t>(ctx->arena_mb) << 20
                                  : static_cast<size_t>(static_cast<double>(free_b) * 0.85);
  want = want / 256 * 256;
  if (want > free_b) return fail(ctx, SVS_ERR_CAPACITY, "arena_mb exceeds free device memory");
  SVS_CUDA(ctx, cudaMalloc(&ctx->arena, want));
  ctx->arena_bytes = want;
  if (!ctx->slot_flags) {
    SVS_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->slot_flags), 4096 * sizeof(int)));
    SVS_CUDA(ctx, cudaMemset(ctx->slot_flags, 0, 4096 * sizeof(int)));
  }
  return SVS_OK;
}

}  // namespace svs

using namespace svs;

extern "C" {

const char* svs_version(void) { return "svscope_b200 0.1.0 (sm_100a)"; }

int svs_create(int device, svs_ctx** out) {
  if (!out) return SVS_ERR_ARG;
  *out = nullptr;
  // more hardware queues, so that rounds on different streams really overlap (only effective
  // if the CUDA context of this process does not exist yet)
  setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
  int n = 0;
  cudaError_t err = cudaGetDeviceCount(&n);
  if (err != cudaSuccess || device < 0 || device >= n) return SVS_ERR_CUDA;
  if (cudaSetDevice(device) != cudaSuccess) return SVS_ERR_CUDA;
  svs_ctx* ctx = new svs_ctx();
  ctx->device = device;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
  // %smid indexes the per-SM scratch slots of the persistent alignment kernel: size them by
  // %nsmid (upper bound of %smid), which may exceed the number of enabled SMs
  ctx->n_smid = ctx->sm_count;
  unsigned* d_n = nullptr;
  if (cudaMalloc(&d_n, sizeof(unsigned)) == cudaSuccess) {
    unsigned h_n = 0;
    nsmid_kernel<<<1, 1>>>(d_n);
    if (cudaMemcpy(&h_n, d_n, sizeof(unsigned), cudaMemcpyDeviceToHost) == cudaSuccess && h_n > 0)
      ctx->n_smid = std::max<int>(ctx->sm_count, static_cast<int>(h_n));
    cudaFree(d_n);
  }
  *out = ctx;
  return SVS_OK;
}

void svs_destroy(svs_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->arena) cudaFree(ctx->arena);
  if (ctx->slot_flags) cudaFree(ctx->slot_flags);
  delete ctx;
}

const char* svs_last_error(const svs_ctx* ctx) { return ctx ? ctx->error.c_str() : "no context"; }

int svs_set_option(svs_ctx* ctx, const char* key, int64_t value) {
  if (!ctx || !key) return SVS_ERR_ARG;
  const std::string k(key);
  if (k == "poa_threads") {
    if (value != 128 && value != 256 && value != 384 && value != 512) return fail(ctx, SVS_ERR_ARG, "poa_threads must be 128, 256, 384 or 512");
    ctx->poa_threads = static_cast<int>(value);
  } else if (k == "prune") {
    ctx->prune = value != 0;
  } else if (k == "poa_cols") {
    if (value != 8) return fail(ctx, SVS_ERR_ARG, "poa_cols must be 8");
    ctx->poa_cols = static_cast<int>(value);
  } else if (k == "ring_rows") {
    if (value < 1 || value > 64) return fail(ctx, SVS_ERR_ARG, "ring_rows out of range");
    ctx->ring_rows = static_cast<int>(value);
  } else if (k == "dp_kernel") {
    if (value != 2) return fail(ctx, SVS_ERR_ARG, "dp_kernel must be 2 (the barrier-per-row kernel is gone)");
    ctx->dp_kernel = static_cast<int>(value);
  } else if (k == "workers") {
    if (value < 1 || value > 64) return fail(ctx, SVS_ERR_ARG, "workers out of range");
    ctx->workers = static_cast<int>(value);
  } else if (k == "lane_jobs") {
    if (value < 0 || value > 4096) return fail(ctx, SVS_ERR_ARG, "lane_jobs out of range");
    ctx->lane_jobs = static_cast<int>(value);
  } else if (k == "streams") {
    if (value < 0 || value > 8) return fail(ctx, SVS_ERR_ARG, "streams out of range");
    ctx->streams = static_cast<int>(value);
  } else if (k == "inflight") {
    if (value < 0 || value > 100000) return fail(ctx, SVS_ERR_ARG, "inflight out of range");
    ctx->inflight = static_cast<int>(value);
  } else if (k == "arena_mb") {
    if (value < 0) return fail(ctx, SVS_ERR_ARG, "arena_mb negative");
    if (ctx->arena) {
      cudaSetDevice(ctx->device);
      cudaFree(ctx->arena);
      ctx->arena = nullptr;
      ctx->arena_bytes = 0;
    }
    ctx->arena_mb = value;
  } else {
    return fail(ctx, SVS_ERR_ARG, "unknown option " + k);
  }
  return SVS_OK;
}

int64_t svs_get_option(const svs_ctx* ctx, const char* key) {
  if (!ctx || !key) return -1;
  const std::string k(key);
  if (k == "poa_threads") return ctx->poa_threads;
  if (k == "ring_rows") return ctx->ring_rows;
  if (k == "poa_cols") return ctx->poa_cols;
  if (k == "prune") return ctx->prune;
  if (k == "dp_kernel") return ctx->dp_kernel;
  if (k == "workers") return ctx->workers;
  if (k == "lane_jobs") return ctx->lane_jobs;
  if (k == "inflight") return ctx->inflight;
  if (k == "streams") return ctx->streams;
  if (k == "arena_mb") return ctx->arena_bytes ? static_cast<int64_t>(ctx->arena_bytes >> 20) : ctx->arena_mb;
  if (k == "sm_count") return ctx->sm_count;
  if (k == "n_smid") return ctx->n_smid;
  return -1;
}

int svs_int_alu_probe(svs_ctx* ctx, double* gops, int n) {
  if (!ctx || !gops || n < 4) return fail(ctx, SVS_ERR_ARG, "need room for 4 rates");
  std::lock_guard<std::mutex> lock(ctx->mu);
  SVS_CUDA(ctx, cudaSetDevice(ctx->device));
  int32_t* d = nullptr;
  SVS_CUDA(ctx, cudaMalloc(&d, 64));
  cudaEvent_t e0, e1;
  SVS_CUDA(ctx, cudaEventCreate(&e0));
  SVS_CUDA(ctx, cudaEventCreate(&e1));
  const int blocks = ctx->sm_count * 8, threads = 256, iters = 4096;
  for (int kind = 0; kind < 4; ++kind) {
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(e0);
      if (kind == 0) alu_probe_kernel<0><<<blocks, threads>>>(d, iters, 12345);
      else if (kind == 1) alu_probe_kernel<1><<<blocks, threads>>>(d, iters, 12345);
      else if (kind == 2) alu_probe_kernel<2><<<blocks, threads>>>(d, iters, 12345);
      else alu_probe_kernel<3><<<blocks, threads>>>(d, iters, 12345);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0 && ms < best) best = ms;
    }
    const double ops = static_cast<double>(blocks) * threads * 8.0 * iters;
    gops[kind] = ops / (best * 1e-3) / 1e9;  // thread-level operations per second, in G
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  SVS_CUDA(ctx, cudaGetLastError());
  return SVS_OK;
}

int svs_reads_upload(svs_ctx* ctx, const uint8_t* seqs, const int64_t* off, int64_t n_seqs, svs_reads** out) {
  if (!ctx || !off || !out || n_seqs < 0) return fail(ctx, SVS_ERR_ARG, "null argument");
  SVS_CUDA(ctx, cudaSetDevice(ctx->device));
  svs_reads* r = new svs_reads();
  r->ctx = ctx;
  r->n = n_seqs;
  r->off.assign(off, off + n_seqs + 1);
  const int64_t total = off[n_seqs] - off[0];
  if (off[0] != 0 || total < 0) {
    delete r;
    return fail(ctx, SVS_ERR_ARG, "offsets must start at 0 and be non-decreasing");
  }
  r->host.assign(seqs, seqs + total);
  cudaError_t err = cudaMalloc(reinterpret_cast<void**>(&r->dev), static_cast<size_t>(total) + 64);
  if (err == cudaSuccess && total > 0) {
    // page-lock the host copy for the transfer (it stays in use by the host graph code)
    const bool pinned = cudaHostRegister(r->host.data(), static_cast<size_t>(total), cudaHostRegisterDefault) == cudaSuccess;
    err = cudaMemcpy(r->dev, r->host.data(), static_cast<size_t>(total), cudaMemcpyHostToDevice);
    if (pinned) cudaHostUnregister(r->host.data());
    else cudaGetLastError();
  }
  if (err == cudaSuccess) err = cudaMalloc(reinterpret_cast<void**>(&r->dev_off), sizeof(int64_t) * (static_cast<size_t>(n_seqs) + 1));
  if (err == cudaSuccess)
    err = cudaMemcpy(r->dev_off, r->off.data(), sizeof(int64_t) * (static_cast<size_t>(n_seqs) + 1), cudaMemcpyHostToDevice);
  if (err != cudaSuccess) {
    if (r->dev) cudaFree(r->dev);
    if (r->dev_off) cudaFree(r->dev_off);
    delete r;
    return fail(ctx, SVS_ERR_CUDA, std::string("reads upload: ") + cudaGetErrorString(err));
  }
  *out = r;
  return SVS_OK;
}

void svs_reads_free(svs_reads* reads) {
  if (!reads) return;
  if (reads->ctx) cudaSetDevice(reads->ctx->device);
  if (reads->dev) cudaFree(reads->dev);
  if (reads->dev_off) cudaFree(reads->dev_off);
  delete reads;
}

}  // extern "C"
