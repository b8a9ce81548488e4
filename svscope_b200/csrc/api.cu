// Context, options and read-set entry points of the C ABI (include/svscope_b200.h).
#include <cstring>
#include <string>

#include "context.h"

namespace svs {

int ensure_arena(svs_ctx* ctx) {
  if (ctx->arena) return SVS_OK;
  SVS_CUDA(ctx, cudaSetDevice(ctx->device));
  size_t free_b = 0, total_b = 0;
  SVS_CUDA(ctx, cudaMemGetInfo(&free_b, &total_b));
  size_t want = ctx->arena_mb > 0 ? static_cast<size_t>(ctx->arena_mb) << 20
                                  : static_cast<size_t>(static_cast<double>(free_b) * 0.70);
  want = want / 256 * 256;
  if (want > free_b) return fail(ctx, SVS_ERR_CAPACITY, "arena_mb exceeds free device memory");
  SVS_CUDA(ctx, cudaMalloc(&ctx->arena, want));
  ctx->arena_bytes = want;
  return SVS_OK;
}

}  // namespace svs

using namespace svs;

extern "C" {

const char* svs_version(void) { return "svscope_b200 0.1.0 (sm_100a)"; }

int svs_create(int device, svs_ctx** out) {
  if (!out) return SVS_ERR_ARG;
  *out = nullptr;
  int n = 0;
  cudaError_t err = cudaGetDeviceCount(&n);
  if (err != cudaSuccess || device < 0 || device >= n) return SVS_ERR_CUDA;
  if (cudaSetDevice(device) != cudaSuccess) return SVS_ERR_CUDA;
  svs_ctx* ctx = new svs_ctx();
  ctx->device = device;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
  *out = ctx;
  return SVS_OK;
}

void svs_destroy(svs_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->arena) cudaFree(ctx->arena);
  delete ctx;
}

const char* svs_last_error(const svs_ctx* ctx) { return ctx ? ctx->error.c_str() : "no context"; }

int svs_set_option(svs_ctx* ctx, const char* key, int64_t value) {
  if (!ctx || !key) return SVS_ERR_ARG;
  const std::string k(key);
  if (k == "poa_threads") {
    if (value != 128 && value != 256 && value != 512) return fail(ctx, SVS_ERR_ARG, "poa_threads must be 128, 256 or 512");
    ctx->poa_threads = static_cast<int>(value);
  } else if (k == "ring_rows") {
    if (value < 1 || value > 64) return fail(ctx, SVS_ERR_ARG, "ring_rows out of range");
    ctx->ring_rows = static_cast<int>(value);
  } else if (k == "workers") {
    if (value < 1 || value > 64) return fail(ctx, SVS_ERR_ARG, "workers out of range");
    ctx->workers = static_cast<int>(value);
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
  if (k == "workers") return ctx->workers;
  if (k == "arena_mb") return ctx->arena_bytes ? static_cast<int64_t>(ctx->arena_bytes >> 20) : ctx->arena_mb;
  if (k == "sm_count") return ctx->sm_count;
  return -1;
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
  if (err == cudaSuccess && total > 0)
    err = cudaMemcpy(r->dev, r->host.data(), static_cast<size_t>(total), cudaMemcpyHostToDevice);
  if (err != cudaSuccess) {
    if (r->dev) cudaFree(r->dev);
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
  delete reads;
}

}  // extern "C"
