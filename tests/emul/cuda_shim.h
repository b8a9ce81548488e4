// TEST INFRASTRUCTURE: just enough of the CUDA execution model to run the product's warp-pipelined
// DP source (svscope_b200/csrc/poa_dp2.cuh) on the CPU.
//   * one OS thread per warp, so warps run concurrently and ahead of one another as on the device
//     (a warp polling a neighbour's progress word really waits for another thread);
//   * the 32 lanes of a warp are user-level contexts of that thread, switched round-robin at every warp
//     collective (__shfl_sync, __shfl_up_sync, __ballot_sync, __reduce_add_sync, __syncwarp): a lane that
//     reaches a collective deposits its operand and hands the thread to the next lane; when it gets the
//     thread back all 32 operands are there (operands are double-buffered by collective parity).  Lane-
//     divergent code between collectives (lane 0 polling while the others wait at the next collective)
//     behaves as on the device.  Collectives must be reached by all 32 lanes, which the product code
//     guarantees (full masks, warp-uniform control flow around them);
//   * __syncthreads = collective, barrier over the warp threads by lane 0, collective;
//   * shared memory is an ordinary buffer, volatile polls are ordinary volatile loads (x86 ordering).
// Nothing here is shipped.  Include the system headers first: the macros at the end would break them.
#pragma once
#include <sched.h>
#include <x86intrin.h>

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <string>
#include <thread>
#include <vector>

struct alignas(8) int2 { int x, y; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) int4 { int x, y, z, w; };
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }

extern "C" void shim_switch(void** save_sp, void* load_sp);   // cuda_shim.cpp (x86-64 System V)

namespace shim {
struct Dim { unsigned x = 0, y = 0, z = 0; };
constexpr size_t kStackBytes = 256 << 10;

struct CtaRt {
  std::atomic<int> count{0};
  std::atomic<int> gen{0};
  int n_warps = 0;
};
struct WarpRt {
  void* sp[32];
  void* main_sp = nullptr;
  char* stacks = nullptr;
  bool done[32];
  unsigned ncoll[32];      // collectives passed, per lane
  int32_t slot[2][32];
  int cur = 0, n_done = 0;
  unsigned base_tid = 0;
  CtaRt* cta = nullptr;
  std::function<void(int)> body;
};
extern thread_local WarpRt* rt;
extern thread_local int lane;
void run_warp(WarpRt* w);   // cuda_shim.cpp
void to_lane(int next);     // hands the thread to lane `next` (or to the warp's main context if all are done)

// the calling lane has deposited its operand: let the other 31 lanes reach the same collective
static inline void round() {
  WarpRt* w = rt;
  const int me = w->cur;
  to_lane((me + 1) & 31);
  (void)me;
}
static inline int32_t exchange(int32_t v, int src_lane) {
  WarpRt* w = rt;
  const int me = w->cur;
  const unsigned par = w->ncoll[me]++ & 1u;
  w->slot[par][me] = v;
  round();
  return w->slot[par][src_lane & 31];
}
static inline void cta_barrier() {   // lane 0 only
  CtaRt& c = *rt->cta;
  const int g = c.gen.load(std::memory_order_acquire);
  if (c.count.fetch_add(1, std::memory_order_acq_rel) + 1 == c.n_warps) { c.count.store(0, std::memory_order_relaxed); c.gen.store(g + 1, std::memory_order_release); }
  else { int k = 0; while (c.gen.load(std::memory_order_acquire) == g) { if (++k > 256) { sched_yield(); k = 0; } } }
}
}  // namespace shim

extern thread_local shim::Dim threadIdx, blockIdx, blockDim;

static inline void __syncwarp(unsigned = 0xffffffffu) { shim::exchange(0, 0); }
static inline void __syncthreads() {
  shim::exchange(0, 0);
  if (shim::rt->cur == 0) shim::cta_barrier();
  shim::exchange(0, 0);
}
static inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline long long clock64() { return static_cast<long long>(__rdtsc()); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }

template <class T> static inline T shim_exchange(T v, int src_lane) {
  static_assert(sizeof(T) == 4, "32-bit values only");
  int32_t raw;
  std::memcpy(&raw, &v, 4);
  const int32_t got = shim::exchange(raw, src_lane);
  T out;
  std::memcpy(&out, &got, 4);
  return out;
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src_lane) { return shim_exchange(v, src_lane); }
static inline int __shfl_sync(unsigned, bool v, int src_lane) { return shim_exchange(static_cast<int>(v), src_lane); }
template <class T> static inline T __shfl_up_sync(unsigned, T v, int d) {
  const int me = shim::rt->cur;
  return shim_exchange(v, me >= d ? me - d : me);
}
static inline unsigned __ballot_sync(unsigned, bool pred) {
  shim::WarpRt* w = shim::rt;
  const int me = w->cur;
  const unsigned par = w->ncoll[me]++ & 1u;
  w->slot[par][me] = pred ? 1 : 0;
  shim::round();
  unsigned m = 0;
  for (int l = 0; l < 32; ++l) m |= static_cast<unsigned>(w->slot[par][l] & 1) << l;
  return m;
}
static inline unsigned __reduce_add_sync(unsigned, unsigned v) {
  shim::WarpRt* w = shim::rt;
  const int me = w->cur;
  const unsigned par = w->ncoll[me]++ & 1u;
  w->slot[par][me] = static_cast<int32_t>(v);
  shim::round();
  unsigned s = 0;
  for (int l = 0; l < 32; ++l) s += static_cast<unsigned>(w->slot[par][l]);
  return s;
}
template <class T> static inline T __ldcg(const T* p) {
  T v;
  std::memcpy(&v, const_cast<const T*>(p), sizeof(T));
  std::atomic_signal_fence(std::memory_order_seq_cst);
  return v;
}
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) {
  return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
}

// Runs body(tid) for n_threads CUDA threads as one CTA: one OS thread per warp, 32 contexts each.
template <class F> static inline void shim_run_cta(int n_threads, F body) {
  const int nw = n_threads / 32;
  shim::CtaRt cta;
  cta.n_warps = nw;
  std::vector<std::unique_ptr<shim::WarpRt>> warps;
  for (int w = 0; w < nw; ++w) {
    warps.emplace_back(new shim::WarpRt());
    warps.back()->base_tid = static_cast<unsigned>(32 * w);
    warps.back()->cta = &cta;
    warps.back()->body = body;
  }
  std::vector<std::thread> th;
  for (int w = 0; w < nw; ++w) {
    shim::WarpRt* p = warps[w].get();
    th.emplace_back([p, n_threads]() {
      blockIdx.x = 0;
      blockDim.x = static_cast<unsigned>(n_threads);
      shim::run_warp(p);
    });
  }
  for (auto& t : th) t.join();
}

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __align__(n) alignas(n)
