// Partial-order alignment on the device (replaces `spoa.poa(sequences, 1)`; reference call
// sites src/DataScanner.py:206,213 and src/DecisionMaker.py:160,171).
//
// poa_window_kernel   one CTA per window (group of sequences): for every read, rank-ordered export
//                     of the device-resident graph (poa_dgraph.h), bands, the warp-pipelined dynamic
//                     programme and its traceback (poa_dp2.cuh), merge of the path, rank order; at
//                     the end MSA rows and consensus.  T threads x 8 read columns per pass.
//
// Integer arithmetic only; no tensor cores (nothing here is a dense contraction).
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

#include "poa_cell.h"
#include "poa_kernels.h"
#include "poa_task.h"
#include "poa_window.h"

namespace svs {

namespace {

constexpr int32_t kSrcGlobal = 1 << 30;   // predecessor source: exported row in global memory (low bits: its slot)

// ---------------------------------------------------------------------------------------------
// Window kernel: the CTA owns a window (group of sequences) from its first to its last read.
// Per read: rank-ordered export of the graph (poa_dgraph.h) -> bands -> dynamic programme ->
// traceback -> merge of the path into the graph -> rank order; at the end MSA rows and the
// heaviest-bundle consensus are written to the output arena.  The graph never leaves the
// device; the host only sees the result record of the window.
struct CtaExec {
  uint32_t* warp_tot;   // shared memory, 32 words
  int32_t* sweep_ring;  // shared memory, 2 x 64 x 2 words (depth sweeps)
  template <class F> __device__ __forceinline__ void run(F f) { f(threadIdx.x, blockDim.x); __syncthreads(); }
  template <class F> __device__ __forceinline__ void one(F f) { if (threadIdx.x == 0) f(); __syncthreads(); }
  template <class F, class G> __device__ __forceinline__ void two(F f, G g) {
    if (threadIdx.x == 0) f(); else if (threadIdx.x == 32) g();
    __syncthreads();
  }
  __device__ __forceinline__ void atomic_max(uint32_t* p, uint32_t v) { atomicMax(p, v); }
  __device__ __forceinline__ void atomic_min(int32_t* p, int32_t v) { atomicMin(p, v); }
  __device__ __forceinline__ void atomic_add(uint32_t* p, uint32_t v) { atomicAdd(p, v); }
  __device__ void scan(uint32_t* a, uint32_t n) {
    const uint32_t nt = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    uint32_t carry = 0;
    for (uint32_t base = 0; base < n; base += nt) {
      const uint32_t i = base + tid;
      uint32_t v = i < n ? a[i] : 0;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= static_cast<uint32_t>(d)) v += o;
      }
      if (lane == 31) warp_tot[warp] = v;
      __syncthreads();
      uint32_t add = carry, total = 0;
      for (uint32_t w = 0; w < nw; ++w) {
        const uint32_t t = warp_tot[w];
        if (w < warp) add += t;
        total += t;
      }
      if (i < n) a[i] = v + add;
      carry += total;
      __syncthreads();
    }
  }
  // a[i] = min(a[i..n))
  __device__ void suffix_min(int32_t* a, uint32_t n) {
    const uint32_t nt = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    int32_t* tot = reinterpret_cast<int32_t*>(warp_tot);
    int32_t carry = INT32_MAX;
    if (n == 0) return;
    for (int64_t base = static_cast<int64_t>((n - 1) / nt) * nt; base >= 0; base -= nt) {
      const uint32_t i = static_cast<uint32_t>(base) + tid;
      int32_t v = i < n ? a[i] : INT32_MAX;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int32_t o = __shfl_down_sync(0xffffffffu, v, d);
        if (lane + d < 32) v = min(v, o);
      }
      if (lane == 0) tot[warp] = v;
      __syncthreads();
      int32_t add = carry, total = INT32_MAX;
      for (uint32_t w = 0; w < nw; ++w) {
        const int32_t t = tot[w];
        if (w > warp) add = min(add, t);
        total = min(total, t);
      }
      if (i < n) a[i] = min(v, add);
      carry = min(carry, total);
      __syncthreads();
    }
  }
  // Path-length intervals (dg_depth_forward / dg_depth_backward): warp 0 sweeps forward, warp 1
  // backward (a CTA of one warp does both in turn).  A row is one step of the warp: the lanes
  // fetch the row's predecessors in parallel; values of the last 64 rows live in a ring in
  // shared memory, older ones in global memory.
  __device__ void depth_sweeps(const WinMem& m, uint32_t R) {
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    int32_t* dp = m.depth;
    // partial values of the backward sweep start as "no successor seen"
    for (uint32_t i = tid; i <= R; i += blockDim.x) { dp[4 * i + 2] = INT32_MAX; dp[4 * i + 3] = -1; }
    __syncthreads();
    if (warp == 0) {
      int32_t* ring = sweep_ring;                       // [64][2]
      if (lane == 0) { ring[0] = 0; ring[1] = 0; }      // row 0
      __syncwarp();
      for (uint32_t i0 = 1; i0 <= R; i0 += 32) {
        const uint32_t nrows = min(32u, R - i0 + 1);
        const uint32_t pb = lane < nrows ? m.pred_off[i0 + lane] : 0;
        const uint32_t pe = lane < nrows ? m.pred_off[i0 + lane + 1] : 0;
        for (uint32_t r = 0; r < nrows; ++r) {
          const uint32_t i = i0 + r;
          const uint32_t b = __shfl_sync(0xffffffffu, pb, r), e = __shfl_sync(0xffffffffu, pe, r);
          int32_t lo = INT32_MAX, hi = 0;
          for (uint32_t k = b + lane; k < e; k += 32) {
            const uint32_t p = m.preds[k];
            int32_t a, c;
            if (i - p < 64) { a = ring[2 * (p & 63)]; c = ring[2 * (p & 63) + 1]; }
            else { a = dp[4 * p]; c = dp[4 * p + 1]; }
            lo = min(lo, a);
            hi = max(hi, c);
          }
          lo = __reduce_min_sync(0xffffffffu, lo) + 1;
          hi = __reduce_max_sync(0xffffffffu, hi) + 1;
          if (lane == 0) {
            ring[2 * (i & 63)] = lo; ring[2 * (i & 63) + 1] = hi;
            dp[4 * i] = lo; dp[4 * i + 1] = hi;
          }
          __syncwarp();
        }
      }
    }
    if (warp == (nw > 1 ? 1u : 0u)) {
      // ring slot of row r: what successors less than 64 rows away found; successors further away wrote
      // to global memory, and they are all done before the 32-row batch of r starts, so that part is
      // fetched per batch (one row per lane) and joined when the row is finalised
      int32_t* ring = sweep_ring + 128;                 // [64][2]
      for (uint32_t k = lane; k < 64; k += 32) { ring[2 * k] = INT32_MAX; ring[2 * k + 1] = -1; }
      __syncwarp();
      for (uint32_t i1 = R; i1 >= 1; i1 -= min(i1, 32u)) {
        // rows i1, i1-1, ... (up to 32), lane l holds the predecessor range and the far part of row i1 - l
        const uint32_t nrows = min(32u, i1);
        const uint32_t pb = lane < nrows ? m.pred_off[i1 - lane] : 0;
        const uint32_t pe = lane < nrows ? m.pred_off[i1 - lane + 1] : 0;
        const int32_t far_lo = lane < nrows ? dp[4 * (i1 - lane) + 2] : INT32_MAX;
        const int32_t far_hi = lane < nrows ? dp[4 * (i1 - lane) + 3] : -1;
        for (uint32_t r = 0; r < nrows; ++r) {
          const uint32_t i = i1 - r;
          const uint32_t b = __shfl_sync(0xffffffffu, pb, r), e = __shfl_sync(0xffffffffu, pe, r);
          int32_t lo = min(ring[2 * (i & 63)], __shfl_sync(0xffffffffu, far_lo, r));
          int32_t hi = max(ring[2 * (i & 63) + 1], __shfl_sync(0xffffffffu, far_hi, r));
          if (lo == INT32_MAX) { lo = 0; hi = 0; }   // no successor
          const int32_t a = lo + 1, c = hi + 1;
          __syncwarp();
          if (lane == 0) {
            dp[4 * i + 2] = lo; dp[4 * i + 3] = hi;
            ring[2 * (i & 63)] = INT32_MAX; ring[2 * (i & 63) + 1] = -1;   // the slot is free for row i - 64
          }
          __syncwarp();
          for (uint32_t k = b + lane; k < e; k += 32) {   // the predecessors of a row are distinct rows
            const uint32_t p = m.preds[k];
            if (p == 0) continue;
            if (i - p < 64) {
              ring[2 * (p & 63)] = min(ring[2 * (p & 63)], a);
              ring[2 * (p & 63) + 1] = max(ring[2 * (p & 63) + 1], c);
            } else {
              dp[4 * p + 2] = min(dp[4 * p + 2], a);
              dp[4 * p + 3] = max(dp[4 * p + 3], c);
            }
          }
          __syncwarp();
        }
        if (i1 <= 32) break;
      }
      if (lane == 0) { dp[2] = 0; dp[3] = 0; }
    }
    __syncthreads();
  }
};

#include "poa_dp2.cuh"

template <int T, int kC>
__global__ void __launch_bounds__(T, 512 / T)
poa_window_kernel(const WinParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ WinState S;
  __shared__ WinMem m;
  __shared__ PoaTask tk;
  __shared__ WinCaps caps;
  __shared__ uint32_t warp_tot[32];
  __shared__ unsigned long long cyc[8];
  __shared__ int s_next, s_slot;
  __shared__ uint64_t s_cells, s_rows, s_exported, s_need, s_pairs, s_bases, s_steps, s_preds;
  __shared__ uint32_t s_nalign, s_retries;
  __shared__ unsigned long long s_need2, s_eval[5];
  __shared__ int32_t sweep_ring[256];
  CtaExec x{warp_tot, sweep_ring};
  const int tid = threadIdx.x;
  if (P.sm_limit > 0) {   // profiling aid: full per-SM occupancy on a few SMs only (small arena, short ncu replays)
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    if (smid >= static_cast<unsigned>(P.sm_limit)) return;
  }
  // scratch slot: any free one (at most n_slots CTAs of this kernel are resident at a time)
  if (tid == 0) {
    int k = static_cast<int>((blockIdx.x * 7919u) % static_cast<unsigned>(P.n_slots));
    while (atomicCAS(&P.slot_flags[k], 0, 1) != 0) k = (k + 1 == P.n_slots) ? 0 : k + 1;
    s_slot = k;
  }
  __syncthreads();
  uint8_t* const slot = P.slot_base + static_cast<uint64_t>(s_slot) * P.slot_bytes;
  const Scores s = P.s;
  while (true) {
    __syncthreads();
    if (tid == 0) s_next = atomicAdd(P.counter, 1);
    __syncthreads();
    if (s_next >= P.n_windows) break;
    const int widx = P.order[s_next];
    const WinDesc d = P.desc[widx];
    if (tid == 0) {
      caps = d.caps;
      const uint64_t fixed = win_layout(slot, P.slot_bytes, d.caps, &m);
      S.nv = S.ne = S.nseq = 0;
      S.path_off = 0;
      S.err = fixed + 4096 > P.slot_bytes ? kWinNodeCap : kWinOk;
      S.max_indeg = 1; S.n_export = S.n_single = S.n_new = S.msa_cols = 0;
      S.last_score = 0; S.last_len = 0;
      S.err_pending = 0; S.topo_serial = 0; S.topo_rounds = 0; S.chg[0] = S.chg[1] = 0;
      for (int k = 0; k < 8; ++k) cyc[k] = 0;
      s_cells = s_rows = s_exported = s_need = s_pairs = s_bases = s_steps = s_preds = 0;
      s_nalign = s_retries = 0;
      for (int k = 0; k < 5; ++k) s_eval[k] = 0;
    }
    __syncthreads();
    for (uint32_t q = 0; q < d.caps.nseq && S.err == kWinOk; ++q) {
      const int64_t id = P.members[d.member_begin + q];
      const int64_t o0 = P.read_off[id];
      const uint32_t L = static_cast<uint32_t>(P.read_off[id + 1] - o0);
      const uint8_t* seq = P.reads + o0;
      if (P.pair_cnt != nullptr && tid == 0) P.pair_cnt[d.member_begin + q] = 0;
      if (L == 0) continue;                       // empty sequences are ignored (no MSA row)
      if (S.nseq == 0) {
        dg_init_chain(x, m, caps, &S, seq, L);
        continue;
      }
      long long t0 = clock64();
      dg_export(x, m, caps, &S, s, static_cast<uint32_t>(P.ring_rows), false);
      if (tid == 0) {
        const long long t1 = clock64();
        cyc[0] += static_cast<unsigned long long>(t1 - t0);
        const uint32_t R = S.nv;
        tk.letter = m.r_letter; tk.pred_off = m.pred_off; tk.preds = m.preds; tk.flags = m.r_flags;
        tk.xslot = m.xslot; tk.h0 = m.h0; tk.col0code = m.col0code; tk.node_id = m.node_id;
        tk.single_before = m.single_before; tk.depth = m.depth; tk.read = seq;
        tk.R = R; tk.L = L;
        const uint32_t cpp = T * kC;
        tk.npass = (L + cpp - 1) / cpp;
        tk.strip = ((L + tk.npass - 1) / tk.npass + kC - 1) / kC * kC;
        tk.w1 = static_cast<uint32_t>((static_cast<uint64_t>(L) + kC - 1 + 15) / 16 * 16);
        tk.w2 = static_cast<uint32_t>((static_cast<uint64_t>(L) + kC - 1 + 7) / 8 * 8 * 2);
        tk.ldx = (static_cast<uint64_t>(L) + 3 + kC + 7) / 8 * 8;
        const uint64_t xbytes = dg_align(static_cast<uint64_t>(S.n_export) * tk.ldx * 4, 256);
        tk.xrows = reinterpret_cast<int32_t*>(m.dyn);
        tk.codes = m.dyn + xbytes;
        tk.codes_cap = m.dyn_bytes > xbytes ? m.dyn_bytes - xbytes : 0;
        tk.coff = m.coff;
        tk.bnd = m.bnd; tk.result = m.result; tk.path = m.path; tk.path_cap = m.path_cap;
        const int64_t span = 10ll * (static_cast<int64_t>(R) + L + 2);
        tk.prune = (P.prune && L >= 1024 && R >= 1024 && span < (1 << 21)) ? 1u : 0u;
        const double spb = S.last_len ? static_cast<double>(S.last_score) / S.last_len : 4.0;
        tk.lb_guess = static_cast<int32_t>((spb - P.prune_margin) * static_cast<double>(L)) - 40;
        const uint64_t n1 = S.n_single;
        const uint64_t full = n1 * tk.w1 + (static_cast<uint64_t>(R) - n1) * tk.w2 + 64;
        if (S.max_indeg > kMaxIndeg) S.err = kWinIndeg;
        else if (span >= kMaxKeySpan) S.err = kWinScoreSpan;
        else if (xbytes + 4096 > m.dyn_bytes || (!tk.prune && false && full > tk.codes_cap)) { S.err = kWinCodesCap; s_need = full + xbytes; }
        s_cells += (static_cast<uint64_t>(R) + 1) * (static_cast<uint64_t>(L) + 1);
        s_rows += R;
        s_exported += S.n_export;
        s_bases += L;
        s_preds += m.pred_off[R + 1];
        s_nalign += 1;
        m.result[3] = 0;
      }
      __syncthreads();
      if (S.err != kWinOk) break;
      t0 = clock64();
      bool v2_overflow = false;
      {
        // bands (pruned: from the guessed lower bound; else full rows) and exact-size code rows; a
        // result below the guess repeats the alignment with the score found (a true lower bound)
        int32_t lb = tk.lb_guess;
        bool have_lb = tk.prune != 0;
        for (int attempt = 0; attempt < 3; ++attempt) {
          if (tid == 0) { tk.result[0] = 0; tk.result[1] = INT32_MIN; }
          compute_bands2<T>(x, tk, s, lb, have_lb, m.band, m.coff, static_cast<TbRow*>(m.tbrow), &s_need2);
          if (s_need2 > tk.codes_cap) { v2_overflow = true; break; }
          dp2_align<T>(tk, s, P.ring_rows, smem_raw, m.band, m.coff, s_eval);
          const int32_t found_row = tk.result[0], found = tk.result[1];
          if (!have_lb || (found_row > 0 && found >= lb)) break;
          if (tid == 0) s_retries += 1;
          have_lb = found_row > 0 && found > kNegBand / 2;
          lb = found;
          __syncthreads();
        }
        if (v2_overflow) {
          if (tid == 0) { S.err = kWinCodesCap; s_need = s_need2; }
          __syncthreads();
          break;
        }
      }
      long long t1 = clock64();
      if (tid < 32) tb3_walk_warp(tk, s, static_cast<const TbRow*>(m.tbrow));
      __syncthreads();
      long long t2 = clock64();
      if (tid == 0) {
        cyc[1] += static_cast<unsigned long long>(t1 - t0);
        cyc[2] += static_cast<unsigned long long>(t2 - t1);
        if (m.result[2] < 0 || m.result[0] <= 0) S.err = kWinTraceback;
        S.last_score = m.result[1];
        if (m.result[2] > 0) s_steps += static_cast<uint64_t>(m.result[2]);
        S.last_len = L;
      }
      __syncthreads();
      if (S.err != kWinOk) break;
      const int32_t np = m.result[2];
      if (P.pairs_out != nullptr && d.pairs_off >= 0) {   // debug: forward pairs of this sequence
        int32_t* dst = P.pairs_out + 2 * (d.pairs_off + static_cast<int64_t>(s_pairs));
        for (int32_t a = tid; a < np; a += T) {
          dst[2 * a] = m.path[2 * (np - 1 - a)];
          dst[2 * a + 1] = m.path[2 * (np - 1 - a) + 1];
        }
        __syncthreads();
        if (tid == 0) { P.pair_cnt[d.member_begin + q] = np; s_pairs += static_cast<uint64_t>(np); }
      }
      const uint32_t n_old = S.nv;
      dg_add_alignment(x, m, caps, &S, m.path, np, seq, L);
      long long t3 = clock64();
      if (S.err != kWinOk) break;
      dg_toposort(x, m, caps, &S, L, n_old);
      if (tid == 0) {
        const long long t4 = clock64();
        cyc[3] += static_cast<unsigned long long>(t3 - t2);
        cyc[4] += static_cast<unsigned long long>(t4 - t3);
      }
    }
    // ---- consensus, MSA, result record --------------------------------------------------------
    __syncthreads();
    const long long t5 = clock64();
    __shared__ uint32_t s_cons_len;
    __shared__ unsigned long long s_out_off;
    if (tid == 0) { s_cons_len = 0; s_out_off = 0; }
    __syncthreads();
    if (S.err == kWinOk && S.nseq > 0) {
      dg_export(x, m, caps, &S, s, static_cast<uint32_t>(P.ring_rows), true);
      uint8_t* cons_tmp = reinterpret_cast<uint8_t*>(m.band);
      const uint64_t V1 = static_cast<uint64_t>(caps.vcap) + 2;
      x.one([&]() { s_cons_len = dg_consensus_serial(m, &S, m.depth, m.depth + V1, cons_tmp); });
      uint32_t* head = m.single_before;
      uint32_t* col_of = reinterpret_cast<uint32_t*>(m.xslot);
      if (P.want_msa) dg_msa_columns(x, m, &S, head, col_of);
      x.one([&]() {
        const uint64_t bytes = (P.want_msa ? static_cast<uint64_t>(S.nseq) * S.msa_cols : 0) + s_cons_len;
        const unsigned long long need = dg_align(bytes, 16);
        const unsigned long long at = atomicAdd(P.out_cursor, need);
        s_out_off = at;
        if (at + need > P.out_cap) { S.err = kWinOutCap; s_need = need; }
      });
      if (S.err == kWinOk) {
        uint8_t* out = P.out_base + s_out_off;
        const uint64_t msa_bytes = P.want_msa ? static_cast<uint64_t>(S.nseq) * S.msa_cols : 0;
        if (P.want_msa) dg_msa_rows(x, m, &S, col_of, m.seq_len, S.nseq, out);
        for (uint32_t k = tid; k < s_cons_len; k += T) out[msa_bytes + k] = cons_tmp[k];
      }
    }
    __syncthreads();
    if (tid == 0) {
      WinResult r;
      r.status = S.err;
      r.msa_rows = (S.err == kWinOk && P.want_msa) ? S.nseq : 0;
      r.msa_cols = (S.err == kWinOk && P.want_msa) ? S.msa_cols : 0;
      r.cons_len = S.err == kWinOk ? s_cons_len : 0;
      r.out_off = s_out_off;
      r.n_align = s_nalign; r.retries = s_retries;
      r.nodes = S.nv; r.edges = S.ne;
      r.cells = s_cells; r.rows = s_rows; r.exported = s_exported; r.need_bytes = s_need;
      r.eval_cells = 8ull * s_eval[0];
      for (int k = 0; k < 4; ++k) r.warp_cyc[k] = s_eval[1 + k];
      r.read_bases = s_bases; r.path_steps = s_steps; r.pred_entries = s_preds;
      cyc[5] += static_cast<unsigned long long>(clock64() - t5);
      for (int k = 0; k < 8; ++k) r.cyc[k] = cyc[k];
      P.results[widx] = r;
    }
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    atomicExch(&P.slot_flags[s_slot], 0);
  }
}

}  // namespace

int poa_cols_per_thread(int, int) { return 8; }

size_t poa_dp_smem_bytes(int threads, int ring_rows, int) { return dp2_smem_bytes(threads, ring_rows); }

int poa_dp_cols_per_pass(int threads, int) { return threads * 8; }

// ---- window kernel launchers ----------------------------------------------------------------
size_t poa_window_smem_bytes(int threads, int ring_rows, int) { return dp2_smem_bytes(threads, ring_rows); }

int poa_window_ctas_per_sm(int threads, int ring_rows, int cols) {
  if (cols != 8 || (threads != 128 && threads != 256 && threads != 384 && threads != 512)) return 0;
  const size_t smem = poa_window_smem_bytes(threads, ring_rows, cols) + 2048;   // + static shared memory
  if (smem > 227 * 1024) return 0;
  const int by_smem = static_cast<int>((227 * 1024) / (smem + 1024));
  const int cap = 512 / threads;   // 128 registers per thread
  return by_smem < cap ? by_smem : cap;
}

template <int T>
static cudaError_t window_cfg(int bytes) {
  return cudaFuncSetAttribute(poa_window_kernel<T, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

cudaError_t poa_window_configure(int threads, int ring_rows, int cols) {
  if (poa_window_ctas_per_sm(threads, ring_rows, cols) <= 0) return cudaErrorInvalidValue;
  const int bytes = static_cast<int>(poa_window_smem_bytes(threads, ring_rows, cols));
  if (threads == 128) return window_cfg<128>(bytes);
  if (threads == 256) return window_cfg<256>(bytes);
  if (threads == 384) return window_cfg<384>(bytes);
  return window_cfg<512>(bytes);
}

cudaError_t poa_window_launch(const WinParams& p, int grid, int threads, int cols, cudaStream_t stream) {
  if (grid <= 0) return cudaSuccess;
  const size_t smem = poa_window_smem_bytes(threads, p.ring_rows, cols);
  if (threads == 128) poa_window_kernel<128, 8><<<grid, 128, smem, stream>>>(p);
  else if (threads == 256) poa_window_kernel<256, 8><<<grid, 256, smem, stream>>>(p);
  else if (threads == 384) poa_window_kernel<384, 8><<<grid, 384, smem, stream>>>(p);
  else if (threads == 512) poa_window_kernel<512, 8><<<grid, 512, smem, stream>>>(p);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

}  // namespace svs
