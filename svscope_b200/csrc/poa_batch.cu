// Batched partial-order alignment: host side of svs_poa_batch (include/svscope_b200.h).
//
// A group of sequences is one graph.  Its alignments are strictly sequential (each read is
// aligned to the graph that already contains the previous ones), so the parallelism is
// across groups.  Groups wait in one cost-sorted queue; every worker thread keeps two
// "lanes" in flight (stream + pinned staging + arena slice) and for each lane repeats
//   export rank-ordered graphs -> one H2D copy -> DP kernel (one CTA per alignment, largest
//   first) -> traceback kernel -> one D2H copy -> merge the paths into the graphs (host)
// in ping-pong, so the host work of one lane hides behind the kernels of the other, and a
// lane pulls the next group from the queue as soon as one of its groups is finished.
#include <algorithm>
#include <atomic>
#include <map>
#include <mutex>
#include <chrono>
#include <cstring>
#include <memory>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "context.h"
#include "poa_cell.h"
#include "poa_graph.h"
#include "poa_kernels.h"
#include "poa_task.h"

struct svs_poa_result {
  std::vector<std::string> consensus;
  std::vector<std::vector<std::string>> msa;
  double stats[24] = {0};
};

namespace svs {
namespace {

struct PoaJob {
  int64_t group = 0;
  bool done = false;
  bool force_block = false;      // the next alignment overflowed its slot: give it a full-size block
  double score_per_base = 4.0;   // of the last alignment (drives the pruning guess)
  std::vector<int64_t> seq_ids;
  size_t next = 0;
  PoaGraph graph;
  RankedGraph rg;
  std::vector<std::vector<int32_t>>* record = nullptr;  // forward pairs of every alignment
  double cost = 0;
};

struct TaskPlan {
  PoaJob* job;
  int64_t seq_id;
  uint32_t R, L, strip, npass, path_cap, w1, w2;
  uint64_t ldx;
  size_t in_bytes, scratch_bytes, out_bytes, codes_bytes, xrows_bytes, bnd_bytes, band_bytes;
  bool prune, in_slot;
  double cells;
};

struct WorkerStats {
  double cells = 0, alignments = 0, dp_ms = 0, tb_ms = 0, dp_launches = 0, tb_launches = 0,
         h2d = 0, d2h = 0, algo_bytes = 0, exported_rows = 0, rows = 0,
         host_wait_ms = 0, host_merge_ms = 0, host_plan_ms = 0, host_pack_ms = 0,
         refill_ms = 0, starved = 0, launch_ms = 0, final_ms = 0, inflight_ms = 0, h2d_ms = 0, d2h_ms = 0,
         prune_retries = 0, slot_overflows = 0;
};

inline double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// First-fit allocator over the device arena; a lane takes one block per round.
class ArenaAlloc {
 public:
  void reset(size_t bytes) {
    std::lock_guard<std::mutex> g(mu_);
    free_.clear();
    free_[0] = bytes;
  }
  bool alloc(size_t bytes, size_t* off) {
    std::lock_guard<std::mutex> g(mu_);
    for (auto it = free_.begin(); it != free_.end(); ++it) {
      if (it->second >= bytes) {
        *off = it->first;
        const size_t rest = it->second - bytes;
        const size_t at = it->first + bytes;
        free_.erase(it);
        if (rest) free_[at] = rest;
        return true;
      }
    }
    return false;
  }
  void release(size_t off, size_t bytes) {
    std::lock_guard<std::mutex> g(mu_);
    auto it = free_.emplace(off, bytes).first;
    auto nx = std::next(it);
    if (nx != free_.end() && it->first + it->second == nx->first) {
      it->second += nx->second;
      free_.erase(nx);
    }
    if (it != free_.begin()) {
      auto pv = std::prev(it);
      if (pv->first + pv->second == it->first) {
        pv->second += it->second;
        free_.erase(it);
      }
    }
  }

 private:
  std::mutex mu_;
  std::map<size_t, size_t> free_;
};

// ---------------------------------------------------------------------------------------------
// Minimal fork-join helper: runs fn(i) for i in [0, n) on up to `threads` std::threads.
template <class F>
void parallel_for(int n, int threads, F&& fn) {
  if (n <= 0) return;
  threads = std::max(1, std::min(threads, n));
  if (threads == 1) {
    for (int i = 0; i < n; ++i) fn(i);
    return;
  }
  std::atomic<int> next{0};
  auto body = [&]() {
    for (int i; (i = next.fetch_add(1)) < n;) fn(i);
  };
  std::vector<std::thread> th;
  th.reserve(threads - 1);
  for (int t = 1; t < threads; ++t) th.emplace_back(body);
  body();
  for (auto& t : th) t.join();
}

// One stream of rounds.  A round aligns the next read of every unfinished group of the
// stream: one persistent-kernel launch (largest alignment first) plus, for alignments whose
// scratch does not fit a per-SM slot, a classic launch with explicit arena blocks.
struct Stream {
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};  // begin, dp end, tb end, done
  uint8_t* h_in = nullptr;
  size_t h_in_cap = 0;
  uint8_t* h_out = nullptr;
  size_t h_out_cap = 0;
  int* d_counter = nullptr;
  size_t blk_off = 0, blk_bytes = 0;
  std::vector<PoaJob*> jobs;          // unfinished groups of this stream
  std::vector<TaskPlan> inflight;
  std::vector<size_t> out_offs;
  size_t in_bytes = 0, out_bytes = 0;
  int n_slot = 0, n_big = 0;
  bool busy = false;
};

class Scheduler {
 public:
  Scheduler(svs_ctx* ctx, const svs_reads* reads, const Scores& s, svs_poa_result* res, bool want_msa)
      : ctx_(ctx), reads_(reads), s_(s), res_(res), want_msa_(want_msa) {}
  ~Scheduler() {
    for (Stream& st : streams_) {
      if (st.h_in) cudaFreeHost(st.h_in);
      if (st.h_out) cudaFreeHost(st.h_out);
      if (st.d_counter) cudaFree(st.d_counter);
      for (auto& e : st.ev) if (e) cudaEventDestroy(e);
      if (st.stream) cudaStreamDestroy(st.stream);
    }
  }
  WorkerStats stats;
  int err = 0;
  std::string errmsg;

  int run(std::vector<PoaJob>& jobs) {
    const int njobs = static_cast<int>(jobs.size());
    threads_ = std::max(1, ctx_->workers);
    if (const char* pm = getenv("SVS_PRUNE_MARGIN")) prune_margin_ = atof(pm);
    // Device arena: [per-SM scratch slots | blocks for graph arrays, paths and oversized alignments]
    n_sm_ = std::max(1, ctx_->sm_count);
    // persistent mode needs exactly one resident CTA per SM (512 threads, > 114 KB shared memory);
    // otherwise every alignment takes the classic path with explicit arena blocks
    persistent_ = poa_persistent_supported(ctx_->poa_threads, ctx_->ring_rows, ctx_->poa_cols);
    cols_ = static_cast<uint32_t>(poa_cols_per_thread(ctx_->poa_threads, ctx_->poa_cols));
    ctas_per_sm_ = std::max(1, poa_persistent_ctas_per_sm(ctx_->poa_threads, ctx_->ring_rows, ctx_->poa_cols));
    n_slots_ = std::max(n_sm_, ctx_->n_smid) * ctas_per_sm_;
    slot_bytes_ = persistent_ ? (static_cast<size_t>(static_cast<double>(ctx_->arena_bytes) * 0.88) / n_slots_) / 4096 * 4096 : 0;
    slot_base_ = static_cast<uint8_t*>(ctx_->arena);
    block_base_ = slot_base_ + slot_bytes_ * n_slots_;
    block_bytes_ = ctx_->arena_bytes - slot_bytes_ * n_slots_;
    blocks_.reset(block_bytes_);
    const int ns = std::max(1, std::min(ctx_->streams > 0 ? ctx_->streams : 2, njobs));
    streams_.resize(ns);
    for (Stream& st : streams_) {
      if (!check(cudaStreamCreateWithFlags(&st.stream, cudaStreamNonBlocking), "stream")) return err;
      for (auto& e : st.ev) if (!check(cudaEventCreate(&e), "event")) return err;
      if (!check(cudaMalloc(reinterpret_cast<void**>(&st.d_counter), sizeof(int)), "counter")) return err;
    }
    // cost-sorted, dealt round-robin: every stream sees the same size distribution
    std::vector<PoaJob*> order;
    for (auto& j : jobs) order.push_back(&j);
    std::stable_sort(order.begin(), order.end(), [](const PoaJob* a, const PoaJob* b) { return a->cost > b->cost; });
    for (int k = 0; k < njobs; ++k) streams_[k % ns].jobs.push_back(order[k]);
    while (true) {
      bool any = false;
      for (Stream& st : streams_) {
        if (st.busy) {
          const double t0 = now_ms();
          if (!check(cudaEventSynchronize(st.ev[3]), "poa round")) return drain();
          stats.host_wait_ms += now_ms() - t0;
          if (!collect(st)) return drain();
        }
        if (!st.jobs.empty()) {
          if (!prepare_and_launch(st)) return drain();
          if (st.busy || !st.jobs.empty()) any = true;
        }
      }
      if (!any) break;
    }
    return err;
  }

 private:
  void set_err(int code, const std::string& msg) {
    std::lock_guard<std::mutex> g(err_mu_);
    if (!err) { err = code; errmsg = msg; }
  }
  int drain() {
    for (Stream& st : streams_) {
      if (st.stream) cudaStreamSynchronize(st.stream);
      if (st.busy && st.blk_bytes) blocks_.release(st.blk_off, st.blk_bytes);
      st.busy = false;
    }
    return err ? err : SVS_ERR_INTERNAL;
  }
  bool check(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return true;
    set_err(SVS_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
    return false;
  }
  const uint8_t* seq_ptr(int64_t id) const { return reads_->host.data() + reads_->off[id]; }
  uint32_t seq_len(int64_t id) const { return static_cast<uint32_t>(reads_->off[id + 1] - reads_->off[id]); }

  // Consumes sequences that need no alignment (empty ones, and the first one of a graph);
  // returns true when job->seq_ids[job->next] has to be aligned on the device.
  bool advance_to_alignment(PoaJob* job) {
    while (job->next < job->seq_ids.size()) {
      const int64_t id = job->seq_ids[job->next];
      const uint32_t len = seq_len(id);
      if (len == 0) {
        if (job->record) job->record->emplace_back();
        ++job->next;
      } else if (job->graph.empty()) {
        job->graph.add_alignment(nullptr, nullptr, 0, seq_ptr(id), len);
        if (job->record) job->record->emplace_back();
        ++job->next;
      } else {
        return true;
      }
    }
    return false;
  }

  void finalize(PoaJob* job) {
    if (res_) {
      res_->consensus[job->group] = job->graph.consensus();
      if (want_msa_) res_->msa[job->group] = job->graph.msa();
    }
    job->graph = PoaGraph();   // release host memory early
    job->rg = RankedGraph();
    job->done = true;
  }

  bool plan(PoaJob* job, TaskPlan* tp) {
    job->graph.export_ranked(PoaScoring{s_.m, s_.n, s_.g, s_.e, s_.q, s_.c},
                             static_cast<uint32_t>(ctx_->ring_rows), &job->rg);
    const RankedGraph& g = job->rg;
    if (g.max_indeg > kMaxIndeg) {
      set_err(SVS_ERR_UNSUPPORTED, "graph node with more than 31 in-edges");
      return false;
    }
    tp->job = job;
    tp->seq_id = job->seq_ids[job->next];
    tp->R = g.R;
    tp->L = seq_len(tp->seq_id);
    const int64_t worst = 10;  // |penalties| <= 10 checked at entry
    if (worst * (static_cast<int64_t>(tp->R) + tp->L + 2) >= kMaxKeySpan) {
      set_err(SVS_ERR_UNSUPPORTED, "alignment too large for 25-bit scores (|V| + L must stay below 1.6 M)");
      return false;
    }
    const uint32_t cpp = static_cast<uint32_t>(poa_dp_cols_per_pass(ctx_->poa_threads, ctx_->poa_cols));
    const uint32_t C = cols_;   // columns per thread: every thread stores C codes / words at once
    tp->npass = (tp->L + cpp - 1) / cpp;
    tp->strip = ((tp->L + tp->npass - 1) / tp->npass + C - 1) / C * C;
    tp->w1 = static_cast<uint32_t>((static_cast<uint64_t>(tp->L) + C - 1 + 15) / 16 * 16);       // 1 B per cell
    tp->w2 = static_cast<uint32_t>((static_cast<uint64_t>(tp->L) + C - 1 + 7) / 8 * 8 * 2);       // 2 B per cell
    tp->ldx = (static_cast<uint64_t>(tp->L) + 3 + C + 7) / 8 * 8;
    tp->path_cap = tp->R + tp->L + 2;
    const size_t R1 = static_cast<size_t>(g.R) + 1;
    tp->in_bytes = align_up(R1 * 16, 16) /*depth*/ +
                   align_up(R1, 16) * 2 /*letter, flags*/ + align_up((R1 + 1) * 4, 16) * 2 /*pred_off, single_before*/ +
                   align_up(g.preds.size() * 4, 16) + align_up(R1 * 4, 16) * 3 /*xslot,h0,node_id*/ +
                   align_up(R1 * 2, 16);
    const size_t n1 = g.single_before[g.R + 1];
    tp->codes_bytes = align_up(n1 * tp->w1 + (static_cast<size_t>(tp->R) - n1) * tp->w2 + 64, 256);
    tp->xrows_bytes = align_up(static_cast<size_t>(g.n_export) * tp->ldx * 4, 256);
    tp->bnd_bytes = align_up(R1 * 4 * 8, 256);
    tp->band_bytes = align_up(R1 * 8, 256);
    tp->scratch_bytes = tp->codes_bytes + tp->xrows_bytes + tp->bnd_bytes + tp->band_bytes;
    // exact pruning pays off on long alignments; scores must stay above -2^21 (kNegBand = -2^22)
    tp->prune = ctx_->prune && tp->L >= 1024 && tp->R >= 1024 &&
                worst * (static_cast<int64_t>(tp->R) + tp->L + 2) < (1 << 21);
    // Slot path: pruned alignments store band-limited code rows (the kernel checks the real
    // size and reports an overflow, which sends the alignment to the block path next round).
    const size_t small = tp->xrows_bytes + tp->bnd_bytes + tp->band_bytes;
    const size_t codes_est = tp->prune ? tp->codes_bytes / 2 : tp->codes_bytes;
    tp->in_slot = persistent_ && !job->force_block && small + codes_est <= slot_bytes_;
    tp->out_bytes = align_up(16 + static_cast<size_t>(tp->path_cap) * 8, 16);
    tp->cells = (static_cast<double>(tp->R) + 1) * (static_cast<double>(tp->L) + 1);
    return true;
  }

  bool ensure_pinned(uint8_t** buf, size_t* cap, size_t need) {
    if (*cap >= need) return true;
    if (*buf) cudaFreeHost(*buf);
    *buf = nullptr;
    *cap = 0;
    const size_t want = need + need / 2 + (4 << 20);
    if (!check(cudaMallocHost(reinterpret_cast<void**>(buf), want), "cudaMallocHost")) return false;
    *cap = want;
    return true;
  }

  // Exports the graphs of all unfinished groups of the stream (in parallel), packs them into
  // the pinned staging buffer (in parallel) and enqueues H2D, kernels and D2H.
  bool prepare_and_launch(Stream& st) {
    const double t_plan0 = now_ms();
    // drop finished groups
    st.jobs.erase(std::remove_if(st.jobs.begin(), st.jobs.end(), [](PoaJob* j) { return j->done; }), st.jobs.end());
    const int nj = static_cast<int>(st.jobs.size());
    std::vector<TaskPlan> plans(nj);
    std::vector<uint8_t> live(nj, 0);
    parallel_for(nj, threads_, [&](int k) {
      PoaJob* job = st.jobs[k];
      try {
        if (!advance_to_alignment(job)) { finalize(job); return; }
        if (plan(job, &plans[k])) live[k] = 1;
      } catch (const std::exception& ex) {
        set_err(SVS_ERR_INTERNAL, std::string("graph: ") + ex.what());
      }
    });
    if (err) return false;
    st.inflight.clear();
    for (int k = 0; k < nj; ++k) if (live[k]) st.inflight.push_back(plans[k]);
    st.jobs.erase(std::remove_if(st.jobs.begin(), st.jobs.end(), [](PoaJob* j) { return j->done; }), st.jobs.end());
    std::vector<TaskPlan>& chunk = st.inflight;
    if (chunk.empty()) { st.busy = false; return true; }
    // slot tasks first (largest first), then oversized ones
    std::sort(chunk.begin(), chunk.end(), [&](const TaskPlan& a, const TaskPlan& b) {
      const bool ba = !a.in_slot, bb = !b.in_slot;
      if (ba != bb) return bb;
      return a.cells > b.cells;
    });
    // Oversized alignments take their scratch from the block region; keep the round within
    // ~45 % of it (two streams) and defer the rest to the next round of this stream.
    const size_t budget = static_cast<size_t>(static_cast<double>(block_bytes_) * 0.45);
    {
      size_t acc = sizeof(PoaTask) * chunk.size() + 8192;
      size_t keep = 0;
      for (; keep < chunk.size(); ++keep) {
        const TaskPlan& t = chunk[keep];
        const size_t need = t.in_bytes + t.out_bytes + (!t.in_slot ? t.scratch_bytes : 0);
        if (keep > 0 && acc + need > budget) break;
        acc += need;
      }
      chunk.resize(keep);
    }
    const int n = static_cast<int>(chunk.size());
    size_t in_total = align_up(sizeof(PoaTask) * n, 256), out_total = 0, big_total = 0;
    st.n_slot = 0;
    for (auto& t : chunk) {
      in_total += t.in_bytes;
      out_total += t.out_bytes;
      if (!t.in_slot) big_total += t.scratch_bytes; else ++st.n_slot;
    }
    st.n_big = n - st.n_slot;
    in_total = align_up(in_total, 256);
    out_total = align_up(out_total, 256);
    const size_t block = align_up(in_total + out_total + big_total + 4096, 4096);
    if (block > block_bytes_) {
      set_err(SVS_ERR_CAPACITY, "one alignment needs " + std::to_string(block >> 20) + " MiB of arena blocks, " +
                                    std::to_string(block_bytes_ >> 20) + " MiB available (raise arena_mb)");
      return false;
    }
    int spins = 0;
    while (!blocks_.alloc(block, &st.blk_off)) {
      // another stream holds the space: finish its round first
      bool freed = false;
      for (Stream& o : streams_) {
        if (&o != &st && o.busy) {
          if (!check(cudaEventSynchronize(o.ev[3]), "poa round")) return false;
          if (!collect(o)) return false;
          freed = true;
          break;
        }
      }
      if (!freed || ++spins > 64) {
        set_err(SVS_ERR_CAPACITY, "arena blocks exhausted (raise arena_mb)");
        return false;
      }
    }
    st.blk_bytes = block;
    if (!ensure_pinned(&st.h_in, &st.h_in_cap, in_total) || !ensure_pinned(&st.h_out, &st.h_out_cap, out_total)) {
      blocks_.release(st.blk_off, st.blk_bytes);
      return false;
    }
    const double t_pack0 = now_ms();
    stats.host_plan_ms += t_pack0 - t_plan0;
    uint8_t* d_in = block_base_ + st.blk_off;
    uint8_t* d_out = d_in + in_total;
    uint8_t* d_big = d_out + out_total;
    PoaTask* h_tasks = reinterpret_cast<PoaTask*>(st.h_in);
    // offsets first (serial, cheap), then the copies in parallel
    std::vector<size_t> in_offs(n), big_offs(n, 0);
    st.out_offs.assign(n, 0);
    size_t in_off = align_up(sizeof(PoaTask) * n, 256), out_off = 0, big_off = 0;
    for (int k = 0; k < n; ++k) {
      in_offs[k] = in_off;
      in_off += chunk[k].in_bytes;
      st.out_offs[k] = out_off;
      out_off += chunk[k].out_bytes;
      if (!chunk[k].in_slot) { big_offs[k] = big_off; big_off += chunk[k].scratch_bytes; }
    }
    parallel_for(n, threads_, [&](int k) {
      const TaskPlan& tp = chunk[k];
      const RankedGraph& g = tp.job->rg;
      PoaTask& t = h_tasks[k];
      const size_t R1 = static_cast<size_t>(g.R) + 1;
      size_t off = in_offs[k];
      auto put = [&](const void* src, size_t bytes) -> const uint8_t* {
        std::memcpy(st.h_in + off, src, bytes);
        const uint8_t* dptr = d_in + off;
        off += align_up(bytes, 16);
        return dptr;
      };
      t.depth = reinterpret_cast<const int32_t*>(put(g.depth.data(), R1 * 16));
      t.letter = put(g.letter.data(), R1);
      t.flags = put(g.flags.data(), R1);
      t.pred_off = reinterpret_cast<const uint32_t*>(put(g.pred_off.data(), (R1 + 1) * 4));
      t.preds = reinterpret_cast<const uint32_t*>(put(g.preds.data(), g.preds.size() * 4));
      t.xslot = reinterpret_cast<const int32_t*>(put(g.xslot.data(), R1 * 4));
      t.h0 = reinterpret_cast<const int32_t*>(put(g.h0.data(), R1 * 4));
      t.node_id = reinterpret_cast<const uint32_t*>(put(g.node_id.data(), R1 * 4));
      t.single_before = reinterpret_cast<const uint32_t*>(put(g.single_before.data(), (R1 + 1) * 4));
      t.col0code = reinterpret_cast<const uint16_t*>(put(g.col0code.data(), R1 * 2));
      t.read = reads_->dev + reads_->off[tp.seq_id];
      t.R = tp.R; t.L = tp.L; t.strip = tp.strip; t.npass = tp.npass;
      t.w1 = tp.w1;
      t.w2 = tp.w2;
      t.ldx = tp.ldx;
      // slot layout: [exported rows | strip boundaries | bands | traceback codes (rest of the slot)]
      t.off_xrows = 0;
      t.off_bnd = tp.xrows_bytes;
      t.off_band = tp.xrows_bytes + tp.bnd_bytes;
      t.off_codes = tp.xrows_bytes + tp.bnd_bytes + tp.band_bytes;
      t.codes_cap = tp.in_slot ? slot_bytes_ - t.off_codes : tp.codes_bytes;
      t.prune = (tp.prune && tp.in_slot) ? 1u : 0u;   // persistent path only
      // guess: score per read base of the previous alignment of this graph, minus a margin
      t.lb_guess = static_cast<int32_t>((tp.job->score_per_base - prune_margin_) * static_cast<double>(tp.L)) - 40;
      if (!tp.in_slot) {
        uint8_t* base = d_big + big_offs[k];
        t.codes = base + t.off_codes;
        t.xrows = reinterpret_cast<int32_t*>(base + t.off_xrows);
        t.bnd = reinterpret_cast<int32_t*>(base + t.off_bnd);
      } else {
        t.codes = nullptr; t.xrows = nullptr; t.bnd = nullptr;
      }
      t.result = reinterpret_cast<int32_t*>(d_out + st.out_offs[k]);
      t.path = reinterpret_cast<int32_t*>(d_out + st.out_offs[k] + 16);
      t.path_cap = tp.path_cap;
      t.pad_ = 0;
    });
    for (int k = 0; k < n; ++k) {
      const TaskPlan& tp = chunk[k];
      stats.cells += tp.cells;
      stats.rows += tp.R;
      stats.exported_rows += tp.job->rg.n_export;
      stats.algo_bytes += tp.L + static_cast<double>(tp.R) + 4.0 * tp.job->rg.preds.size();
    }
    st.in_bytes = in_off;
    st.out_bytes = out_off;
    const double t_launch0 = now_ms();
    stats.host_pack_ms += t_launch0 - t_pack0;
    cudaStream_t cs = st.stream;
    const PoaTask* d_tasks = reinterpret_cast<const PoaTask*>(d_in);
    bool ok = check(cudaMemcpyAsync(d_in, st.h_in, in_off, cudaMemcpyHostToDevice, cs), "H2D graphs") &&
              check(cudaMemsetAsync(st.d_counter, 0, sizeof(int), cs), "counter") &&
              check(cudaEventRecord(st.ev[0], cs), "event");
    if (ok && st.n_slot > 0)
      ok = check(poa_persistent_launch(d_tasks, st.n_slot, st.d_counter, slot_base_, slot_bytes_, ctx_->slot_flags, n_sm_, s_,
                                       ctx_->poa_threads, ctx_->ring_rows, ctx_->poa_cols, cs), "poa_persistent_kernel");
    if (ok && st.n_big > 0)
      ok = check(poa_dp_launch(d_tasks + st.n_slot, st.n_big, s_, ctx_->poa_threads, ctx_->ring_rows, ctx_->poa_cols, cs), "poa_dp_kernel");
    ok = ok && check(cudaEventRecord(st.ev[1], cs), "event");
    if (ok && st.n_big > 0) ok = check(poa_tb_launch(d_tasks + st.n_slot, st.n_big, s_, cs), "poa_tb_kernel");
    ok = ok && check(cudaEventRecord(st.ev[2], cs), "event") &&
         check(cudaMemcpyAsync(st.h_out, d_out, out_off, cudaMemcpyDeviceToHost, cs), "D2H paths") &&
         check(cudaEventRecord(st.ev[3], cs), "event");
    if (!ok) {
      blocks_.release(st.blk_off, st.blk_bytes);
      return false;
    }
    stats.launch_ms += now_ms() - t_launch0;
    stats.dp_launches += (st.n_slot > 0) + (st.n_big > 0);
    stats.tb_launches += (st.n_big > 0);
    st.busy = true;
    return true;
  }

  // Merges every path of the finished round into its graph (in parallel over the groups).
  bool collect(Stream& st) {
    const double t_merge0 = now_ms();
    st.busy = false;
    float ms = 0;
    cudaEventElapsedTime(&ms, st.ev[0], st.ev[1]); stats.dp_ms += ms;
    cudaEventElapsedTime(&ms, st.ev[1], st.ev[2]); stats.tb_ms += ms;
    cudaEventElapsedTime(&ms, st.ev[2], st.ev[3]); stats.d2h_ms += ms;
    const int n = static_cast<int>(st.inflight.size());
    stats.h2d += st.in_bytes; stats.d2h += st.out_bytes;
    stats.alignments += n;
    std::vector<double> path_pairs(n, 0.0);
    std::vector<int32_t> retries(n, 0);
    std::vector<uint8_t> overflowed(n, 0);
    parallel_for(n, threads_, [&](int k) {
      const TaskPlan& tp = st.inflight[k];
      const int32_t* res = reinterpret_cast<const int32_t*>(st.h_out + st.out_offs[k]);
      const int32_t* path = res + 4;
      const int32_t np = res[2];
      if (res[0] == -2) {   // band-limited codes did not fit the slot: repeat with a full-size block
        tp.job->force_block = true;
        overflowed[k] = 1;
        return;
      }
      if (np < 0 || static_cast<uint32_t>(np) > tp.path_cap || res[0] <= 0) {
        set_err(SVS_ERR_INTERNAL, "traceback failed (best_row=" + std::to_string(res[0]) + ", n=" + std::to_string(np) + ")");
        return;
      }
      std::vector<int32_t> nodes(np), pos(np);
      for (int32_t a = 0; a < np; ++a) {
        nodes[a] = path[2 * (np - 1 - a)];
        pos[a] = path[2 * (np - 1 - a) + 1];
      }
      path_pairs[k] = np;
      PoaJob* job = tp.job;
      job->score_per_base = static_cast<double>(res[1]) / std::max<uint32_t>(1, tp.L);
      retries[k] = res[3];
      if (job->record) {
        std::vector<int32_t> rec(2 * static_cast<size_t>(np));
        for (int32_t a = 0; a < np; ++a) { rec[2 * a] = nodes[a]; rec[2 * a + 1] = pos[a]; }
        job->record->push_back(std::move(rec));
      }
      try {
        job->graph.add_alignment(nodes.data(), pos.data(), static_cast<size_t>(np),
                                 reads_->host.data() + reads_->off[tp.seq_id], tp.L);
        ++job->next;
        job->force_block = false;
      } catch (const std::exception& ex) {
        set_err(SVS_ERR_INTERNAL, std::string("add_alignment: ") + ex.what());
      }
    });
    for (double v : path_pairs) stats.algo_bytes += 8.0 * v;
    for (int32_t v : retries) stats.prune_retries += v;
    for (uint8_t v : overflowed) { stats.slot_overflows += v; stats.alignments -= v; }
    blocks_.release(st.blk_off, st.blk_bytes);
    st.blk_bytes = 0;
    st.inflight.clear();
    stats.host_merge_ms += now_ms() - t_merge0;
    return err == 0;
  }

  svs_ctx* ctx_;
  const svs_reads* reads_;
  Scores s_;
  svs_poa_result* res_;
  bool want_msa_;
  int threads_ = 1, n_sm_ = 1, n_slots_ = 1, ctas_per_sm_ = 1;
  uint32_t cols_ = 8;
  bool persistent_ = false;
  double prune_margin_ = 0.10;   // score per base subtracted from the previous alignment's rate
  size_t slot_bytes_ = 0, block_bytes_ = 0;
  uint8_t* slot_base_ = nullptr;
  uint8_t* block_base_ = nullptr;
  ArenaAlloc blocks_;
  std::vector<Stream> streams_;
  std::mutex err_mu_;
};

int validate_scoring(svs_ctx* ctx, int algorithm, const Scores& s) {
  if (algorithm != 1)
    return fail(ctx, SVS_ERR_UNSUPPORTED, "only algorithm=1 (global alignment) is on the hot path");
  const bool convex = (s.g < s.e) && (s.g > s.q) && (s.e < s.c);
  if (!convex) return fail(ctx, SVS_ERR_UNSUPPORTED, "only the convex (two-piece) gap mode is supported");
  if (s.e - s.g > 2 || s.c - s.q > 6)
    return fail(ctx, SVS_ERR_UNSUPPORTED, "gap parameters need e-g <= 2 and c-q <= 6 (packed cell format)");
  if (s.e >= 0 || s.c >= 0 || s.m <= 0)
    return fail(ctx, SVS_ERR_UNSUPPORTED, "gap extensions must be negative and the match score positive");
  for (int v : {s.m, s.n, s.g, s.e, s.q, s.c})
    if (v > 10 || v < -10) return fail(ctx, SVS_ERR_UNSUPPORTED, "|score parameter| > 10");
  return SVS_OK;
}

int run_jobs(svs_ctx* ctx, const svs_reads* reads, std::vector<PoaJob>& jobs, const Scores& s,
             WorkerStats* total, svs_poa_result* res, bool want_msa) {
  int rc = ensure_arena(ctx);
  if (rc) return rc;
  SVS_CUDA(ctx, poa_dp_configure(ctx->poa_threads, ctx->ring_rows, ctx->poa_cols));
  if (poa_persistent_supported(ctx->poa_threads, ctx->ring_rows, ctx->poa_cols))
    SVS_CUDA(ctx, poa_persistent_configure(ctx->poa_threads, ctx->ring_rows, ctx->poa_cols));
  Scheduler sched(ctx, reads, s, res, want_msa);
  rc = sched.run(jobs);
  if (rc || sched.err) return fail(ctx, sched.err ? sched.err : rc, sched.errmsg);
  *total = sched.stats;
  return SVS_OK;
}

}  // namespace
}  // namespace svs

using namespace svs;

extern "C" {

int svs_poa_batch(svs_ctx* ctx, const svs_reads* reads, const int64_t* members, const int64_t* group_off,
                  int64_t n_groups, int algorithm, int m, int n, int g, int e, int q, int c, int want_msa,
                  svs_poa_result** out) {
  if (!ctx || !reads || !group_off || !out || n_groups < 0) return fail(ctx, SVS_ERR_ARG, "null argument");
  std::lock_guard<std::mutex> lock(ctx->mu);
  const Scores s{m, n, g, e, q, c};
  int rc = validate_scoring(ctx, algorithm, s);
  if (rc) return rc;
  SVS_CUDA(ctx, cudaSetDevice(ctx->device));
  const auto t0 = std::chrono::steady_clock::now();
  std::vector<PoaJob> jobs(static_cast<size_t>(n_groups));
  for (int64_t k = 0; k < n_groups; ++k) {
    PoaJob& j = jobs[k];
    j.group = k;
    double sum = 0, cnt = 0;
    for (int64_t a = group_off[k]; a < group_off[k + 1]; ++a) {
      const int64_t id = members[a];
      if (id < 0 || id >= reads->n) return fail(ctx, SVS_ERR_ARG, "sequence index out of range");
      j.seq_ids.push_back(id);
      sum += static_cast<double>(reads->off[id + 1] - reads->off[id]);
      cnt += 1;
    }
    const double lbar = cnt ? sum / cnt : 0;
    j.cost = cnt * lbar * lbar * (1.0 + 0.02 * cnt);
  }
  WorkerStats ws;
  std::unique_ptr<svs_poa_result> holder(new svs_poa_result());
  svs_poa_result* res = holder.get();
  res->consensus.resize(n_groups);
  res->msa.resize(n_groups);
  rc = run_jobs(ctx, reads, jobs, s, &ws, res, want_msa != 0);
  if (rc) return rc;
  holder.release();
  const double wall = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  double* st = res->stats;
  st[0] = ws.cells; st[1] = ws.alignments; st[2] = ws.dp_ms; st[3] = ws.tb_ms; st[4] = wall;
  st[5] = ws.dp_launches; st[6] = ws.tb_launches; st[7] = ws.h2d; st[8] = ws.d2h; st[9] = ws.algo_bytes;
  st[10] = ws.exported_rows; st[11] = ws.rows;
  st[12] = ws.host_wait_ms; st[13] = ws.host_merge_ms; st[14] = ws.host_plan_ms; st[15] = ws.host_pack_ms;
  st[16] = ws.refill_ms; st[17] = ws.starved; st[18] = ws.launch_ms; st[19] = ws.final_ms;
  st[20] = ws.inflight_ms; st[21] = ws.h2d_ms; st[22] = ws.d2h_ms; st[23] = ws.prune_retries;
  *out = res;
  return SVS_OK;
}

int svs_poa_result_sizes(const svs_poa_result* res, int64_t* cons_len, int64_t* msa_rows, int64_t* msa_cols) {
  if (!res) return SVS_ERR_ARG;
  for (size_t k = 0; k < res->consensus.size(); ++k) {
    if (cons_len) cons_len[k] = static_cast<int64_t>(res->consensus[k].size());
    if (msa_rows) msa_rows[k] = static_cast<int64_t>(res->msa[k].size());
    if (msa_cols) msa_cols[k] = res->msa[k].empty() ? 0 : static_cast<int64_t>(res->msa[k][0].size());
  }
  return SVS_OK;
}

int svs_poa_result_copy(const svs_poa_result* res, uint8_t* consensus, uint8_t* msa) {
  if (!res) return SVS_ERR_ARG;
  size_t co = 0, mo = 0;
  for (size_t k = 0; k < res->consensus.size(); ++k) {
    if (consensus) {
      std::memcpy(consensus + co, res->consensus[k].data(), res->consensus[k].size());
      co += res->consensus[k].size();
    }
    if (msa) {
      for (const auto& row : res->msa[k]) {
        std::memcpy(msa + mo, row.data(), row.size());
        mo += row.size();
      }
    }
  }
  return SVS_OK;
}

int svs_poa_result_stats(const svs_poa_result* res, double* stats, int n_stats) {
  if (!res || !stats) return SVS_ERR_ARG;
  for (int k = 0; k < n_stats && k < 24; ++k) stats[k] = res->stats[k];
  return SVS_OK;
}

void svs_poa_result_free(svs_poa_result* res) { delete res; }

int svs_poa_align_pairs(svs_ctx* ctx, const uint8_t* seqs, const int64_t* off, int64_t n_seqs,
                        int32_t* pair_node, int32_t* pair_pos, int64_t cap, int64_t* n_pairs,
                        int64_t* seq_pair_off) {
  if (!ctx || !off || n_seqs < 0) return fail(ctx, SVS_ERR_ARG, "null argument");
  svs_reads* reads = nullptr;
  int rc = svs_reads_upload(ctx, seqs, off, n_seqs, &reads);
  if (rc) return rc;
  std::lock_guard<std::mutex> lock(ctx->mu);
  const Scores s{5, -4, -8, -6, -10, -4};
  std::vector<PoaJob> jobs(1);
  std::vector<std::vector<int32_t>> record;
  jobs[0].record = &record;
  for (int64_t k = 0; k < n_seqs; ++k) jobs[0].seq_ids.push_back(k);
  WorkerStats ws;
  rc = run_jobs(ctx, reads, jobs, s, &ws, nullptr, false);
  svs_reads_free(reads);
  if (rc) return rc;
  int64_t total = 0;
  for (int64_t k = 0; k < n_seqs; ++k) {
    if (seq_pair_off) seq_pair_off[k] = total;
    const auto& r = record[k];
    for (size_t a = 0; a + 1 < r.size(); a += 2) {
      if (total < cap) {
        if (pair_node) pair_node[total] = r[a];
        if (pair_pos) pair_pos[total] = r[a + 1];
      }
      ++total;
    }
  }
  if (seq_pair_off) seq_pair_off[n_seqs] = total;
  if (n_pairs) *n_pairs = total;
  return SVS_OK;
}

}  // extern "C"
