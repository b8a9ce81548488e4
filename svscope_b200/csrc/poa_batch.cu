// Batched partial-order alignment: host side of svs_poa_batch (include/svscope_b200.h).
//
// A group of sequences is one graph.  Its alignments are strictly sequential (each read is
// aligned to the graph that already contains the previous ones), so the parallelism is
// across groups: every worker thread owns a share of the groups, a CUDA stream and a slice
// of the device arena, and advances all its unfinished groups by one alignment per round:
//   export rank-ordered graphs -> one H2D copy -> DP kernel (one CTA per alignment, largest
//   first) -> traceback kernel -> one D2H copy -> merge the paths into the graphs (host).
// Rounds of different workers overlap on the device, which also fills the tail of a round.
#include <algorithm>
#include <atomic>
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
  double stats[16] = {0};
};

namespace svs {
namespace {

struct PoaJob {
  int64_t group = 0;
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
  uint32_t R, L, strip, npass, path_cap;
  uint64_t ldc, ldx;
  size_t in_bytes, scratch_bytes, out_bytes;
  double cells;
};

struct WorkerStats {
  double cells = 0, alignments = 0, dp_ms = 0, tb_ms = 0, dp_launches = 0, tb_launches = 0,
         h2d = 0, d2h = 0, algo_bytes = 0, exported_rows = 0, rows = 0;
};

class Worker {
 public:
  Worker(svs_ctx* ctx, const svs_reads* reads, uint8_t* arena, size_t arena_bytes, const Scores& s)
      : ctx_(ctx), reads_(reads), arena_(arena), arena_bytes_(arena_bytes), s_(s) {}
  ~Worker() {
    if (h_in_) cudaFreeHost(h_in_);
    if (h_out_) cudaFreeHost(h_out_);
    if (ev_[0]) for (auto& e : ev_) cudaEventDestroy(e);
    if (stream_) cudaStreamDestroy(stream_);
  }
  std::vector<PoaJob*> jobs;
  WorkerStats stats;
  int err = 0;
  std::string errmsg;

  void run() {
    if (cudaSetDevice(ctx_->device) != cudaSuccess) return set_err(SVS_ERR_CUDA, "cudaSetDevice failed");
    if (!check(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking), "stream")) return;
    for (auto& e : ev_) if (!check(cudaEventCreate(&e), "event")) return;
    std::vector<TaskPlan> chunk;
    size_t used = 0;
    while (true) {
      bool any = false;
      for (PoaJob* job : jobs) {
        if (!advance_to_alignment(job)) continue;
        any = true;
        TaskPlan tp;
        if (!plan(job, &tp)) return;
        const size_t need = tp.in_bytes + tp.scratch_bytes + tp.out_bytes + sizeof(PoaTask) + 1024;
        if (need + 4096 > arena_bytes_) {
          return set_err(SVS_ERR_CAPACITY, "one alignment needs " + std::to_string(need >> 20) +
                                               " MiB, arena slice is " + std::to_string(arena_bytes_ >> 20) +
                                               " MiB (raise arena_mb or lower workers)");
        }
        if (used + need + 4096 > arena_bytes_) {
          if (!flush(chunk)) return;
          chunk.clear();
          used = 0;
          // the flushed jobs moved on; this job's export is still valid (it was not in the chunk)
        }
        chunk.push_back(tp);
        used += need;
      }
      if (!chunk.empty()) {
        if (!flush(chunk)) return;
        chunk.clear();
        used = 0;
      }
      if (!any) break;
    }
  }

 private:
  void set_err(int code, const std::string& msg) {
    err = code;
    errmsg = msg;
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
    if (worst * (static_cast<int64_t>(tp->R) + tp->L + 2) >= kMaxScoreSpan) {
      set_err(SVS_ERR_UNSUPPORTED, "alignment too large for 27-bit scores");
      return false;
    }
    const uint32_t cpp = static_cast<uint32_t>(poa_dp_cols_per_pass(ctx_->poa_threads));
    tp->npass = (tp->L + cpp - 1) / cpp;
    tp->strip = ((tp->L + tp->npass - 1) / tp->npass + 7) / 8 * 8;
    tp->ldc = (static_cast<uint64_t>(tp->L) + 7 + 7) / 8 * 8;
    tp->ldx = (static_cast<uint64_t>(tp->L) + 11 + 7) / 8 * 8;
    tp->path_cap = tp->R + tp->L + 2;
    const size_t R1 = static_cast<size_t>(g.R) + 1;
    tp->in_bytes = align_up(R1, 16) * 2 /*letter, flags*/ + align_up((R1 + 1) * 4, 16) +
                   align_up(g.preds.size() * 4, 16) + align_up(R1 * 4, 16) * 3 /*xslot,h0,node_id*/ +
                   align_up(R1 * 2, 16);
    tp->scratch_bytes = align_up(static_cast<size_t>(tp->R) * tp->ldc * 2, 256) +
                        align_up(static_cast<size_t>(g.n_export) * tp->ldx * 4, 256) +
                        align_up(R1 * 4 * 8, 256);
    tp->out_bytes = align_up(16 + static_cast<size_t>(tp->path_cap) * 8, 16);
    tp->cells = (static_cast<double>(tp->R) + 1) * (static_cast<double>(tp->L) + 1);
    return true;
  }

  bool ensure_pinned(uint8_t** buf, size_t* cap, size_t need) {
    if (*cap >= need) return true;
    if (*buf) cudaFreeHost(*buf);
    *buf = nullptr;
    *cap = 0;
    const size_t want = need + need / 2 + (1 << 20);
    if (!check(cudaMallocHost(reinterpret_cast<void**>(buf), want), "cudaMallocHost")) return false;
    *cap = want;
    return true;
  }

  bool flush(std::vector<TaskPlan>& chunk) {
    const int n = static_cast<int>(chunk.size());
    std::sort(chunk.begin(), chunk.end(), [](const TaskPlan& a, const TaskPlan& b) { return a.cells > b.cells; });
    size_t in_total = align_up(sizeof(PoaTask) * n, 256), scratch_total = 0, out_total = 0;
    for (auto& t : chunk) { in_total += t.in_bytes; scratch_total += t.scratch_bytes; out_total += t.out_bytes; }
    in_total = align_up(in_total, 256);
    out_total = align_up(out_total, 256);
    if (in_total + scratch_total + out_total > arena_bytes_)
      { set_err(SVS_ERR_INTERNAL, "arena accounting"); return false; }
    if (!ensure_pinned(&h_in_, &h_in_cap_, in_total)) return false;
    if (!ensure_pinned(&h_out_, &h_out_cap_, out_total)) return false;
    uint8_t* d_in = arena_;
    uint8_t* d_out = arena_ + in_total;
    uint8_t* d_scratch = d_out + out_total;
    PoaTask* h_tasks = reinterpret_cast<PoaTask*>(h_in_);
    size_t in_off = align_up(sizeof(PoaTask) * n, 256), out_off = 0, sc_off = 0;
    std::vector<size_t> out_offs(n);
    auto put = [&](const void* src, size_t bytes) -> const uint8_t* {
      std::memcpy(h_in_ + in_off, src, bytes);
      const uint8_t* dptr = d_in + in_off;
      in_off += align_up(bytes, 16);
      return dptr;
    };
    for (int k = 0; k < n; ++k) {
      const TaskPlan& tp = chunk[k];
      const RankedGraph& g = tp.job->rg;
      PoaTask& t = h_tasks[k];
      const size_t R1 = static_cast<size_t>(g.R) + 1;
      t.letter = put(g.letter.data(), R1);
      t.flags = put(g.flags.data(), R1);
      t.pred_off = reinterpret_cast<const uint32_t*>(put(g.pred_off.data(), (R1 + 1) * 4));
      t.preds = reinterpret_cast<const uint32_t*>(put(g.preds.data(), g.preds.size() * 4));
      t.xslot = reinterpret_cast<const int32_t*>(put(g.xslot.data(), R1 * 4));
      t.h0 = reinterpret_cast<const int32_t*>(put(g.h0.data(), R1 * 4));
      t.node_id = reinterpret_cast<const uint32_t*>(put(g.node_id.data(), R1 * 4));
      t.col0code = reinterpret_cast<const uint16_t*>(put(g.col0code.data(), R1 * 2));
      t.read = reads_->dev + reads_->off[tp.seq_id];
      t.R = tp.R; t.L = tp.L; t.strip = tp.strip; t.npass = tp.npass;
      t.codes = reinterpret_cast<uint16_t*>(d_scratch + sc_off);
      sc_off += align_up(static_cast<size_t>(tp.R) * tp.ldc * 2, 256);
      t.ldc = tp.ldc;
      t.xrows = reinterpret_cast<int32_t*>(d_scratch + sc_off);
      sc_off += align_up(static_cast<size_t>(g.n_export) * tp.ldx * 4, 256);
      t.ldx = tp.ldx;
      t.bnd = reinterpret_cast<int32_t*>(d_scratch + sc_off);
      sc_off += align_up(R1 * 4 * 8, 256);
      out_offs[k] = out_off;
      t.result = reinterpret_cast<int32_t*>(d_out + out_off);
      t.path = reinterpret_cast<int32_t*>(d_out + out_off + 16);
      t.path_cap = tp.path_cap;
      t.pad_ = 0;
      out_off += tp.out_bytes;
      stats.cells += tp.cells;
      stats.rows += tp.R;
      stats.exported_rows += g.n_export;
      stats.algo_bytes += tp.L + static_cast<double>(tp.R) + 4.0 * g.preds.size();
    }
    if (!check(cudaMemcpyAsync(d_in, h_in_, in_off, cudaMemcpyHostToDevice, stream_), "H2D graphs")) return false;
    if (!check(cudaEventRecord(ev_[0], stream_), "event")) return false;
    if (!check(poa_dp_launch(reinterpret_cast<const PoaTask*>(d_in), n, s_, ctx_->poa_threads, ctx_->ring_rows, stream_), "poa_dp_kernel")) return false;
    if (!check(cudaEventRecord(ev_[1], stream_), "event")) return false;
    if (!check(poa_tb_launch(reinterpret_cast<const PoaTask*>(d_in), n, s_, stream_), "poa_tb_kernel")) return false;
    if (!check(cudaEventRecord(ev_[2], stream_), "event")) return false;
    if (!check(cudaMemcpyAsync(h_out_, d_out, out_off, cudaMemcpyDeviceToHost, stream_), "D2H paths")) return false;
    if (!check(cudaStreamSynchronize(stream_), "poa round")) return false;
    float ms = 0;
    cudaEventElapsedTime(&ms, ev_[0], ev_[1]); stats.dp_ms += ms;
    cudaEventElapsedTime(&ms, ev_[1], ev_[2]); stats.tb_ms += ms;
    stats.dp_launches += 1; stats.tb_launches += 1;
    stats.h2d += in_off; stats.d2h += out_off;
    stats.alignments += n;
    std::vector<int32_t> nodes, pos;
    for (int k = 0; k < n; ++k) {
      const TaskPlan& tp = chunk[k];
      const int32_t* res = reinterpret_cast<const int32_t*>(h_out_ + out_offs[k]);
      const int32_t* path = res + 4;
      const int32_t np = res[2];
      if (np < 0 || static_cast<uint32_t>(np) > tp.path_cap || res[0] <= 0) {
        set_err(SVS_ERR_INTERNAL, "traceback failed (best_row=" + std::to_string(res[0]) + ", n=" + std::to_string(np) + ")");
        return false;
      }
      nodes.resize(np); pos.resize(np);
      for (int32_t a = 0; a < np; ++a) {
        nodes[a] = path[2 * (np - 1 - a)];
        pos[a] = path[2 * (np - 1 - a) + 1];
      }
      stats.algo_bytes += 8.0 * np;
      PoaJob* job = tp.job;
      if (job->record) {
        std::vector<int32_t> rec(2 * static_cast<size_t>(np));
        for (int32_t a = 0; a < np; ++a) { rec[2 * a] = nodes[a]; rec[2 * a + 1] = pos[a]; }
        job->record->push_back(std::move(rec));
      }
      try {
        job->graph.add_alignment(nodes.data(), pos.data(), static_cast<size_t>(np),
                                 reads_->host.data() + reads_->off[tp.seq_id], tp.L);
      } catch (const std::exception& ex) {
        set_err(SVS_ERR_INTERNAL, std::string("add_alignment: ") + ex.what());
        return false;
      }
      ++job->next;
    }
    return true;
  }

  svs_ctx* ctx_;
  const svs_reads* reads_;
  uint8_t* arena_;
  size_t arena_bytes_;
  Scores s_;
  cudaStream_t stream_ = nullptr;
  cudaEvent_t ev_[3] = {nullptr, nullptr, nullptr};
  uint8_t* h_in_ = nullptr;
  size_t h_in_cap_ = 0;
  uint8_t* h_out_ = nullptr;
  size_t h_out_cap_ = 0;
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
             WorkerStats* total) {
  int rc = ensure_arena(ctx);
  if (rc) return rc;
  SVS_CUDA(ctx, poa_dp_configure(ctx->poa_threads, ctx->ring_rows));
  const int nw = std::max(1, std::min<int>(ctx->workers, static_cast<int>(jobs.size())));
  const size_t slice = (ctx->arena_bytes / nw) / 256 * 256;
  std::vector<std::unique_ptr<Worker>> workers;
  for (int w = 0; w < nw; ++w)
    workers.emplace_back(new Worker(ctx, reads, static_cast<uint8_t*>(ctx->arena) + slice * w, slice, s));
  // longest-processing-time assignment of groups to workers
  std::vector<size_t> order(jobs.size());
  std::iota(order.begin(), order.end(), 0);
  std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return jobs[a].cost > jobs[b].cost; });
  std::vector<double> load(nw, 0.0);
  for (size_t idx : order) {
    const int w = static_cast<int>(std::min_element(load.begin(), load.end()) - load.begin());
    workers[w]->jobs.push_back(&jobs[idx]);
    load[w] += jobs[idx].cost + 1.0;
  }
  std::vector<std::thread> threads;
  for (int w = 1; w < nw; ++w) threads.emplace_back([&, w]() { workers[w]->run(); });
  workers[0]->run();
  for (auto& t : threads) t.join();
  for (auto& w : workers) {
    if (w->err) return fail(ctx, w->err, w->errmsg);
    total->cells += w->stats.cells; total->alignments += w->stats.alignments;
    total->dp_ms += w->stats.dp_ms; total->tb_ms += w->stats.tb_ms;
    total->dp_launches += w->stats.dp_launches; total->tb_launches += w->stats.tb_launches;
    total->h2d += w->stats.h2d; total->d2h += w->stats.d2h; total->algo_bytes += w->stats.algo_bytes;
    total->exported_rows += w->stats.exported_rows; total->rows += w->stats.rows;
  }
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
  rc = run_jobs(ctx, reads, jobs, s, &ws);
  if (rc) return rc;
  auto* res = new svs_poa_result();
  res->consensus.resize(n_groups);
  res->msa.resize(n_groups);
  {
    // consensus / MSA extraction is independent per graph
    std::atomic<int64_t> nextk{0};
    const int nt = std::max(1, std::min<int>(ctx->workers, static_cast<int>(n_groups)));
    std::vector<std::thread> th;
    auto body = [&]() {
      for (int64_t k; (k = nextk.fetch_add(1)) < n_groups;) {
        res->consensus[k] = jobs[k].graph.consensus();
        if (want_msa) res->msa[k] = jobs[k].graph.msa();
      }
    };
    for (int t = 1; t < nt; ++t) th.emplace_back(body);
    body();
    for (auto& t : th) t.join();
  }
  const double wall = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  double* st = res->stats;
  st[0] = ws.cells; st[1] = ws.alignments; st[2] = ws.dp_ms; st[3] = ws.tb_ms; st[4] = wall;
  st[5] = ws.dp_launches; st[6] = ws.tb_launches; st[7] = ws.h2d; st[8] = ws.d2h; st[9] = ws.algo_bytes;
  st[10] = ws.exported_rows; st[11] = ws.rows;
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
  for (int k = 0; k < n_stats && k < 16; ++k) stats[k] = res->stats[k];
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
  rc = run_jobs(ctx, reads, jobs, s, &ws);
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
