// Batched partial-order alignment: host side of svs_poa_batch / svs_poa_submit / svs_poa_wait
// (include/svscope_b200.h).  Replaces `spoa.poa(sequences, 1)` (reference call sites
// src/DataScanner.py:206,213 and src/DecisionMaker.py:160,171) for many sequence groups at once.
//
// A group (window) is ONE task of the window kernel (poa_kernels.cu: poa_window_kernel): a
// resident CTA aligns all its sequences one after the other with the graph kept in its scratch
// slot, so the host does nothing between submit and wait but sort the groups by cost, lay out
// the slots and read back one result record per group.  Memory tiers: the arena is cut into
// equal slots; tier 0 has one slot per resident CTA (all SMs busy), higher tiers have fewer,
// larger slots.  A window whose graph or traceback codes outgrow its slot reports that in its
// record (nothing else is affected) and is repeated in the next tier; what does not fit the
// largest tier is reported per window (status), never by failing the batch.
#include <algorithm>
#include <chrono>
#include <cstring>
#include <memory>
#include <numeric>
#include <string>
#include <vector>

#include "context.h"
#include "poa_cell.h"
#include "poa_kernels.h"
#include "poa_window.h"

struct svs_poa_result {
  svs_ctx* ctx = nullptr;
  const svs_reads* reads = nullptr;
  svs::Scores s{};
  bool want_msa = false;
  bool pending = false;
  bool debug_pairs = false;
  int64_t n_groups = 0;
  std::vector<int64_t> members;
  std::vector<svs::WinDesc> desc;
  std::vector<svs::WinResult> res;          // by group
  std::vector<int> out_buf;                 // index into out_bufs per group
  std::vector<uint64_t> est_codes;          // estimated traceback-code bytes of the largest alignment of the group
  std::vector<uint8_t*> out_bufs;           // device output arenas (one per launch round)
  std::vector<int> round_groups;            // groups of the running round
  int tier = 0;
  // device-side inputs of the running round
  int64_t* d_members = nullptr;
  svs::WinDesc* d_desc = nullptr;
  int32_t* d_order = nullptr;
  svs::WinResult* d_res = nullptr;
  int* d_counter = nullptr;                 // [0] window counter, [2..3] output cursor (64 bit)
  int32_t* d_pairs = nullptr;
  int64_t* d_pair_cnt = nullptr;
  int64_t pairs_cap = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::chrono::steady_clock::time_point t_submit;
  double stats[40] = {0};
  double h2d = 0, d2h = 0, kernel_ms = 0, launches = 0;
};

namespace svs {
namespace {

struct Tier {
  int n_slots;
  size_t slot_bytes;
};

int window_cps(svs_ctx* ctx) { return poa_window_ctas_per_sm(ctx->poa_threads, ctx->ring_rows, ctx->poa_cols); }

// tier 0: one slot per resident CTA; then 1 per SM, 1 per 4 SMs, 8, 2, 1 slots
std::vector<Tier> make_tiers(svs_ctx* ctx) {
  std::vector<Tier> t;
  const int cps = std::max(1, window_cps(ctx));
  int sm = std::max(1, ctx->sm_count);
  if (const char* sl = getenv("SVS_SM_LIMIT")) {   // profiling aid: slots for the SMs that take work only
    if (atoi(sl) > 0) sm = std::min(sm, atoi(sl));
  }
  for (int n : {sm * cps, sm, std::max(1, sm / 4), 8, 2, 1}) {
    if (!t.empty() && n >= t.back().n_slots) continue;
    t.push_back(Tier{n, (ctx->arena_bytes / static_cast<size_t>(n)) / 4096 * 4096});
  }
  return t;
}

uint64_t fixed_bytes(const WinCaps& c) {
  WinMem m;
  return win_layout(nullptr, 0, c, &m);
}

// capacity estimate of a group: `full` = no overflow possible (every base a new node)
WinCaps estimate_caps(const svs_reads* reads, const int64_t* mem, int64_t n, bool full) {
  WinCaps c;
  c.nseq = static_cast<uint32_t>(n);
  uint64_t first = 0;
  for (int64_t k = 0; k < n; ++k) {
    const uint64_t len = static_cast<uint64_t>(reads->off[mem[k] + 1] - reads->off[mem[k]]);
    c.sumlen += len;
    if (len > c.lmax) c.lmax = static_cast<uint32_t>(len);
    if (first == 0) first = len;
  }
  uint64_t v = c.sumlen;
  if (!full && c.sumlen > (1u << 16)) v = std::min<uint64_t>(c.sumlen, first + c.lmax + (c.sumlen - first) / 4 + 4096);
  c.vcap = static_cast<uint32_t>(std::min<uint64_t>(v + 1, 0x7fffff00u));
  c.ecap = static_cast<uint32_t>(std::min<uint64_t>(full ? c.sumlen + 1 : std::min<uint64_t>(c.sumlen + 1, 2ull * v + 1), 0x7fffff00u));
  return c;
}

double group_cost(const WinCaps& c) {
  const double n = c.nseq, lbar = c.nseq ? static_cast<double>(c.sumlen) / c.nseq : 0;
  return n * lbar * lbar * (1.0 + 0.02 * n);
}

// (stream-ordered frees: cudaFree would wait for every kernel on the device, i.e. for the window
// kernels of the other sub-batches that are still running)
void free_round(svs_poa_result* r) {
  if (r->d_members) cudaFreeAsync(r->d_members, r->stream);
  if (r->d_desc) cudaFreeAsync(r->d_desc, r->stream);
  if (r->d_order) cudaFreeAsync(r->d_order, r->stream);
  if (r->d_res) cudaFreeAsync(r->d_res, r->stream);
  if (r->d_counter) cudaFreeAsync(r->d_counter, r->stream);
  r->d_members = nullptr; r->d_desc = nullptr; r->d_order = nullptr; r->d_res = nullptr; r->d_counter = nullptr;
}

// Launches the window kernel for `groups` (indices into r->desc) in `tier`.  Asynchronous.
int launch_round(svs_poa_result* r, const std::vector<int>& groups, int tier_idx, double out_factor) {
  svs_ctx* ctx = r->ctx;
  const std::vector<Tier> tiers = make_tiers(ctx);
  const Tier& tier = tiers[tier_idx];
  const int n = static_cast<int>(groups.size());
  r->round_groups = groups;
  r->tier = tier_idx;
  if (n == 0) return SVS_OK;
  // launch order: largest first
  std::vector<int32_t> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::vector<double> cost(n);
  for (int k = 0; k < n; ++k) cost[k] = group_cost(r->desc[groups[k]].caps);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
  std::vector<WinDesc> desc(n);
  uint64_t out_need = 4096;
  for (int k = 0; k < n; ++k) {
    desc[k] = r->desc[groups[k]];
    const WinCaps& c = desc[k].caps;
    const double cols = std::min<double>(static_cast<double>(c.sumlen), out_factor * c.lmax + 1024);
    out_need += static_cast<uint64_t>((r->want_msa ? c.nseq * cols : 0) + c.lmax + cols + 64);
  }
  free_round(r);
  SVS_CUDA(ctx, cudaMallocAsync(reinterpret_cast<void**>(&r->d_desc), sizeof(WinDesc) * n, r->stream));
  SVS_CUDA(ctx, cudaMallocAsync(reinterpret_cast<void**>(&r->d_order), sizeof(int32_t) * n, r->stream));
  SVS_CUDA(ctx, cudaMallocAsync(reinterpret_cast<void**>(&r->d_res), sizeof(WinResult) * n, r->stream));
  SVS_CUDA(ctx, cudaMallocAsync(reinterpret_cast<void**>(&r->d_counter), 64, r->stream));
  SVS_CUDA(ctx, cudaMallocAsync(reinterpret_cast<void**>(&r->d_members), sizeof(int64_t) * std::max<size_t>(1, r->members.size()), r->stream));
  uint8_t* d_out = nullptr;
  SVS_CUDA(ctx, cudaMallocAsync(reinterpret_cast<void**>(&d_out), out_need, r->stream));
  r->out_bufs.push_back(d_out);
  cudaStream_t st = r->stream;
  SVS_CUDA(ctx, cudaMemcpyAsync(r->d_desc, desc.data(), sizeof(WinDesc) * n, cudaMemcpyHostToDevice, st));
  SVS_CUDA(ctx, cudaMemcpyAsync(r->d_order, order.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
  SVS_CUDA(ctx, cudaMemcpyAsync(r->d_members, r->members.data(), sizeof(int64_t) * r->members.size(), cudaMemcpyHostToDevice, st));
  SVS_CUDA(ctx, cudaMemsetAsync(r->d_counter, 0, 64, st));
  SVS_CUDA(ctx, cudaMemsetAsync(r->d_res, 0xff, sizeof(WinResult) * n, st));
  r->h2d += sizeof(WinDesc) * n + sizeof(int32_t) * n + sizeof(int64_t) * r->members.size();
  WinParams p{};
  p.reads = r->reads->dev;
  p.read_off = r->reads->dev_off;
  p.members = r->d_members;
  p.desc = r->d_desc;
  p.order = r->d_order;
  p.n_windows = n;
  p.counter = r->d_counter;
  p.slot_base = static_cast<uint8_t*>(ctx->arena);
  p.slot_bytes = tier.slot_bytes;
  p.slot_flags = ctx->slot_flags;
  p.n_slots = tier.n_slots;
  p.out_base = d_out;
  p.out_cap = out_need;
  p.out_cursor = reinterpret_cast<unsigned long long*>(r->d_counter + 2);
  p.results = r->d_res;
  p.pairs_out = r->d_pairs;
  p.pair_cnt = r->d_pair_cnt;
  p.s = r->s;
  p.ring_rows = ctx->ring_rows;
  p.prune = ctx->prune;
  p.want_msa = r->want_msa ? 1 : 0;
  p.prune_margin = 0.10f;
  p.dp_version = ctx->dp_kernel;
  if (const char* pm = getenv("SVS_PRUNE_MARGIN")) p.prune_margin = static_cast<float>(atof(pm));
  int grid = std::min(n, tier.n_slots);
  if (const char* sl = getenv("SVS_SM_LIMIT")) {   // profiling aid (scripts/ncu_dp.sh): see WinParams::sm_limit
    p.sm_limit = atoi(sl);
    if (p.sm_limit > 0) grid = std::max(1, ctx->sm_count) * std::max(1, window_cps(ctx));
  }
  SVS_CUDA(ctx, cudaEventRecord(r->ev0, st));
  SVS_CUDA(ctx, poa_window_launch(p, grid, ctx->poa_threads, ctx->poa_cols, st));
  SVS_CUDA(ctx, cudaEventRecord(r->ev1, st));
  r->launches += 1;
  return SVS_OK;
}

// Waits for the running round and files its records; returns the groups to repeat.
int collect_round(svs_poa_result* r, std::vector<int>* again) {
  svs_ctx* ctx = r->ctx;
  again->clear();
  const int n = static_cast<int>(r->round_groups.size());
  if (n == 0) return SVS_OK;
  SVS_CUDA(ctx, cudaEventSynchronize(r->ev1));
  float ms = 0;
  cudaEventElapsedTime(&ms, r->ev0, r->ev1);
  r->kernel_ms += ms;
  std::vector<WinResult> res(n);
  SVS_CUDA(ctx, cudaMemcpy(res.data(), r->d_res, sizeof(WinResult) * n, cudaMemcpyDeviceToHost));
  r->d2h += sizeof(WinResult) * n;
  const int buf = static_cast<int>(r->out_bufs.size()) - 1;
  for (int k = 0; k < n; ++k) {
    const int g = r->round_groups[k];
    WinResult& w = res[k];
    if (w.status < 0 || w.status > kWinIndeg) w.status = kWinPending;   // never written: kernel fault upstream
    // accumulate the work counters over the rounds of this group
    const WinResult prev = r->res[g];
    r->res[g] = w;
    r->out_buf[g] = buf;
    if (prev.status != kWinPending) {   // a repeated window: keep counting what was spent
      r->res[g].cells += prev.cells; r->res[g].eval_cells += prev.eval_cells; r->res[g].n_align += prev.n_align; r->res[g].retries += prev.retries;
      r->res[g].rows += prev.rows; r->res[g].exported += prev.exported;
      r->res[g].read_bases += prev.read_bases; r->res[g].path_steps += prev.path_steps; r->res[g].pred_entries += prev.pred_entries;
      for (int c = 0; c < 8; ++c) r->res[g].cyc[c] += prev.cyc[c];
      for (int c = 0; c < 4; ++c) r->res[g].warp_cyc[c] += prev.warp_cyc[c];
    }
    switch (w.status) {
      case kWinNodeCap: case kWinEdgeCap: case kWinStackCap: case kWinCodesCap: case kWinOutCap:
        again->push_back(g);
        break;
      default: break;
    }
  }
  return SVS_OK;
}

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

int submit(svs_ctx* ctx, const svs_reads* reads, const int64_t* members, const int64_t* group_off, int64_t n_groups,
           const Scores& s, bool want_msa, bool debug_pairs, svs_poa_result** out) {
  int rc = ensure_arena(ctx);
  if (rc) return rc;
  if (window_cps(ctx) <= 0) return fail(ctx, SVS_ERR_ARG, "poa_threads / poa_cols / ring_rows: no such kernel configuration");
  SVS_CUDA(ctx, poa_window_configure(ctx->poa_threads, ctx->ring_rows, ctx->poa_cols));
  std::unique_ptr<svs_poa_result> r(new svs_poa_result());
  r->ctx = ctx; r->reads = reads; r->s = s; r->want_msa = want_msa; r->n_groups = n_groups;
  r->debug_pairs = debug_pairs;
  r->t_submit = std::chrono::steady_clock::now();
  r->members.assign(members, members + group_off[n_groups]);
  for (int64_t id : r->members)
    if (id < 0 || id >= reads->n) return fail(ctx, SVS_ERR_ARG, "sequence index out of range");
  r->desc.resize(n_groups);
  r->res.resize(n_groups);
  r->out_buf.assign(n_groups, -1);
  r->est_codes.assign(n_groups, 0);
  const std::vector<Tier> tiers = make_tiers(ctx);
  std::vector<int> first;
  int64_t pairs_total = 0;
  for (int64_t g = 0; g < n_groups; ++g) {
    WinDesc& d = r->desc[g];
    d.member_begin = group_off[g];
    d.caps = estimate_caps(reads, members + group_off[g], group_off[g + 1] - group_off[g], false);
    d.pairs_off = -1;
    if (debug_pairs) {
      d.pairs_off = pairs_total;
      pairs_total += static_cast<int64_t>(d.caps.nseq) * (static_cast<int64_t>(d.caps.vcap) + d.caps.lmax + 2);
    }
    {   // largest alignment of the group: the last read against a graph of first + ~8 % of the other bases,
        // about 0.6 B of codes per nominal cell (band-limited rows, 1-2 B per evaluated cell)
      uint64_t first = 0;
      for (int64_t k = group_off[g]; k < group_off[g + 1] && first == 0; ++k)
        first = static_cast<uint64_t>(reads->off[members[k] + 1] - reads->off[members[k]]);
      const double rows = std::min<double>(d.caps.vcap, static_cast<double>(first) + 0.08 * static_cast<double>(d.caps.sumlen - first));
      r->est_codes[g] = static_cast<uint64_t>(0.6 * rows * static_cast<double>(d.caps.lmax));
    }
    std::memset(&r->res[g], 0, sizeof(WinResult));
    r->res[g].status = kWinPending;
    first.push_back(static_cast<int>(g));
  }
  SVS_CUDA(ctx, cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking));
  SVS_CUDA(ctx, cudaEventCreate(&r->ev0));
  SVS_CUDA(ctx, cudaEventCreate(&r->ev1));
  if (debug_pairs) {
    r->pairs_cap = pairs_total;
    SVS_CUDA(ctx, cudaMallocAsync(reinterpret_cast<void**>(&r->d_pairs), sizeof(int32_t) * 2 * std::max<int64_t>(1, pairs_total), r->stream));
    SVS_CUDA(ctx, cudaMallocAsync(reinterpret_cast<void**>(&r->d_pair_cnt), sizeof(int64_t) * std::max<size_t>(1, r->members.size()), r->stream));
    SVS_CUDA(ctx, cudaMemsetAsync(r->d_pair_cnt, 0, sizeof(int64_t) * std::max<size_t>(1, r->members.size()), r->stream));
  }
  // windows whose fixed part does not even fit a tier-0 slot start from the first tier that holds them
  std::vector<int> t0;
  for (int g : first) {
    if (fixed_bytes(r->desc[g].caps) + r->est_codes[g] + (8u << 20) <= tiers[0].slot_bytes) t0.push_back(g);
  }
  rc = launch_round(r.get(), t0, 0, 2.5);
  if (rc) return rc;
  r->pending = true;
  *out = r.release();
  return SVS_OK;
}

int wait(svs_poa_result* r) {
  svs_ctx* ctx = r->ctx;
  if (!r->pending) return SVS_OK;
  r->pending = false;
  const std::vector<Tier> tiers = make_tiers(ctx);
  std::vector<int> again;
  int rc = collect_round(r, &again);
  if (rc) return rc;
  // groups that never ran in tier 0 (fixed part too large)
  for (int64_t g = 0; g < r->n_groups; ++g)
    if (r->res[g].status == kWinPending && std::find(again.begin(), again.end(), static_cast<int>(g)) == again.end())
      again.push_back(static_cast<int>(g));
  int tier = 0;
  double out_factor = 2.5;
  while (!again.empty()) {
    // what failed, and the smallest tier that can hold every repeated window
    bool only_out = true;
    for (int g : again) only_out = only_out && r->res[g].status == kWinOutCap;
    if (only_out) out_factor *= 4;
    else ++tier;
    if (tier >= static_cast<int>(tiers.size()) || out_factor > 200) break;
    std::vector<int> run;
    for (int g : again) {
      WinDesc& d = r->desc[g];
      const int st = r->res[g].status;
      if (st == kWinNodeCap || st == kWinEdgeCap || st == kWinStackCap)
        d.caps = estimate_caps(r->reads, r->members.data() + d.member_begin, d.caps.nseq, true);
      const uint64_t need = fixed_bytes(d.caps) + (8u << 20) +
                            (st == kWinCodesCap ? r->res[g].need_bytes + (r->res[g].need_bytes >> 3)
                                                : (st == kWinPending ? r->est_codes[g] : 0));
      if (need <= tiers[tier].slot_bytes || tier + 1 == static_cast<int>(tiers.size())) run.push_back(g);
    }
    std::vector<int> skipped;
    for (int g : again) if (std::find(run.begin(), run.end(), g) == run.end()) skipped.push_back(g);
    if (!run.empty()) {
      // larger slots alias the slots of every other launch: this tier runs alone on the arena
      SVS_CUDA(ctx, cudaDeviceSynchronize());
      rc = launch_round(r, run, tier, out_factor);
      if (rc) return rc;
      rc = collect_round(r, &again);
      if (rc) return rc;
    } else {
      again.clear();
    }
    again.insert(again.end(), skipped.begin(), skipped.end());
  }
  free_round(r);
  // totals
  double* st = r->stats;
  std::fill(st, st + 40, 0.0);
  for (const WinResult& w : r->res) {
    st[0] += static_cast<double>(w.cells);
    st[12] += static_cast<double>(w.eval_cells);
    st[1] += w.n_align;
    st[10] += static_cast<double>(w.exported);
    st[11] += static_cast<double>(w.rows);
    st[23] += w.retries;
    for (int c = 0; c < 8; ++c) st[24 + c] += static_cast<double>(w.cyc[c]);
    for (int c = 0; c < 4; ++c) st[33 + c] += static_cast<double>(w.warp_cyc[c]);
    if (w.status != kWinOk) st[32] += 1;
  }
  st[2] = r->kernel_ms;
  st[4] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - r->t_submit).count();
  st[5] = r->launches;
  st[7] = r->h2d;
  st[8] = r->d2h;
  // algorithmic bytes: read + rank-ordered graph (letter + 4 B per in-edge) + 8 B per path step
  for (const WinResult& w : r->res)
    st[9] += static_cast<double>(w.read_bases) + static_cast<double>(w.rows) + 4.0 * static_cast<double>(w.pred_entries) + 8.0 * static_cast<double>(w.path_steps);
  return SVS_OK;
}

}  // namespace
}  // namespace svs

using namespace svs;

extern "C" {

int svs_poa_submit(svs_ctx* ctx, const svs_reads* reads, const int64_t* members, const int64_t* group_off,
                   int64_t n_groups, int algorithm, int m, int n, int g, int e, int q, int c, int want_msa,
                   svs_poa_result** out) {
  if (!ctx || !reads || !group_off || !out || n_groups < 0) return fail(ctx, SVS_ERR_ARG, "null argument");
  std::lock_guard<std::mutex> lock(ctx->mu);
  const Scores s{m, n, g, e, q, c};
  int rc = validate_scoring(ctx, algorithm, s);
  if (rc) return rc;
  SVS_CUDA(ctx, cudaSetDevice(ctx->device));
  return submit(ctx, reads, members, group_off, n_groups, s, want_msa != 0, false, out);
}

int svs_poa_wait(svs_poa_result* res) {
  if (!res) return SVS_ERR_ARG;
  std::lock_guard<std::mutex> lock(res->ctx->mu);
  SVS_CUDA(res->ctx, cudaSetDevice(res->ctx->device));
  return wait(res);
}

int svs_poa_batch(svs_ctx* ctx, const svs_reads* reads, const int64_t* members, const int64_t* group_off,
                  int64_t n_groups, int algorithm, int m, int n, int g, int e, int q, int c, int want_msa,
                  svs_poa_result** out) {
  if (!out) return fail(ctx, SVS_ERR_ARG, "null argument");
  svs_poa_result* r = nullptr;
  int rc = svs_poa_submit(ctx, reads, members, group_off, n_groups, algorithm, m, n, g, e, q, c, want_msa, &r);
  if (rc) return rc;
  rc = svs_poa_wait(r);
  if (rc) { svs_poa_result_free(r); return rc; }
  *out = r;
  return SVS_OK;
}

int svs_poa_result_status(const svs_poa_result* res, int32_t* status) {
  if (!res || !status) return SVS_ERR_ARG;
  for (int64_t k = 0; k < res->n_groups; ++k) status[k] = res->res[k].status;
  return SVS_OK;
}

int svs_poa_result_sizes(const svs_poa_result* res, int64_t* cons_len, int64_t* msa_rows, int64_t* msa_cols) {
  if (!res) return SVS_ERR_ARG;
  for (int64_t k = 0; k < res->n_groups; ++k) {
    const WinResult& w = res->res[k];
    const bool ok = w.status == kWinOk;
    if (cons_len) cons_len[k] = ok ? w.cons_len : 0;
    if (msa_rows) msa_rows[k] = ok ? w.msa_rows : 0;
    if (msa_cols) msa_cols[k] = ok ? w.msa_cols : 0;
  }
  return SVS_OK;
}

int svs_poa_result_copy(const svs_poa_result* res, uint8_t* consensus, uint8_t* msa) {
  if (!res) return SVS_ERR_ARG;
  svs_ctx* ctx = res->ctx;
  SVS_CUDA(ctx, cudaSetDevice(ctx->device));
  size_t co = 0, mo = 0;
  for (int64_t k = 0; k < res->n_groups; ++k) {
    const WinResult& w = res->res[k];
    if (w.status != kWinOk || res->out_buf[k] < 0) continue;
    const uint8_t* base = res->out_bufs[res->out_buf[k]] + w.out_off;
    const size_t mb = static_cast<size_t>(w.msa_rows) * w.msa_cols;
    if (msa && mb) SVS_CUDA(ctx, cudaMemcpyAsync(msa + mo, base, mb, cudaMemcpyDeviceToHost, res->stream));
    if (consensus && w.cons_len)
      SVS_CUDA(ctx, cudaMemcpyAsync(consensus + co, base + mb, w.cons_len, cudaMemcpyDeviceToHost, res->stream));
    mo += mb;
    co += w.cons_len;
  }
  SVS_CUDA(ctx, cudaStreamSynchronize(res->stream));
  return SVS_OK;
}

int svs_poa_result_stats(const svs_poa_result* res, double* stats, int n_stats) {
  if (!res || !stats) return SVS_ERR_ARG;
  for (int k = 0; k < n_stats && k < 40; ++k) stats[k] = res->stats[k];
  return SVS_OK;
}

void svs_poa_result_free(svs_poa_result* res) {
  if (!res) return;
  cudaSetDevice(res->ctx->device);
  if (res->pending && res->ev1) cudaEventSynchronize(res->ev1);
  free_round(res);
  for (uint8_t* b : res->out_bufs) cudaFreeAsync(b, res->stream);
  if (res->d_pairs) cudaFreeAsync(res->d_pairs, res->stream);
  if (res->d_pair_cnt) cudaFreeAsync(res->d_pair_cnt, res->stream);
  if (res->ev0) cudaEventDestroy(res->ev0);
  if (res->ev1) cudaEventDestroy(res->ev1);
  if (res->stream) cudaStreamDestroy(res->stream);
  delete res;
}

int svs_poa_align_pairs(svs_ctx* ctx, const uint8_t* seqs, const int64_t* off, int64_t n_seqs,
                        int32_t* pair_node, int32_t* pair_pos, int64_t cap, int64_t* n_pairs,
                        int64_t* seq_pair_off) {
  if (!ctx || !off || n_seqs < 0) return fail(ctx, SVS_ERR_ARG, "null argument");
  svs_reads* reads = nullptr;
  int rc = svs_reads_upload(ctx, seqs, off, n_seqs, &reads);
  if (rc) return rc;
  std::unique_lock<std::mutex> lock(ctx->mu);
  const Scores s{5, -4, -8, -6, -10, -4};
  std::vector<int64_t> members(n_seqs);
  std::iota(members.begin(), members.end(), 0);
  const int64_t goff[2] = {0, n_seqs};
  svs_poa_result* r = nullptr;
  rc = submit(ctx, reads, members.data(), goff, 1, s, false, true, &r);
  if (!rc) rc = wait(r);
  int64_t total = 0;
  if (!rc && r->res[0].status != kWinOk) rc = fail(ctx, SVS_ERR_CAPACITY, "window status " + std::to_string(r->res[0].status));
  if (!rc) {
    std::vector<int64_t> cnt(n_seqs);
    cudaMemcpy(cnt.data(), r->d_pair_cnt, sizeof(int64_t) * n_seqs, cudaMemcpyDeviceToHost);
    int64_t all = 0;
    for (int64_t k = 0; k < n_seqs; ++k) all += cnt[k];
    std::vector<int32_t> pairs(2 * std::max<int64_t>(1, all));
    cudaMemcpy(pairs.data(), r->d_pairs, sizeof(int32_t) * 2 * all, cudaMemcpyDeviceToHost);
    int64_t src = 0;
    for (int64_t k = 0; k < n_seqs; ++k) {
      if (seq_pair_off) seq_pair_off[k] = total;
      for (int64_t a = 0; a < cnt[k]; ++a, ++src) {
        if (total < cap) {
          if (pair_node) pair_node[total] = pairs[2 * src];
          if (pair_pos) pair_pos[total] = pairs[2 * src + 1];
        }
        ++total;
      }
    }
    if (seq_pair_off) seq_pair_off[n_seqs] = total;
    if (n_pairs) *n_pairs = total;
  }
  lock.unlock();
  if (r) svs_poa_result_free(r);
  svs_reads_free(reads);
  return rc;
}

}  // extern "C"
