// TEST INFRASTRUCTURE: sequential CPU emulation of the GPU alignment path.  It runs the
// product's host graph (poa_graph.cpp) and the product's cell arithmetic / traceback walker
// (poa_cell.h) with the same information loss as the kernels (predecessor rows are only
// visible as packed words; E opened from A), so that the algorithmic equivalence with the
// oracle's five-matrix equality traceback can be checked without a GPU.  Not shipped.
#include <chrono>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../svscope_b200/csrc/poa_cell.h"
#include "poa_graph.h"

using namespace svs;

struct Emu {
  int prune = 0;         // 0 off, >0: half-width of the first (narrow) pass
  int dyn = 0;           // 1: dynamic score-bound pruning with a guessed lower bound (lb_ratio * L)
  double lb_ratio = 4.0;
  int dyn_ext = -1;      // -1: provable right extension of every row; >= 0: that many chunks of lookahead,
                         //     a row whose last computed chunk is still relevant repeats the alignment with more
  int overflows = 0;
  bool overflow = false;
  int retries = 0;
  double kept_cells = 0, all_cells = 0, static_cells = 0;  // static = cells the static score-bound band would keep
  PoaGraph graph;
  PoaScoring sc;
  uint32_t ring_rows = 4;
  std::vector<int32_t> last;  // forward pairs
  RankedGraph rg;
  // > 0: run the product's warp-pipelined DP source (poa_dp2.cuh) on that many OS threads (dp2_threads.cpp)
  int warp_threads = 0;
  int warp_prune = 0;
  int warp_retries = 0;
  int32_t last_score = 0;
  uint32_t last_len = 0;
};

namespace svs {
int dp2_threads_align(const RankedGraph& G, const PoaScoring& sc, const uint8_t* read, uint32_t L, int threads, int ring_rows,
                      bool prune, int32_t lb_guess, std::vector<int32_t>* rev_pairs, int32_t* score, int* retries);
}

static bool emu_align_dyn(Emu* E, const uint8_t* read, uint32_t L, int32_t lb, int32_t* score_out);

static void emu_align(Emu* E, const uint8_t* read, uint32_t L) {
  E->last.clear();
  if (E->graph.empty() || L == 0) return;
  E->graph.export_ranked(E->sc, E->ring_rows, &E->rg);
  if (E->warp_threads > 0) {
    // poa_kernels.cu: pruning from a guessed lower bound (score per base of the previous read minus a margin);
    // the size thresholds are lowered here so that small test inputs run the banded path too
    const bool prune = E->warp_prune != 0 && L >= 64 && E->rg.R >= 64;
    const double spb = E->last_len ? static_cast<double>(E->last_score) / E->last_len : 4.0;
    const int32_t lb_guess = static_cast<int32_t>((spb - 0.10) * static_cast<double>(L)) - 40;
    std::vector<int32_t> rev;
    int32_t score = 0;
    int retries = 0;
    const int rc = dp2_threads_align(E->rg, E->sc, read, L, E->warp_threads, static_cast<int>(E->ring_rows), prune, lb_guess, &rev,
                                     &score, &retries);
    if (rc != 0) { E->last.assign(2, -999); return; }   // shows up as a mismatch in the test
    E->warp_retries += retries;
    E->last_score = score; E->last_len = L;
    for (int64_t k = static_cast<int64_t>(rev.size() / 2) - 1; k >= 0; --k) {
      E->last.push_back(rev[2 * k]);
      E->last.push_back(rev[2 * k + 1]);
    }
    return;
  }
  if (E->dyn) {
    int32_t lb = static_cast<int32_t>(E->lb_ratio * L), score = 0;
    const int ext0 = E->dyn_ext;
    bool ok = false;
    for (;;) {
      E->overflow = false;
      ok = emu_align_dyn(E, read, L, lb, &score);
      if (E->overflow) {       // lookahead too short somewhere: widen it (at most up to the provable one)
        ++E->overflows;
        E->dyn_ext = E->dyn_ext < 64 ? (E->dyn_ext + 1) * 4 : -1;
        continue;
      }
      if (!ok || score < lb) {   // the guess was above the optimum: repeat with a feasible score (or unpruned)
        ++E->retries;
        lb = ok ? score : INT32_MIN;
        continue;
      }
      break;
    }
    E->dyn_ext = ext0;
    return;
  }
  const RankedGraph& G = E->rg;
  const Scores s{E->sc.m, E->sc.n, E->sc.g, E->sc.e, E->sc.q, E->sc.c};
  const uint32_t R = G.R;
  const uint64_t W = L + 1;
  std::vector<int32_t> P((R + 1) * W);          // packed cells, all rows
  // code rows: 1 byte per cell for single-predecessor rows, 2 bytes otherwise (as on the device)
  const uint32_t w1 = (L + 7 + 15) / 16 * 16, w2 = (L + 7 + 7) / 8 * 8 * 2;
  const uint64_t n1_total = G.single_before[R + 1];
  std::vector<uint8_t> codes(n1_total * w1 + (R - n1_total) * w2 + 64);
  P[0] = pack_cell(0, kNeg, kNeg);
  for (uint32_t j = 1; j <= L; ++j) P[j] = pack_cell(row0_h(s, j), kNeg, kNeg);
  PredLutEntry lut[32];
  for (int low = 0; low < 32; ++low) lut[low] = pred_lut_entry(s, low);
  // depths for the pruning bounds
  std::vector<int32_t> dmin(R + 1, 0), dmax(R + 1, 0), smin(R + 1, 0), smax(R + 1, 0);
  std::vector<uint8_t> has_succ(R + 1, 0);
  for (uint32_t i = 1; i <= R; ++i) {
    int32_t lo = INT32_MAX, hi = 0;
    for (uint32_t k = G.pred_off[i]; k < G.pred_off[i + 1]; ++k) {
      const uint32_t p = G.preds[k];
      lo = std::min(lo, dmin[p]); hi = std::max(hi, dmax[p]);
    }
    dmin[i] = lo + 1; dmax[i] = hi + 1;
  }
  for (uint32_t i = R; i >= 1; --i) {
    if (!has_succ[i]) { smin[i] = 0; smax[i] = 0; }
    for (uint32_t k = G.pred_off[i]; k < G.pred_off[i + 1]; ++k) {
      const uint32_t p = G.preds[k];
      if (p == 0) continue;
      if (!has_succ[p]) { smin[p] = smin[i] + 1; smax[p] = smax[i] + 1; has_succ[p] = 1; }
      else { smin[p] = std::min(smin[p], smin[i] + 1); smax[p] = std::max(smax[p], smax[i] + 1); }
    }
  }
  int32_t best = INT32_MIN;
  uint32_t best_row = 0;
  std::vector<int32_t> lo_col(R + 1, 1), hi_col(R + 1, static_cast<int32_t>(L));
  const int n_pass = E->prune > 0 ? 2 : 1;
  for (int pass = 0; pass < n_pass; ++pass) {
    if (E->prune > 0) {
      if (pass == 0) {
        for (uint32_t i = 1; i <= R; ++i) { lo_col[i] = std::max(1, dmin[i] - E->prune); hi_col[i] = std::min<int32_t>(L, dmax[i] + E->prune); }
      } else {
        const int32_t LB = best;  // score of a feasible alignment (or INT32_MIN: keep everything)
        for (uint32_t i = 1; i <= R; ++i) {
          int32_t lo = L + 1, hi = 0;
          for (int32_t j = 1; j <= static_cast<int32_t>(L); ++j) {
            if (LB == INT32_MIN || cell_bound(s, dmin[i], dmax[i], smin[i], smax[i], j, L) >= LB) { lo = std::min(lo, j); hi = std::max(hi, j); }
          }
          lo_col[i] = lo; hi_col[i] = hi;
          E->kept_cells += std::max(0, hi - lo + 1);
          E->all_cells += L;
        }
      }
    }
    best = INT32_MIN; best_row = 0;
    const int32_t NEGW = pack_cell(kNegBand, kNeg, kNeg);
    for (uint32_t i = 1; i <= R; ++i) {
      const bool col0_in = true;  // column 0 is exact (host values)
      (void)col0_in;
      P[i * W] = pack_cell(G.h0[i], kNeg, kNeg);
      RowCarry cy{G.h0[i], kNeg, kNeg, G.h0[i]};
      for (uint32_t j = 1; j <= L; ++j) {
        if (E->prune > 0 && (static_cast<int32_t>(j) < lo_col[i] || static_cast<int32_t>(j) > hi_col[i])) {
          P[i * W + j] = NEGW;
          cy = RowCarry{kNegBand, kNeg, kNeg, kNegBand};
          continue;
        }
        CellAcc a;
        const int32_t sub = (G.letter[i] == read[j - 1]) ? s.m : s.n;
        const bool single = (G.pred_off[i + 1] - G.pred_off[i] == 1);
        int32_t H;
        uint16_t cd;
        if (single) {  // the kernels' fast path
          const uint32_t p = G.preds[G.pred_off[i]];
          const int32_t w = P[p * W + j];
          const PredLutEntry& t = lut[w & 31];
          cell_pred_single(a, w, unpack_h(P[p * W + j - 1]), sub, t.sf, t.so, t.sm);
          cd = static_cast<uint16_t>(cell_finish_single(a, cy, s, H));
        } else {
          cell_key_init(a);
          for (uint32_t k = G.pred_off[i]; k < G.pred_off[i + 1]; ++k) {
            const uint32_t p = G.preds[k];
            const int32_t w = P[p * W + j];
            const PredLutEntry& t = lut[w & 31];
            cell_pred_key(a, 31 - static_cast<int32_t>(k - G.pred_off[i]), w, P[p * W + j - 1], t.tf, t.to, t.tv);
          }
          cell_key_add_sub(a, sub);
          int32_t Fo, Oo;
          cd = cell_finish_key(a, cy, s, H, Fo, Oo);
          a.Fm = Fo; a.Om = Oo;
        }
        const uint64_t n1 = G.single_before[i];
        uint8_t* crow = codes.data() + n1 * w1 + (static_cast<uint64_t>(i - 1) - n1) * w2;
        if (single) crow[j - 1] = static_cast<uint8_t>(cd);
        else reinterpret_cast<uint16_t*>(crow)[j - 1] = cd;
        P[i * W + j] = pack_cell(H, a.Fm, a.Om);
      }
      if ((G.flags[i] & kFlagSink) && cy.H > best && (E->prune == 0 || hi_col[i] >= static_cast<int32_t>(L))) {
        best = cy.H;
        best_row = i;
      }
    }
    if (E->prune > 0 && pass == 0 && (best_row == 0 || best <= kNegBand + (1 << 20))) best = INT32_MIN;  // no feasible path in the narrow band
  }
  std::vector<int32_t> rev(2 * (static_cast<uint64_t>(R) + L + 2));
  const int32_t n = traceback_walk(best_row, L, codes.data(), w1, w2, G.single_before.data(), G.col0code.data(), G.pred_off.data(),
                                   G.preds.data(), G.node_id.data(), s, rev.data(),
                                   static_cast<int32_t>(R + L + 2));
  for (int32_t k = n - 1; k >= 0; --k) {
    E->last.push_back(rev[2 * k]);
    E->last.push_back(rev[2 * k + 1]);
  }
}

// Dynamic exact pruning (emulation of the planned kernel scheme): a cell is RELEVANT when
// H + suffix bound >= lb.  Row i computes the chunk interval [min rlo_p, max rhi_p + 1 + ext]
// over its predecessors p (ext = provable maximum length of a relevant horizontal run), cells
// outside are minus infinity; afterwards its own relevant interval is recorded.  lb is a guess;
// when the result scores below the guess the pass is repeated with the found score.
static bool emu_align_dyn(Emu* E, const uint8_t* read, uint32_t L, int32_t lb, int32_t* score_out) {
  const RankedGraph& G = E->rg;
  const Scores s{E->sc.m, E->sc.n, E->sc.g, E->sc.e, E->sc.q, E->sc.c};
  const uint32_t R = G.R;
  const uint64_t W = L + 1;
  const int C = 8;
  PredLutEntry lut[32];
  for (int low = 0; low < 32; ++low) lut[low] = pred_lut_entry(s, low);
  std::vector<int32_t> P((R + 1) * W, pack_cell(kNegBand, kNeg, kNeg));
  const uint32_t w1 = (L + 7 + 15) / 16 * 16, w2 = (L + 7 + 7) / 8 * 8 * 2;
  const uint64_t n1_total = G.single_before[R + 1];
  static std::vector<uint8_t> codes;
  codes.assign(n1_total * w1 + (R - n1_total) * w2 + 64, 0);
  P[0] = pack_cell(0, kNeg, kNeg);
  for (uint32_t j = 1; j <= L; ++j) P[j] = pack_cell(row0_h(s, j), kNeg, kNeg);
  const int nchunk = (L + C - 1) / C;
  std::vector<int> rlo(R + 1, nchunk), rhi(R + 1, -1);     // relevant chunk interval per row
  std::vector<int> clo(R + 1, nchunk), chi(R + 1, -1);     // computed chunk interval per row
  // row 0: everything is available (closed form)
  rlo[0] = 0; rhi[0] = nchunk - 1; clo[0] = 0; chi[0] = nchunk - 1;
  const bool have_lb = lb > INT32_MIN / 2;
  int ext = nchunk;
  if (have_lb) {
    const int64_t slack = static_cast<int64_t>(s.m) * L - lb;
    const int64_t dmax = slack <= 0 ? 0 : slack / (s.m - s.c) + 1;
    ext = static_cast<int>(std::min<int64_t>(nchunk, dmax / C + 2));
    if (E->dyn_ext >= 0) ext = std::min(ext, E->dyn_ext);
  }
  const bool lookahead = have_lb && E->dyn_ext >= 0;
  int32_t best = INT32_MIN; uint32_t best_row = 0;
  const int32_t* dp = G.depth.data();
  std::vector<uint8_t> rel0(R + 1, 0);   // column-0 cell of the row is relevant
  rel0[0] = 1;
  for (uint32_t i = 1; i <= R; ++i) {
    P[i * W] = pack_cell(G.h0[i], kNeg, kNeg);
    if (have_lb)
      for (uint32_t j = 1; j <= L; ++j)
        E->static_cells += cell_bound(s, dp[4 * i], dp[4 * i + 1], dp[4 * i + 2], dp[4 * i + 3], static_cast<int32_t>(j),
                                      static_cast<int32_t>(L)) >= lb;
    int lo = nchunk, hi = -1;
    for (uint32_t k = G.pred_off[i]; k < G.pred_off[i + 1]; ++k) {
      const uint32_t p = G.preds[k];
      if (rlo[p] <= rhi[p]) { lo = std::min(lo, rlo[p]); hi = std::max(hi, rhi[p] + 1); }
      // column 0 of the predecessor (leading graph nodes skipped) is not part of any chunk: when it
      // can still lie on a path scoring >= lb, the first chunk of this row has to be computed
      if (rel0[p]) { lo = 0; hi = std::max(hi, 0); }
    }
    rel0[i] = !have_lb || static_cast<int64_t>(G.h0[i]) + side_bound(s, dp[4 * i + 2], dp[4 * i + 3], static_cast<int32_t>(L)) >= lb;
    if (!have_lb) { lo = 0; hi = nchunk - 1; }
    else if (lo <= hi) hi = std::min(nchunk - 1, hi + ext);
    clo[i] = lo; chi[i] = hi;
    if (lo > hi) continue;
    RowCarry cy;
    if (lo == 0) cy = RowCarry{G.h0[i], kNeg, kNeg, G.h0[i]};
    else cy = RowCarry{kNegBand, kNeg, kNeg, kNegBand};
    for (int t = lo; t <= hi; ++t) {
      int32_t hmax = INT32_MIN;
      for (int c = 0; c < C; ++c) {
        const uint32_t j = 1 + t * C + c;
        if (j > L) break;
        CellAcc a;
        const int32_t sub = (G.letter[i] == read[j - 1]) ? s.m : s.n;
        const bool single = (G.pred_off[i + 1] - G.pred_off[i] == 1);
        auto pw = [&](uint32_t p, uint32_t jj) -> int32_t {   // predecessor cell, minus infinity outside its computed interval
          if (p == 0 || jj == 0) return P[p * W + jj];
          const int tc = (static_cast<int>(jj) - 1) / C;
          return (tc >= clo[p] && tc <= chi[p]) ? P[p * W + jj] : pack_cell(kNegBand, kNeg, kNeg);
        };
        int32_t H; uint16_t cd;
        if (single) {
          const uint32_t p = G.preds[G.pred_off[i]];
          const int32_t w = pw(p, j);
          const PredLutEntry& t = lut[w & 31];
          cell_pred_single(a, w, unpack_h(pw(p, j - 1)), sub, t.sf, t.so, t.sm);
          cd = static_cast<uint16_t>(cell_finish_single(a, cy, s, H));
        } else {
          cell_key_init(a);
          for (uint32_t k = G.pred_off[i]; k < G.pred_off[i + 1]; ++k) {
            const uint32_t p = G.preds[k];
            const int32_t w = pw(p, j);
            const PredLutEntry& t = lut[w & 31];
            cell_pred_key(a, 31 - static_cast<int32_t>(k - G.pred_off[i]), w, pw(p, j - 1), t.tf, t.to, t.tv);
          }
          cell_key_add_sub(a, sub);
          int32_t Fo, Oo;
          cd = cell_finish_key(a, cy, s, H, Fo, Oo);
          a.Fm = Fo; a.Om = Oo;
        }
        const uint64_t n1 = G.single_before[i];
        uint8_t* crow = codes.data() + n1 * w1 + (static_cast<uint64_t>(i - 1) - n1) * w2;
        if (single) crow[j - 1] = static_cast<uint8_t>(cd);
        else reinterpret_cast<uint16_t*>(crow)[j - 1] = cd;
        P[i * W + j] = pack_cell(H, a.Fm, a.Om);
        hmax = std::max(hmax, H);
        if (j == L && (G.flags[i] & kFlagSink) && H > best) { best = H; best_row = i; }
      }
      // chunk relevance: best H of the chunk + the largest suffix bound over its columns
      bool rel = !have_lb;
      if (have_lb) {
        const int32_t ja = 1 + t * C, jb = std::min<int32_t>(L, ja + C - 1);
        int32_t ub = INT32_MIN;
        for (int32_t j = ja; j <= jb; ++j) ub = std::max(ub, side_bound(s, dp[4 * i + 2], dp[4 * i + 3], static_cast<int32_t>(L) - j));
        rel = static_cast<int64_t>(hmax) + ub >= lb;
      }
      if (rel) { rlo[i] = std::min(rlo[i], t); rhi[i] = std::max(rhi[i], t); }
      if (rel && lookahead && t == hi && hi < nchunk - 1) E->overflow = true;
      E->kept_cells += C;
    }
    E->all_cells += L;
  }
  *score_out = best;
  if (best_row == 0) return false;
  std::vector<int32_t> rev(2 * (static_cast<uint64_t>(R) + L + 2));
  const int32_t n = traceback_walk(best_row, L, codes.data(), w1, w2, G.single_before.data(), G.col0code.data(), G.pred_off.data(),
                                   G.preds.data(), G.node_id.data(), s, rev.data(),
                                   static_cast<int32_t>(R + L + 2));
  E->last.clear();
  for (int32_t k = n - 1; k >= 0; --k) {
    E->last.push_back(rev[2 * k]);
    E->last.push_back(rev[2 * k + 1]);
  }
  return true;
}

extern "C" {
void* emu_new(int ring_rows) {
  Emu* e = new Emu();
  e->ring_rows = ring_rows;
  return e;
}
void emu_set_warp(void* h, int threads, int prune) { static_cast<Emu*>(h)->warp_threads = threads; static_cast<Emu*>(h)->warp_prune = prune; }
int emu_warp_retries(void* h) { return static_cast<Emu*>(h)->warp_retries; }
void emu_set_prune(void* h, int half_width) { static_cast<Emu*>(h)->prune = half_width; }
void emu_set_dyn_ext(void* h, int chunks) { static_cast<Emu*>(h)->dyn_ext = chunks; }
double emu_static_fraction(void* h) { Emu* e = static_cast<Emu*>(h); return e->all_cells > 0 ? e->static_cells / e->all_cells : 1.0; }
int emu_overflows(void* h) { return static_cast<Emu*>(h)->overflows; }
void emu_set_dyn(void* h, double lb_ratio) { static_cast<Emu*>(h)->dyn = 1; static_cast<Emu*>(h)->lb_ratio = lb_ratio; }
int emu_retries(void* h) { return static_cast<Emu*>(h)->retries; }
double emu_kept_fraction(void* h) { Emu* e = static_cast<Emu*>(h); return e->all_cells > 0 ? e->kept_cells / e->all_cells : 1.0; }
void emu_free(void* h) { delete static_cast<Emu*>(h); }
int64_t emu_add(void* h, const uint8_t* seq, int64_t len) {
  Emu* e = static_cast<Emu*>(h);
  emu_align(e, seq, static_cast<uint32_t>(len));
  const size_t n = e->last.size() / 2;
  std::vector<int32_t> nodes(n), pos(n);
  for (size_t k = 0; k < n; ++k) { nodes[k] = e->last[2 * k]; pos[k] = e->last[2 * k + 1]; }
  e->graph.add_alignment(nodes.data(), pos.data(), n, seq, static_cast<uint32_t>(len));
  return static_cast<int64_t>(n);
}
int64_t emu_last_alignment(void* h, int32_t* nodes, int32_t* pos, int64_t cap) {
  Emu* e = static_cast<Emu*>(h);
  const int64_t n = static_cast<int64_t>(e->last.size() / 2);
  for (int64_t k = 0; k < n && k < cap; ++k) { nodes[k] = e->last[2 * k]; pos[k] = e->last[2 * k + 1]; }
  return n;
}
// merge a given alignment without running the DP (host-graph micro-benchmarks)
void emu_add_pairs(void* h, const int32_t* nodes, const int32_t* pos, int64_t n, const uint8_t* seq, int64_t len) {
  static_cast<Emu*>(h)->graph.add_alignment(nodes, pos, static_cast<size_t>(n), seq, static_cast<uint32_t>(len));
}
double emu_time_export(void* h, int reps) {
  Emu* e = static_cast<Emu*>(h);
  const auto t0 = std::chrono::steady_clock::now();
  for (int k = 0; k < reps; ++k) e->graph.export_ranked(e->sc, 12, &e->rg);
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / reps;
}
int64_t emu_num_nodes(void* h) { return static_cast<Emu*>(h)->graph.num_nodes(); }
void emu_rank_to_node(void* h, int32_t* out) {
  const auto& r = static_cast<Emu*>(h)->graph.rank_to_node();
  for (size_t k = 0; k < r.size(); ++k) out[k] = static_cast<int32_t>(r[k]);
}
int64_t emu_consensus(void* h, char* out, int64_t cap) {
  std::string c = static_cast<Emu*>(h)->graph.consensus();
  std::memcpy(out, c.data(), std::min<size_t>(cap, c.size()));
  return static_cast<int64_t>(c.size());
}
void emu_msa_dims(void* h, int64_t* rows, int64_t* cols) {
  auto m = static_cast<Emu*>(h)->graph.msa();
  *rows = static_cast<int64_t>(m.size());
  *cols = m.empty() ? 0 : static_cast<int64_t>(m[0].size());
}
void emu_msa(void* h, char* out) {
  auto m = static_cast<Emu*>(h)->graph.msa();
  size_t off = 0;
  for (auto& r : m) { std::memcpy(out + off, r.data(), r.size()); off += r.size(); }
}

// Self-check of the shared cell header: the predecessor table against the definitions it tabulates, the
// key fold against a direct "first in-edge attaining the maximum", and the code fields against their decoders.
// Returns the number of mismatches (0 = fine); *checked = number of comparisons made.
int64_t emu_cell_selfcheck(uint32_t seed, int64_t n_random, int64_t* checked) {
  int64_t bad = 0, cnt = 0;
  auto rnd = [&seed]() { seed = seed * 1664525u + 1013904223u; return seed >> 8; };
  for (int g = -10; g <= -2; ++g) for (int e = g + 1; e <= std::min(g + 2, -1); ++e)
    for (int c = e + 1; c <= -1; ++c) for (int q = std::max(-10, c - 6); q < g; ++q) {
      const Scores s{5, -4, g, e, q, c};
      PredLutEntry lut[32];
      for (int low = 0; low < 32; ++low) {
        lut[low] = pred_lut_entry(s, low);
        const int dF = (low >> 3) & 3, dO = low & 7;
        const int Fc = std::max(g, e - dF), Oc = std::max(q, c - dO), V = std::max(Fc, Oc);
        const int xv = ((e - dF == V) || (g != V && c - dO == V)) ? 1 : 0;
        bad += lut[low].sf != Fc || lut[low].so != Oc || lut[low].sm != (xv | ((g >= e - dF) ? 2 : 0));
        bad += lut[low].tf != Fc * 64 + ((g >= e - dF) ? 1 : 0) - 2 * low;
        bad += lut[low].to != Oc * 64 + ((q >= c - dO) ? 1 : 0) - 2 * low;
        bad += lut[low].tv != V * 64 + xv - 2 * low;
        cnt += 4;
      }
      for (int64_t it = 0; it < n_random; ++it) {   // random in-edge sets through the key fold
        const int d = 1 + static_cast<int>(rnd() % 6);
        const int32_t sub = (rnd() & 1) ? s.m : s.n;
        int32_t w[6], wl[6];
        for (int k = 0; k < d; ++k) {
          const int32_t base = static_cast<int32_t>(rnd() % 64) - 32;     // close values: ties are frequent
          w[k] = pack_cell(base, base - static_cast<int32_t>(rnd() % 5), base - static_cast<int32_t>(rnd() % 9));
          wl[k] = pack_cell(static_cast<int32_t>(rnd() % 16) - 8, kNeg, kNeg);
        }
        CellAcc a; cell_key_init(a);
        for (int k = 0; k < d; ++k) {
          const PredLutEntry& t = lut[w[k] & 31];
          cell_pred_key(a, 31 - k, w[k], wl[k], t.tf, t.to, t.tv);
        }
        cell_key_add_sub(a, sub);
        // direct evaluation
        int32_t Fm = INT32_MIN, Om = INT32_MIN, Dm = INT32_MIN, Vm = INT32_MIN; int iF = 0, iO = 0, iD = 0, iV = 0;
        int fF = 0, fO = 0, fV = 0;
        for (int k = 0; k < d; ++k) {
          int32_t H, F, O; unpack_cell(w[k], H, F, O);
          const int32_t G = H + g, Fe = F + e, Oq = H + q, Oe = O + c;
          const int32_t Fc = std::max(G, Fe), Oc = std::max(Oq, Oe), V = std::max(Fc, Oc), D = unpack_h(wl[k]) + sub;
          if (Fc > Fm) { Fm = Fc; iF = k; fF = G >= Fe; }
          if (Oc > Om) { Om = Oc; iO = k; fO = Oq >= Oe; }
          if (D > Dm) { Dm = D; iD = k; }
          if (V > Vm) { Vm = V; iV = k; fV = (Fe == V) || (G != V && Oe == V); }
        }
        bad += (a.Fm >> 6) != Fm || 31 - ((a.Fm >> 1) & 31) != iF || (a.Fm & 1) != fF;
        bad += (a.Om >> 6) != Om || 31 - ((a.Om >> 1) & 31) != iO || (a.Om & 1) != fO;
        bad += (a.D >> 5) != Dm || 31 - (a.D & 31) != iD;
        const int32_t kV = static_cast<int32_t>(a.meta);
        bad += (kV >> 6) != Vm || 31 - ((kV >> 1) & 31) != iV || (kV & 1) != fV;
        cnt += 4;
      }
    }
  for (uint32_t move = 0; move < 3; ++move) for (uint32_t bits = 0; bits < 8; ++bits)
    for (uint32_t km = 0; km < 32; ++km) for (uint32_t ku = 0; ku < 32; ++ku) {
      const uint32_t cd = make_code(move, bits & 1, (bits >> 1) & 1, bits >> 2, km, ku);
      bad += (cd & 3) != move || ((cd >> 2) & 1) != (bits & 1) || ((cd >> 3) & 1) != ((bits >> 1) & 1) ||
             code_stop(cd) != (bits >> 2) || code_kmove(cd) != km || code_kup(cd) != ku;
      ++cnt;
    }
  for (uint32_t b = 0; b < 32; ++b) {   // byte codes of single-predecessor rows
    const uint32_t cd = code_of_single_byte(b);
    bad += (cd & 15) != (b & 15) || code_stop(cd) != ((b >> 4) & 1) || code_kmove(cd) != 0 || code_kup(cd) != 0;
    ++cnt;
  }
  if (checked) *checked = cnt;
  return bad;
}
}
