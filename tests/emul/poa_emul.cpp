// TEST INFRASTRUCTURE: sequential CPU emulation of the GPU alignment path.  It runs the
// product's host graph (poa_graph.cpp) and the product's cell arithmetic / traceback walker
// (poa_cell.h) with the same information loss as the kernels (predecessor rows are only
// visible as packed words; E opened from A), so that the algorithmic equivalence with the
// oracle's five-matrix equality traceback can be checked without a GPU.  Not shipped.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../svscope_b200/csrc/poa_cell.h"
#include "../../svscope_b200/csrc/poa_graph.h"

using namespace svs;

struct Emu {
  PoaGraph graph;
  PoaScoring sc;
  uint32_t ring_rows = 4;
  std::vector<int32_t> last;  // forward pairs
  RankedGraph rg;
};

static void emu_align(Emu* E, const uint8_t* read, uint32_t L) {
  E->last.clear();
  if (E->graph.empty() || L == 0) return;
  E->graph.export_ranked(E->sc, E->ring_rows, &E->rg);
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
  const SingleTables tabs = make_single_tables(s);
  int32_t best = INT32_MIN;
  uint32_t best_row = 0;
  for (uint32_t i = 1; i <= R; ++i) {
    P[i * W] = pack_cell(G.h0[i], kNeg, kNeg);
    RowCarry cy{G.h0[i], kNeg, kNeg, G.h0[i]};
    for (uint32_t j = 1; j <= L; ++j) {
      CellAcc a;
      const int32_t sub = (G.letter[i] == read[j - 1]) ? s.m : s.n;
      const bool single = (G.pred_off[i + 1] - G.pred_off[i] == 1);
      int32_t H;
      uint16_t cd;
      if (single) {  // the kernels' fast path
        const uint32_t p = G.preds[G.pred_off[i]];
        cell_pred_single(a, P[p * W + j], unpack_h(P[p * W + j - 1]), sub, s, tabs);
        cd = static_cast<uint16_t>(cell_finish_single(a, cy, s, H));
      } else {
        for (uint32_t k = G.pred_off[i]; k < G.pred_off[i + 1]; ++k) {
          const uint32_t p = G.preds[k];
          cell_pred_key(a, k - G.pred_off[i], P[p * W + j], unpack_h(P[p * W + j - 1]), sub, s, tabs);
        }
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
    if ((G.flags[i] & kFlagSink) && cy.H > best) {
      best = cy.H;
      best_row = i;
    }
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

extern "C" {
void* emu_new(int ring_rows) {
  Emu* e = new Emu();
  e->ring_rows = ring_rows;
  return e;
}
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
}
