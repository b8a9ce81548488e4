// Cell-level arithmetic of the sequence-to-graph alignment, shared by the CUDA kernels
// (poa_kernels.cu) and by host code (column-0 codes in poa_graph.cpp, the CPU emulation used
// by tests/).  Everything here is integer and exact.
//
// Recurrences (global alignment, two-piece "convex" gaps; behaviour of spoa's scalar engine,
// SURVEY.md Appendix B) for row i (graph node) with predecessor rows p_0..p_{d-1} in stored
// in-edge order, column j (read position j-1), substitution score s:
//   F[i][j] = max_k max(H[p_k][j]+g, F[p_k][j]+e)      O[i][j] = max_k max(H[p_k][j]+q, O[p_k][j]+c)
//   D[i][j] = max_k H[p_k][j-1] + s                    A[i][j] = max(D, F, O)
//   E[i][j] = max(A[i][j-1]+g, E[i][j-1]+e)            Q[i][j] = max(A[i][j-1]+q, Q[i][j-1]+c)
//   H[i][j] = max(A, E, Q)
// E is opened from A instead of H (spoa: E[j] = max(H[j-1]+g, E[j-1]+e)).  Why this is exact, for
// every convex parameter set (q < g < e < c, the only mode supported):
//  * Q: a Q gap opened on an E-derived H is dominated by the single Q gap from the same origin
//    (g+(d-1)e+q+(m-1)c < q+(d+m-1)c since g<c, e<c), one opened on a Q-derived H by extending
//    it (q<c); so Q computed from A equals spoa's Q.
//  * E: spoa's E can exceed the A-opened E only through "Q run of length d, then E opened on
//    it"; then the Q run simply continued scores higher by m(c-e)+(e-g) > 0 at every later
//    column, so H = max(A, E, Q) is the same, and in the traceback tests that involve E
//    (H == E[j-1]+e;  E[j]+e == E[j+1] || Q[j]+c == Q[j+1]) either the E term is false in both
//    versions or the Q term is true in both.
//  * E opened on an E-derived H is dominated by extending (g<e).
// This turns the horizontal dependency into two independent max-plus prefix scans.
// tests/test_host_logic.py runs this arithmetic against the five-matrix oracle.
//
// Instead of keeping five score matrices for an equality-test traceback, every cell emits a
// code holding exactly the decisions that traceback would take there (16 bits):
//   bits 0-1   move that defines H: 0 diagonal, 1 vertical, 2 horizontal
//   bit  2     "extend" flag of a vertical / horizontal move (gap continues)
//   bit  3     horizontal continuation: E[j]+e==E[j+1] || Q[j]+c==Q[j+1]
//   bits 5-9   31 - (in-edge index of the diagonal / vertical move)
//   bit  10    stop flag of the vertical-extension walk at this cell
//   bits 11-15 31 - (in-edge index of the vertical-extension walk); 0 = none: go to the source row
// (the indices are stored complemented because that is how the max-keys of the fold carry them).
// Rows with a single predecessor keep one byte per cell: bits 0-3 as above, bit 4 = stop flag,
// their in-edge indices being 0 (code_of_single_byte widens it).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define SVS_HD __host__ __device__ __forceinline__
#else
#define SVS_HD inline
#endif

namespace svs {

constexpr int32_t kNeg = -(1 << 28);  // "minus infinity": |scores| stay below 2^26
constexpr uint32_t kMoveDiag = 0, kMoveVert = 1, kMoveHorz = 2;
constexpr uint32_t kNoPred = 31;
constexpr uint32_t kMaxIndeg = 31;        // in-edge index must fit 5 bits, 31 is reserved
constexpr int64_t kMaxScoreSpan = 1 << 26;
constexpr uint8_t kFlagSink = 1, kFlagExport = 4;  // per-row flags of the ranked graph
constexpr uint8_t kFlagChain = 8;  // the row has exactly one predecessor and it is the previous row

struct Scores {
  int32_t m, n, g, e, q, c;
};

SVS_HD int32_t imax(int32_t a, int32_t b) { return a > b ? a : b; }
SVS_HD int32_t imin(int32_t a, int32_t b) { return a < b ? a : b; }

SVS_HD uint16_t make_code(uint32_t move, uint32_t ext, uint32_t lcnext, uint32_t upstop,
                          uint32_t k_move, uint32_t k_up) {
  return static_cast<uint16_t>(move | (ext << 2) | (lcnext << 3) | ((31u - k_move) << 5) | (upstop << 10) |
                               ((31u - k_up) << 11));
}
// the byte code of a single-predecessor row in the 16-bit layout (both in-edge indices 0)
SVS_HD uint32_t code_of_single_byte(uint32_t b) { return (b & 15u) | ((b & 16u) << 6) | (31u << 5) | (31u << 11); }
SVS_HD uint32_t code_kmove(uint32_t cd) { return 31u - ((cd >> 5) & 31u); }
SVS_HD uint32_t code_stop(uint32_t cd) { return (cd >> 10) & 1u; }
SVS_HD uint32_t code_kup(uint32_t cd) { return 31u - ((cd >> 11) & 31u); }

// One word per cell carries all a successor row needs: H, and F and O as clamped distances
// below H.  F matters to a successor only through max(H+g, F+e) and the equality tests
// against it, i.e. only while H-F <= e-g (<= 2); likewise O while H-O <= c-q (<= 6).
SVS_HD int32_t pack_cell(int32_t H, int32_t F, int32_t O) {
  const int32_t dF = imin(H - F, 3);
  const int32_t dO = imin(H - O, 7);
  return H * 32 + dF * 8 + dO;
}
SVS_HD void unpack_cell(int32_t w, int32_t& H, int32_t& F, int32_t& O) {
  H = w >> 5;
  F = H - ((w >> 3) & 3);
  O = H - (w & 7);
}
SVS_HD int32_t unpack_h(int32_t w) { return w >> 5; }

// Source row (row 0): H[0][0]=0, H[0][j]=max(g+(j-1)e, q+(j-1)c); F, O = -inf for j>=1.
SVS_HD int32_t row0_e(const Scores& s, int32_t j) { return j == 0 ? 0 : s.g + (j - 1) * s.e; }
SVS_HD int32_t row0_q(const Scores& s, int32_t j) { return j == 0 ? 0 : s.q + (j - 1) * s.c; }
SVS_HD int32_t row0_h(const Scores& s, int32_t j) {
  return j == 0 ? 0 : imax(row0_e(s, j), row0_q(s, j));
}

// Running state of one cell over its predecessor rows (four registers).  Single-predecessor
// rows: Fm, Om, D = the F, O and diagonal candidates, meta = flags.  Rows with several
// predecessors: the same fields hold max-keys (see cell_pred_key).
struct CellAcc {
  int32_t Fm, Om, D;
  uint32_t meta;
};

// State of the previous column of the same row.
struct RowCarry {
  int32_t A, E, Q, H;
};

// ---- what a successor reads from a predecessor's packed word -----------------------------------
// The vertical candidates depend on the predecessor only through H and the five low bits
// (low = dF*8+dO) of its packed word:
//   F = H + max(g, e - dF),   O = H + max(q, c - dO),
// and the flags taken at the predecessor (xV: a vertical move is an extension, sF / sO: the gap
// open attains the F / O maximum, i.e. the extension walk stops there) are pure functions of low.
// All of it comes from a 32-entry table indexed by low.  On the device lane l of a warp holds
// entry l in registers and a look-up is ONE warp shuffle whose source lane is the packed word
// itself (the shuffle uses its five low bits): no field extraction in the row loop.
struct PredLutEntry {
  // rows with several predecessors ("key" form, see below): key of in-edge k = 2*w + 2*(31-k) + t*
  int32_t tf, to, tv;
  // rows with one predecessor: F = H + sf, O = H + so, sm = xV | sF << 1
  int32_t sf, so, sm;
};

SVS_HD PredLutEntry pred_lut_entry(const Scores& s, int32_t low) {
  const int32_t dF = (low >> 3) & 3, dO = low & 7;
  const int32_t G = s.g, Fe = -dF + s.e, Oq = s.q, Oe = -dO + s.c;
  const int32_t Fc = imax(G, Fe), Oc = imax(Oq, Oe), V = imax(Fc, Oc);
  const int32_t xv = ((Fe == V) || (G != V && Oe == V)) ? 1 : 0;
  const int32_t sf = (G >= Fe) ? 1 : 0, so = (Oq >= Oe) ? 1 : 0;
  PredLutEntry t;
  t.tf = Fc * 64 + sf - 2 * low;
  t.to = Oc * 64 + so - 2 * low;
  t.tv = V * 64 + xv - 2 * low;
  t.sf = Fc;
  t.so = Oc;
  t.sm = xv | (sf << 1);
  return t;
}

// ---- fast path for rows with a single predecessor (most rows) --------------------------------
// w = packed word of the predecessor in this column, Hpl = its H in the column on the left,
// (sfv, sov, smv) = table entry of w.  Result: a.Fm, a.Om, a.D = the F, O and diagonal candidates,
// a.meta = xV | sF << 1.
SVS_HD void cell_pred_single(CellAcc& a, int32_t w, int32_t Hpl, int32_t sub, int32_t sfv, int32_t sov, int32_t smv) {
  const int32_t Hp = w >> 5;
  a.Fm = Hp + sfv;
  a.Om = Hp + sov;
  a.D = Hpl + sub;
  a.meta = static_cast<uint32_t>(smv);
}

// Completes a single-predecessor cell; returns the low byte of the code (indices are 0).
// Cells outside a pruning band carry kNegBand and whatever is derived from them stays below every
// real score (|real| < 2^21 while pruning is on, kNegBand = -2^22, drift per row / column < 16):
// no clamping is needed, decisions of cells on a co-optimal path never involve such values.
SVS_HD uint32_t cell_finish_single(const CellAcc& a, RowCarry& cy, const Scores& s, int32_t& H_out) {
  const int32_t eo = cy.A + s.g, ee = cy.E + s.e, qo = cy.A + s.q, qe = cy.Q + s.c;
  const int32_t E = imax(eo, ee), Q = imax(qo, qe);
  const int32_t V = imax(a.Fm, a.Om);
  const int32_t A = imax(a.D, V);
  const int32_t H = imax(A, imax(E, Q));
  const bool is_d = (a.D == H), is_v = (V == H);
  const uint32_t hx = (ee == H) || (cy.H + s.g != H && qe == H);
  const uint32_t move = is_d ? kMoveDiag : (is_v ? kMoveVert : kMoveHorz);
  const uint32_t ext = is_d ? 0u : (is_v ? (a.meta & 1u) : hx);
  const uint32_t lcnext = (E + s.e >= A + s.g) || (Q + s.c >= A + s.q);
  cy.A = A; cy.E = E; cy.Q = Q; cy.H = H;
  H_out = H;
  return move | (ext << 2) | (lcnext << 3) | ((a.meta & 2u) << 3);
}

// ---- rows with several predecessors: "first in-edge attaining the maximum" by key ---------------
// A candidate value v of in-edge k with flag f is folded as the single integer
//   key = v*64 + (31-k)*2 + f ,
// so that one max() keeps the largest value, among equal values the smallest in-edge index,
// and carries the flag of that in-edge along (|v| < 2^24 is checked before the alignment starts).
// With v = H + (table value) and the packed word w = H*32 + low:  key = 2*w + 2*(31-k) + t[low].
// The diagonal candidates differ between in-edges only by H of the left column, so their key is
//   (wl & ~31) | (31-k) = H_left*32 + (31-k);  the substitution score is added once per cell
// after the fold (cell_key_add_sub).
// CellAcc is reused: Fm = key of F, Om = key of O, D = key of the diagonal, meta = key of V.
constexpr int64_t kMaxKeySpan = 1 << 24;

SVS_HD void cell_key_init(CellAcc& a) {
  a.Fm = INT32_MIN; a.Om = INT32_MIN; a.D = INT32_MIN; a.meta = static_cast<uint32_t>(INT32_MIN);
}

// rk = 31 - k;  w, wl = packed words of in-edge k in this column / the column on the left;
// (tfv, tov, tvv) = table entry of w
SVS_HD void cell_pred_key(CellAcc& a, int32_t rk, int32_t w, int32_t wl, int32_t tfv, int32_t tov, int32_t tvv) {
  const int32_t t = 2 * w + 2 * rk;
  a.Fm = imax(a.Fm, t + tfv);
  a.Om = imax(a.Om, t + tov);
  a.meta = static_cast<uint32_t>(imax(static_cast<int32_t>(a.meta), t + tvv));
  a.D = imax(a.D, (wl & ~31) | rk);
}

// after the last in-edge: the diagonal key becomes (H_left + sub)*32 + (31-k)
SVS_HD void cell_key_add_sub(CellAcc& a, int32_t sub) { a.D += sub * 32; }

// value of the cell state a successor needs (for packing / the scan)
SVS_HD int32_t key_value(int32_t key) { return key >> 6; }
SVS_HD int32_t key_value_diag(int32_t key) { return key >> 5; }

SVS_HD uint16_t cell_finish_key(const CellAcc& a, RowCarry& cy, const Scores& s, int32_t& H_out,
                                int32_t& F_out, int32_t& O_out) {
  const int32_t kV = static_cast<int32_t>(a.meta);
  const int32_t Fm = a.Fm >> 6, Om = a.Om >> 6, D = a.D >> 5, V = kV >> 6;
  const int32_t eo = cy.A + s.g, ee = cy.E + s.e, qo = cy.A + s.q, qe = cy.Q + s.c;
  const int32_t E = imax(eo, ee), Q = imax(qo, qe);
  const int32_t A = imax(D, V);
  const int32_t H = imax(A, imax(E, Q));
  const bool is_d = (D == H), is_v = (V == H);
  const uint32_t hx = (ee == H) || (cy.H + s.g != H && qe == H);
  const uint32_t move = is_d ? kMoveDiag : (is_v ? kMoveVert : kMoveHorz);
  const uint32_t ext = is_d ? 0u : (is_v ? static_cast<uint32_t>(kV & 1) : hx);
  // complemented in-edge index of the move, straight from the key (ignored for horizontal moves)
  const uint32_t rmove = (is_d ? static_cast<uint32_t>(a.D) : (static_cast<uint32_t>(kV) >> 1)) & 31u;
  const uint32_t lcnext = (E + s.e >= A + s.g) || (Q + s.c >= A + s.q);
  // vertical-extension walk: first in-edge attaining F or O (F wins ties).  The six low bits of a
  // key are (31-k) << 1 | stop flag, exactly bits 10-15 of the code.
  const uint32_t fF = static_cast<uint32_t>(a.Fm) & 63u, fO = static_cast<uint32_t>(a.Om) & 63u;
  const uint32_t up = ((fF | 1u) >= fO) ? fF : fO;
  cy.A = A; cy.E = E; cy.Q = Q; cy.H = H;
  H_out = H; F_out = Fm; O_out = Om;
  return static_cast<uint16_t>(move | (ext << 2) | (lcnext << 3) | (rmove << 5) | (up << 10));
}

// Traceback code of a source-row cell (0, j), j >= 1: always a horizontal move.
SVS_HD uint16_t row0_code(const Scores& s, int32_t j) {
  const int32_t H = row0_h(s, j);
  const int32_t El = row0_e(s, j - 1), Ql = row0_q(s, j - 1), Hl = row0_h(s, j - 1);
  const uint32_t ext = (El + s.e == H) || (Hl + s.g != H && Ql + s.c == H);
  const uint32_t lcnext = (row0_e(s, j) + s.e == row0_e(s, j + 1)) || (row0_q(s, j) + s.c == row0_q(s, j + 1));
  return make_code(kMoveHorz, ext, lcnext, 0, 0, 0);
}

// ---- exact pruning -----------------------------------------------------------------------------
// Upper bound of the score of ANY global alignment through cell (row, j): the prefix consumes j
// read characters and k1 graph nodes with dmin <= k1 <= dmax (nodes on a source->row path,
// row included), the suffix L-j characters and k2 nodes with smin <= k2 <= smax (row->sink,
// row excluded).  Each side scores at most m per aligned pair and at least |c| (the cheapest
// gap character, no open charged so that a gap may straddle the cell) per unpaired one.
// A cell with bound < score of some feasible alignment cannot lie on a co-optimal path, and
// lowering such cells to "minus infinity" only lowers non-maximal candidates of the cells
// that are, so every traceback decision is unchanged (DESIGN.md "Exact pruning").
constexpr int32_t kNegBand = -(1 << 22);     // pruned cells; real scores stay above -2^21

SVS_HD int32_t side_bound(const Scores& s, int32_t lo, int32_t hi, int32_t x) {
  const int32_t k = x < lo ? lo : (x > hi ? hi : x);
  const int32_t d = k > x ? k - x : x - k;
  return s.m * (k < x ? k : x) + s.c * d;
}

SVS_HD int32_t cell_bound(const Scores& s, int32_t dmin, int32_t dmax, int32_t smin, int32_t smax, int32_t j,
                          int32_t L) {
  return side_bound(s, dmin, dmax, j) + side_bound(s, smin, smax, L - j);
}

// Everything the traceback needs to read.
struct TbView {
  const uint8_t* codes;
  uint32_t w1, w2;            // row pitch in bytes: single-predecessor rows / others
  // band-limited code rows (pruned alignments): row r holds the cells from column
  // 1 + first_chunk(r)*cols on, first_chunk(r) = (band[2r]-1)/cols; nullptr = full-width rows
  const int32_t* band;
  uint32_t cols;
  const uint32_t* single_before;
  const uint16_t* col0code;
  const uint32_t* pred_off;
  const uint32_t* preds;
  const uint32_t* node_id;
};

SVS_HD uint32_t tb_code_at(const TbView& v, const Scores& s, uint32_t ii, uint32_t jj) {
  if (ii == 0) return jj == 0 ? 0u : row0_code(s, static_cast<int32_t>(jj));
  if (jj == 0) return v.col0code[ii];
  const uint64_t n1 = v.single_before[ii];
  const uint8_t* row = v.codes + n1 * v.w1 + (static_cast<uint64_t>(ii - 1) - n1) * v.w2;
  uint32_t col = jj - 1;
  if (v.band != nullptr) col -= (static_cast<uint32_t>(v.band[2 * ii] - 1) / v.cols) * v.cols;
  if (v.pred_off[ii + 1] - v.pred_off[ii] == 1) return code_of_single_byte(row[col]);   // 1-byte code: in-edge indices are 0
  return reinterpret_cast<const uint16_t*>(row)[col];
}

// One iteration of the traceback loop at (i, j) != (0, 0), including the gap-extension walks
// it may trigger.  Appends pairs (node id | -1, read position | -1) in REVERSE order.
// Returns false if `cap` pairs do not suffice.
SVS_HD bool tb_step(const TbView& v, const Scores& s, uint32_t& i, uint32_t& j, int32_t& n, int32_t* out_pairs,
                    int32_t cap) {
  auto pred_row = [&](uint32_t ii, uint32_t k) -> uint32_t {
    return k == kNoPred ? 0u : v.preds[v.pred_off[ii] + k];
  };
  const uint32_t cd = tb_code_at(v, s, i, j);
  const uint32_t move = cd & 3, ext = (cd >> 2) & 1, km = code_kmove(cd);
  uint32_t pi, pj;
  if (move == kMoveDiag) { pi = pred_row(i, km); pj = j - 1; }
  else if (move == kMoveVert) { pi = pred_row(i, km); pj = j; }
  else { pi = i; pj = j - 1; }
  if (n >= cap) return false;
  out_pairs[2 * n] = (i == pi) ? -1 : static_cast<int32_t>(v.node_id[i]);
  out_pairs[2 * n + 1] = (j == pj) ? -1 : static_cast<int32_t>(j - 1);
  ++n;
  i = pi; j = pj;
  if (move == kMoveHorz && ext) {
    while (true) {
      if (n >= cap) return false;
      out_pairs[2 * n] = -1;
      out_pairs[2 * n + 1] = static_cast<int32_t>(j - 1);
      ++n;
      --j;
      if (j == 0 || !((tb_code_at(v, s, i, j) >> 3) & 1)) break;
    }
  } else if (move == kMoveVert && ext) {
    while (i != 0) {
      const uint32_t c2 = tb_code_at(v, s, i, j);
      const uint32_t stop = code_stop(c2), ku = code_kup(c2);
      const uint32_t up = pred_row(i, ku);
      if (n >= cap) return false;
      out_pairs[2 * n] = static_cast<int32_t>(v.node_id[i]);
      out_pairs[2 * n + 1] = -1;
      ++n;
      i = up;
      if (stop || i == 0) break;
    }
  }
  return true;
}

// Walks the stored decisions from (best_row, L) back to (0, 0) and writes the alignment
// pairs in REVERSE order.  Returns the number of pairs, or -1 if `cap` pairs do not suffice.
SVS_HD int32_t traceback_walk(uint32_t best_row, uint32_t L, const uint8_t* codes, uint32_t w1, uint32_t w2,
                              const uint32_t* single_before, const uint16_t* col0code, const uint32_t* pred_off,
                              const uint32_t* preds, const uint32_t* node_id, const Scores& s,
                              int32_t* out_pairs, int32_t cap) {
  const TbView v{codes, w1, w2, nullptr, 8, single_before, col0code, pred_off, preds, node_id};
  uint32_t i = best_row, j = L;
  int32_t n = 0;
  while (!(i == 0 && j == 0)) {
    if (!tb_step(v, s, i, j, n, out_pairs, cap)) return -1;
  }
  return n;
}

}  // namespace svs
