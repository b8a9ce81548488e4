// Device-visible descriptor of one (graph, read) alignment.  All pointers are device
// pointers into the batch arena; the layout of the rank-ordered graph arrays is the one
// produced by PoaGraph::export_ranked (poa_graph.h).
#pragma once
#include <cstdint>

namespace svs {

struct PoaTask {
  // graph in rank order: row r = rank r-1, row 0 = virtual source row
  const uint8_t* letter;      // [R+1]
  const uint32_t* pred_off;   // [R+2]
  const uint32_t* preds;      // predecessor rows in stored in-edge order (0 = source row)
  const uint8_t* flags;       // [R+1] kFlagSink | kFlagExport
  const int32_t* xslot;       // [R+1] slot of the row in `xrows`, or -1
  const int32_t* h0;          // [R+1] H[row][0]
  const uint16_t* col0code;   // [R+1] traceback codes of column 0
  const uint32_t* node_id;    // [R+1]
  const uint32_t* single_before;  // [R+2] number of single-predecessor rows among rows 1..i-1
  const int32_t* depth;       // [R+1][4] nodes on source->row paths (min, max; row included) and on
                              // row->sink paths (min, max; row excluded): bounds of the exact pruning
  const uint8_t* read;        // [L]
  uint32_t R, L;
  uint32_t strip, npass;      // columns per pass (multiple of 8), number of passes
  // scratch
  // traceback codes of columns 1..L: rows with one predecessor store 1 byte per cell (row
  // pitch w1), rows with several store 2 bytes per cell (row pitch w2); row i starts at byte
  // single_before[i]*w1 + (i-1-single_before[i])*w2
  uint8_t* codes;
  uint32_t w1, w2;
  int32_t* xrows;             // [n_export][ldx] packed cells of exported rows, column j at 3+j
  uint64_t ldx;
  int32_t* bnd;               // [2][4][R+1] strip boundary state (H, A, E, Q), ping-pong
  // results
  int32_t* result;            // [4] best_row, best_score, n_pairs, pruning retries
  int32_t* path;              // [2*path_cap] alignment pairs in reverse order
  uint32_t path_cap;
  uint32_t pad_;
  // persistent kernel: scratch offsets inside the per-SM slot (codes/xrows/bnd are patched)
  uint64_t off_codes, off_xrows, off_bnd, off_band;
  uint32_t prune;             // 1: prune with bands from `lb_guess` (retry inside the kernel if it was too high)
  int32_t lb_guess;           // guessed lower bound of the optimal score
  uint64_t codes_cap;         // bytes available for traceback codes (pruned alignments use band-limited rows)
  uint32_t* coff;             // [R+2] warp-pipelined kernel: offset of every row's codes, in units of 8 bytes
};

}  // namespace svs
