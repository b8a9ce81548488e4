// Window kernel: one CTA aligns ALL sequences of a window (group), one after the other, with the
// partial-order graph resident in its scratch slot (poa_dgraph.h) - no host round trip between
// two reads of a window.  Host side: poa_window.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "poa_cell.h"
#include "poa_dgraph.h"

namespace svs {

struct WinDesc {            // one window (group of sequences), written by the host
  int64_t member_begin;     // first entry of the group in `members`
  WinCaps caps;             // slot layout of this window (nseq = number of members)
  int64_t pairs_off;        // debug: first pair slot of the window in `pairs_out`, or -1
};

struct WinResult {
  int32_t status;           // WinStatus
  uint32_t msa_rows, msa_cols, cons_len;
  uint64_t out_off;         // MSA (rows x cols, row-major, letters) at out_base + out_off, consensus right after it
  uint32_t n_align, retries;
  uint32_t nodes, edges;
  uint64_t cells;           // nominal DP cells: sum (R + 1)(L + 1)
  uint64_t eval_cells;      // cells the kernel evaluated (chunks of 8 inside the bands; dp_kernel 2 only)
  uint64_t rows, exported;  // DP rows / rows exported to global memory, summed over the alignments
  uint64_t need_bytes;      // kWinCodesCap: bytes of traceback codes the failing alignment needed
  uint64_t read_bases, path_steps, pred_entries;   // summed over the alignments (algorithmic bytes)
  uint64_t warp_cyc[4];     // DP, summed over warps: cycles in the row loop, polling the left warp, polling the
                            // right warp / strip boundary, waiting for the last warp at the end
  uint64_t cyc[8];          // SM cycles of thread 0 per phase: export, bands+DP, traceback, merge, rank order, finish
};

struct WinParams {
  const uint8_t* reads;       // all sequences, 1 B per base
  const int64_t* read_off;    // [n_reads + 1]
  const int64_t* members;     // sequence ids of all windows
  const WinDesc* desc;
  const int32_t* order;       // windows in launch order (largest first)
  int n_windows;
  int* counter;               // next entry of `order`
  uint8_t* slot_base;
  uint64_t slot_bytes;
  int* slot_flags;
  int n_slots;
  uint8_t* out_base;
  uint64_t out_cap;
  unsigned long long* out_cursor;
  WinResult* results;         // [n_windows], indexed by window
  int32_t* pairs_out;         // debug: forward alignment pairs of every sequence, or nullptr
  int64_t* pair_cnt;          // debug: [members] pairs per sequence
  Scores s;
  int ring_rows;
  int prune;
  int want_msa;
  float prune_margin;
  int sm_limit;               // profiling aid: > 0 = only CTAs that land on SMs below this id take work
  int dp_version;             // 2: warp-pipelined DP (8 columns per thread), 1: barrier-per-row DP
};

size_t poa_window_smem_bytes(int threads, int ring_rows, int cols);
int poa_window_ctas_per_sm(int threads, int ring_rows, int cols);   // 0 = configuration not available
cudaError_t poa_window_configure(int threads, int ring_rows, int cols);
cudaError_t poa_window_launch(const WinParams& p, int grid, int threads, int cols, cudaStream_t stream);

}  // namespace svs
