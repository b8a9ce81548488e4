// TEST INFRASTRUCTURE: the product's warp-pipelined DP (svscope_b200/csrc/poa_dp2.cuh: bands, exact-size
// code rows, the row loop with its hand-over between warps, strips, exported rows, the warp traceback)
// compiled for the CPU through tests/emul/cuda_shim.h (one OS thread per warp, one user-level context per lane).  The
// graph comes from the test-side host graph (poa_graph.cpp); the driver below restates the few lines of
// poa_kernels.cu around the DP (task set-up, pruning attempts).  Not shipped.
#include "cuda_shim.h"

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <unistd.h>

#include "../../svscope_b200/csrc/poa_cell.h"
#include "../../svscope_b200/csrc/poa_task.h"
#include "poa_graph.h"

namespace svs {
namespace {

using std::max;
using std::min;

constexpr int32_t kSrcGlobal = 1 << 30;

// the one CtaExec service compute_bands2 needs: inclusive prefix sum by the whole CTA
struct CtaExec {
  void scan(uint32_t* a, uint32_t n) {
    __syncthreads();
    if (threadIdx.x == 0) for (uint32_t i = 1; i < n; ++i) a[i] += a[i - 1];
    __syncthreads();
  }
};

#include "../../svscope_b200/csrc/poa_dp2.cuh"

template <class T> T* aligned(std::vector<uint8_t>& store, size_t count) {
  store.assign(count * sizeof(T) + 256, 0);
  uintptr_t p = reinterpret_cast<uintptr_t>(store.data());
  p = (p + 255) / 256 * 256;
  return reinterpret_cast<T*>(p);
}

template <int T>
int run(const RankedGraph& G, const Scores& s, const uint8_t* read, uint32_t L, int ring_rows, bool prune, int32_t lb_guess,
        std::vector<int32_t>* rev_pairs, int32_t* score, int* retries) {
  constexpr int kC = 8;
  const uint32_t R = G.R;
  std::vector<uint8_t> st[12];
  PoaTask tk{};
  int32_t* depth = aligned<int32_t>(st[0], 4ull * (R + 2));
  std::memcpy(depth, G.depth.data(), G.depth.size() * 4);
  tk.letter = G.letter.data(); tk.pred_off = G.pred_off.data(); tk.preds = G.preds.data(); tk.flags = G.flags.data();
  tk.xslot = G.xslot.data(); tk.h0 = G.h0.data(); tk.col0code = G.col0code.data(); tk.node_id = G.node_id.data();
  tk.single_before = G.single_before.data(); tk.depth = depth; tk.read = read;
  tk.R = R; tk.L = L;
  const uint32_t cpp = T * kC;
  tk.npass = (L + cpp - 1) / cpp;
  tk.strip = ((L + tk.npass - 1) / tk.npass + kC - 1) / kC * kC;
  tk.ldx = (static_cast<uint64_t>(L) + 3 + kC + 7) / 8 * 8;
  tk.xrows = aligned<int32_t>(st[1], static_cast<size_t>(G.n_export + 1) * tk.ldx);
  tk.codes_cap = 2ull * (static_cast<uint64_t>(R) + 2) * (L + 16) + 4096;
  tk.codes = aligned<uint8_t>(st[2], tk.codes_cap);
  tk.bnd = aligned<int32_t>(st[3], 8ull * (R + 2));
  tk.result = aligned<int32_t>(st[4], 16);
  tk.path_cap = R + L + 2;
  tk.path = aligned<int32_t>(st[5], 2ull * tk.path_cap);
  tk.coff = aligned<uint32_t>(st[6], R + 4);
  int32_t* band = aligned<int32_t>(st[7], 2ull * (R + 2));
  TbRow* tbrows = aligned<TbRow>(st[8], R + 2);
  unsigned char* smem = aligned<unsigned char>(st[9], dp2_smem_bytes(T, ring_rows) + 64);
  unsigned long long need_bytes = 0, eval[5] = {0, 0, 0, 0, 0};
  tk.prune = prune ? 1u : 0u;
  tk.lb_guess = lb_guess;
  int n_retries = 0, rc = 0;
  CtaExec x;
  // debugging aid: SVS_EMU_WATCHDOG=<seconds> prints the progress words of every warp if the CTA has not finished by then
  std::atomic<bool> cta_done{false};
  std::thread watchdog;
  if (const char* wd = getenv("SVS_EMU_WATCHDOG")) {
    const double limit = atof(wd);
    watchdog = std::thread([&, limit]() {
      for (double t = 0; t < limit && !cta_done.load(); t += 0.05) std::this_thread::sleep_for(std::chrono::milliseconds(50));
      if (cta_done.load()) return;
      const int NW = T / 32;
      const volatile int* prog = reinterpret_cast<const volatile int*>(smem + static_cast<size_t>(T) * 32 * (ring_rows + 2) +
                                                                       static_cast<size_t>(NW) * kCarryDepth * sizeof(Carry));
      fprintf(stderr, "dp2 watchdog: R %u L %u npass %u strip %u ring %d prune %u\n", tk.R, tk.L, tk.npass, tk.strip, ring_rows, tk.prune);
      for (int w = 0; w < NW; ++w) fprintf(stderr, "  warp %d prog %d fprog %d\n", w, prog[w], prog[16 + w]);
      for (int w = 0; w < NW; ++w) {
        const uint32_t i = static_cast<uint32_t>(prog[w]) % (tk.R + 1);
        for (uint32_t r = (i > 4 ? i - 4 : 1); r <= std::min(tk.R, i + 40); ++r) {
          fprintf(stderr, "  row %u band [%d,%d] chunks [%d,%d] preds", r, band[2 * r], band[2 * r + 1], (band[2 * r] - 1) >> 3, (band[2 * r + 1] - 1) >> 3);
          for (uint32_t e = tk.pred_off[r]; e < tk.pred_off[r + 1]; ++e) fprintf(stderr, " %u", tk.preds[e]);
          fprintf(stderr, "%s\n", (tk.flags[r] & kFlagExport) ? " export" : "");
        }
        fprintf(stderr, "  --\n");
      }
      fflush(stderr);
      _exit(4);
    });
  }
  shim_run_cta(T, [&](int tid) {
    // poa_kernels.cu: bands (pruned: from the guessed lower bound; else full rows), exact-size code rows;
    // a result below the guess repeats the alignment with the score found
    int32_t lb = tk.lb_guess;
    bool have_lb = tk.prune != 0;
    bool overflow = false;
    for (int attempt = 0; attempt < 3; ++attempt) {
      if (tid == 0) { tk.result[0] = 0; tk.result[1] = INT32_MIN; }
      __syncthreads();
      compute_bands2<T>(x, tk, s, lb, have_lb, band, tk.coff, tbrows, &need_bytes);
      if (need_bytes > tk.codes_cap) { overflow = true; break; }
      dp2_align<T>(tk, s, ring_rows, smem, band, tk.coff, eval);
      const int32_t found_row = tk.result[0], found = tk.result[1];
      if (!have_lb || (found_row > 0 && found >= lb)) break;
      if (tid == 0) ++n_retries;
      have_lb = found_row > 0 && found > kNegBand / 2;
      lb = found;
      __syncthreads();
    }
    if (overflow) { if (tid == 0) rc = -2; return; }
    if (tid < 32) tb3_walk_warp(tk, s, tbrows);
    __syncthreads();
  });
  cta_done.store(true);
  if (watchdog.joinable()) watchdog.join();
  if (rc != 0) return rc;
  if (tk.result[2] < 0 || tk.result[0] <= 0) return -1;
  rev_pairs->assign(tk.path, tk.path + 2 * static_cast<size_t>(tk.result[2]));
  *score = tk.result[1];
  *retries = n_retries;
  return 0;
}

}  // namespace

// Aligns `read` to the rank-ordered graph with the product's DP source on T OS threads.  Pairs come back in
// REVERSE order (node id | -1, read position | -1).  Returns 0, -1 (no alignment) or -2 (code rows overflow).
int dp2_threads_align(const RankedGraph& G, const PoaScoring& sc, const uint8_t* read, uint32_t L, int threads, int ring_rows,
                      bool prune, int32_t lb_guess, std::vector<int32_t>* rev_pairs, int32_t* score, int* retries) {
  const Scores s{sc.m, sc.n, sc.g, sc.e, sc.q, sc.c};
  if (threads == 128) return run<128>(G, s, read, L, ring_rows, prune, lb_guess, rev_pairs, score, retries);
  if (threads == 256) return run<256>(G, s, read, L, ring_rows, prune, lb_guess, rev_pairs, score, retries);
  if (threads == 384) return run<384>(G, s, read, L, ring_rows, prune, lb_guess, rev_pairs, score, retries);
  if (threads == 512) return run<512>(G, s, read, L, ring_rows, prune, lb_guess, rev_pairs, score, retries);
  return -3;
}

}  // namespace svs
