// Row update of the MisScore DP (linear gap: open == extend == -gap), shared by the CUDA kernel
// (misscore.cu) and the CPU emulation of the tests.  A thread owns kC consecutive columns of a
// row; the dependence on the left neighbour, S[c] = max(T[c], S[c-1] - gap), is a max-plus
// prefix scan: pass 1 scans the thread's own columns, the CTA combines the threads' last
// values (X_t = last_t + gap*kC*(t+1), exclusive prefix maximum, Lin_t = that - gap*kC*t), pass 2
// folds the incoming value in and emits the three trace bits per cell (misscore_tb.h).
#pragma once
#include <cstdint>

#include "misscore_tb.h"

namespace svs {

constexpr int kMisNeg = -(1 << 28);

// up[k] = S[r-1][j0+k], upleft = S[r-1][j0-1], eq bit k = (a[r-1] == b[j0+k-1]).
// p_init = S[r][j0-1] if already known (first thread of a strip), else kMisNeg.
template <int kC>
SVS_HD int mis_pass1(const int (&up)[kC], int upleft, uint32_t eq, int s_match, int s_mis, int gap, int p_init,
                     int (&loc)[kC]) {
  int diag = upleft;
  int p = p_init;
#pragma unroll
  for (int k = 0; k < kC; ++k) {
    const int d = diag + (((eq >> k) & 1u) ? s_match : s_mis);
    const int u = up[k] - gap;
    const int t = d > u ? d : u;
    p = p - gap;
    p = t > p ? t : p;
    loc[k] = p;
    diag = up[k];
  }
  return p;
}

// lin = S[r][j0-1].  Replaces up[] / upleft by row r and returns the kC trace nibbles
// (cell k in bits 4k..4k+2: t1 | t2 << 1 | t4 << 2).
template <int kC>
SVS_HD uint64_t mis_pass2(int (&up)[kC], int& upleft, uint32_t eq, int s_match, int s_mis, int gap, int lin,
                          const int (&loc)[kC]) {
  static_assert(kC <= 16, "one 64-bit word of nibbles per thread and row");
  int left = lin;
  int diag = upleft;
  int reach = lin;
  uint64_t nibs = 0;
#pragma unroll
  for (int k = 0; k < kC; ++k) {
    reach -= gap;
    const int s = loc[k] > reach ? loc[k] : reach;
    const int d = diag + (((eq >> k) & 1u) ? s_match : s_mis);
    const int u = up[k] - gap;
    const uint32_t nib = (left - gap == s ? 1u : 0u) | (d == s ? 2u : 0u) | (u == s ? 4u : 0u);
    nibs |= static_cast<uint64_t>(nib) << (4 * k);
    diag = up[k];
    up[k] = s;
    left = s;
  }
  upleft = lin;
  return nibs;
}

// Columns per strip: the fewest strips of at most threads*kC columns, evenly wide, a multiple of kC.
SVS_HD int mis_strip_cols(int lb, int threads, int kC) {
  const int cap = threads * kC;
  const int nstrips = (lb + cap - 1) / cap;
  const int per = (lb + nstrips - 1) / nstrips;
  return (per + kC - 1) / kC * kC;
}

// Bytes per row of trace nibbles (whole 64-bit words).
SVS_HD int64_t mis_trace_pitch(int lb) { return static_cast<int64_t>((lb + 15) / 16) * 8; }

}  // namespace svs
