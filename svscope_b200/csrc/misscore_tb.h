// Traceback of the MisScore alignment (svs_misscore_pairs), shared by the CUDA kernel and the
// CPU emulation of the tests (tests/emul/misscore_emul.cpp).
//
// Contract: the FIRST alignment returned by Bio.pairwise2.align.globalms(a, b, match, mismatch,
// g, g) as the reference's AligmentScore uses it (src/PairwiseCompare.py:19-30, called with
// 1, 0, -1, -1).  pairwise2 walks its trace matrix with a stack: at every cell the choices are
// tried in the order  gap in seqA (bit 1, column-1), match/mismatch (2), gap in seqB (4,
// row-1), extended gaps (8, 16);  a gap in seqA directly after (on the way back) a gap in
// seqB is a dead end, after which the most recent untried choice is resumed.
//
// With equal open and extend penalties (g <= 0) three bits per cell describe the DP:
//   t1: S[r][c-1] + g == S[r][c]   t2: S[r-1][c-1] + s(a,b) == S[r][c]   t4: S[r-1][c] + g == S[r][c]
// (the best alignment ending in a gap is always reached by "opening" it from the neighbour's
// best score, so the extend bits never add a first choice), and the first alignment is the
// plain greedy walk 1 > 2 > 4 without any backtracking:  if the walk takes t4 at (r,c), then
// t1(r,c) is clear, and  t1(r-1,c) and t4(r,c)  would give  S[r][c] = S[r-1][c-1] + 2g <=
// S[r][c-1] + g <= S[r][c], i.e. t1(r,c) - so the cell above never offers the forbidden gap in
// seqA; the same argument on row 1 shows the walk cannot reach the top border through t4
// with columns left.  The walk reports kMisDeadEnd if it ever met the rule (it cannot).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define SVS_HD __host__ __device__ __forceinline__
#else
#define SVS_HD inline
#endif

namespace svs {

// Trace nibbles of DP rows 1..la, two cells per byte, `pitch` bytes per row.
struct MisTrace {
  const uint8_t* bits;
  int64_t pitch;
  SVS_HD int nib(int row, int col) const {
    const uint8_t v = bits[static_cast<int64_t>(row - 1) * pitch + ((col - 1) >> 1)];
    return (v >> (((col - 1) & 1) * 4)) & 7;
  }
};

enum { kMisOk = 0, kMisDeadEnd = 1, kMisBadTrace = 2 };

// result: [0] alignment columns, [1] columns with equal characters ('|' of format_alignment,
// which compares the gapped strings: a literal '-' against a gap counts as equal).
// line (optional): match line of the alignment, written back to front into line[0..len).
SVS_HD int misscore_traceback(const MisTrace& T, const uint8_t* a, int la, const uint8_t* b, int lb,
                              uint8_t* line, int32_t* result) {
  int row = la, col = lb, len = 0, match = 0;
  bool col_gap = false;
  while (row > 0 && col > 0) {
    const int t = T.nib(row, col);
    if (t & 1) {
      if (col_gap) return kMisDeadEnd;
      col -= 1;
      const bool dash = b[col] == '-';  // a literal '-' equals the gap character in format_alignment
      if (line) line[len] = dash ? '|' : ' ';
      match += dash;
    } else if (t & 2) {
      row -= 1;
      col -= 1;
      const uint8_t x = a[row], y = b[col];
      if (line) line[len] = x == y ? '|' : ((x == '-' || y == '-') ? ' ' : '.');
      match += x == y;
      col_gap = false;
    } else if (t & 4) {
      row -= 1;
      const bool dash = a[row] == '-';
      if (line) line[len] = dash ? '|' : ' ';
      match += dash;
      col_gap = true;
    } else {
      return kMisBadTrace;
    }
    len += 1;
  }
  if (col && col_gap) return kMisDeadEnd;
  // _finish_backtrace: the rest of the longer sequence against gaps
  while (row > 0 || col > 0) {
    const bool dash = (row > 0 ? a[--row] : b[--col]) == '-';
    if (line) line[len] = dash ? '|' : ' ';
    match += dash;
    len += 1;
  }
  result[0] = len;
  result[1] = match;
  return kMisOk;
}

}  // namespace svs
