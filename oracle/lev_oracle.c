/* ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Not used by the product path.
 *
 * PARITY UNPINNED: unit-cost Levenshtein distance, the semantic contract of
 * `Levenshtein.distance` (Levenshtein 0.23.0, README.md:23).  The reference never calls it
 * from live code: its only use is the commented-out read-by-read matrix in
 * src/DecisionMaker.py:76-84 (import commented at :34).  The module is not installed here
 * and the reference holds no vectors for it, so parity is defined against the textbook DP
 * below (`lev_dp`); `lev_myers` is the same function computed bit-parallel (Myers 1999 /
 * Hyyro 2003 block form) and exists only so that the CPU baseline is timed with the kind of
 * algorithm the real module uses.  tests/ check lev_myers == lev_dp.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

int64_t lev_dp(const uint8_t* a, int64_t la, const uint8_t* b, int64_t lb) {
  if (la == 0) return lb;
  if (lb == 0) return la;
  int64_t* row = (int64_t*)malloc((size_t)(lb + 1) * sizeof(int64_t));
  for (int64_t j = 0; j <= lb; ++j) row[j] = j;
  for (int64_t i = 1; i <= la; ++i) {
    int64_t diag = row[0];
    row[0] = i;
    for (int64_t j = 1; j <= lb; ++j) {
      int64_t up = row[j];
      int64_t best = diag + (a[i - 1] != b[j - 1]);
      if (up + 1 < best) best = up + 1;
      if (row[j - 1] + 1 < best) best = row[j - 1] + 1;
      row[j] = best;
      diag = up;
    }
  }
  int64_t d = row[lb];
  free(row);
  return d;
}

/* pattern = a (rows), text = b (columns); 64 pattern rows per word */
int64_t lev_myers(const uint8_t* a, int64_t la, const uint8_t* b, int64_t lb) {
  if (la == 0) return lb;
  if (lb == 0) return la;
  const int64_t nb = (la + 63) / 64;
  uint64_t* peq = (uint64_t*)calloc((size_t)nb * 256, sizeof(uint64_t));
  uint64_t* pv = (uint64_t*)malloc((size_t)nb * sizeof(uint64_t));
  uint64_t* mv = (uint64_t*)calloc((size_t)nb, sizeof(uint64_t));
  for (int64_t i = 0; i < la; ++i) peq[(i / 64) * 256 + a[i]] |= (uint64_t)1 << (i % 64);
  for (int64_t k = 0; k < nb; ++k) pv[k] = ~(uint64_t)0;
  const uint64_t last_top = (uint64_t)1 << ((la - 1) % 64);
  int64_t score = la;
  for (int64_t j = 0; j < lb; ++j) {
    int hin = 1; /* D[0][j] - D[0][j-1] = +1 */
    for (int64_t k = 0; k < nb; ++k) {
      uint64_t eq = peq[k * 256 + b[j]];
      const uint64_t Pv = pv[k], Mv = mv[k];
      const uint64_t top = (k == nb - 1) ? last_top : ((uint64_t)1 << 63);
      const uint64_t xv = eq | Mv;
      if (hin < 0) eq |= 1;
      const uint64_t xh = (((eq & Pv) + Pv) ^ Pv) | eq;
      uint64_t ph = Mv | ~(xh | Pv);
      uint64_t mh = Pv & xh;
      int hout = 0;
      if (ph & top) hout = 1;
      if (mh & top) hout = -1;
      ph <<= 1;
      mh <<= 1;
      if (hin < 0) mh |= 1;
      else if (hin > 0) ph |= 1;
      pv[k] = mh | ~(xv | ph);
      mv[k] = ph & xv;
      hin = hout;
    }
    score += hin;
  }
  free(peq);
  free(pv);
  free(mv);
  return score;
}

/* full symmetric matrix for n sequences stored concatenated; dist is n*n int64 */
void lev_matrix(const uint8_t* seqs, const int64_t* off, int64_t n, int use_myers, int64_t* dist) {
  for (int64_t i = 0; i < n; ++i) {
    dist[i * n + i] = 0;
    for (int64_t j = i + 1; j < n; ++j) {
      const uint8_t* a = seqs + off[i];
      const uint8_t* b = seqs + off[j];
      int64_t la = off[i + 1] - off[i], lb = off[j + 1] - off[j];
      int64_t d = use_myers ? lev_myers(a, la, b, lb) : lev_dp(a, la, b, lb);
      dist[i * n + j] = d;
      dist[j * n + i] = d;
    }
  }
}
