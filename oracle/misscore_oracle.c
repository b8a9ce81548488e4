/* ORACLE - TEST INFRASTRUCTURE ONLY.  Parity unpinned (Biopython is absent, see
 * oracle/pairwise2_oracle.py, which this file follows function by function).
 *
 * First alignment of Bio.pairwise2.align.globalms(a, b, match, mismatch, open, extend) for
 * integer parameters, as used by the reference's MisScore (src/PairwiseCompare.py:19-30):
 * full score and trace matrices (trace bits 1 = open gap in seqA, 2 = match/mismatch,
 * 4 = open gap in seqB, 8 = extend gap in seqA, 16 = extend gap in seqB), then the
 * stack-driven traceback with the dead-end rule (a gap in seqA may not follow a gap in seqB
 * on the way back) and the gap-open search for the extend bits.  Only the counts of the
 * alignment and its match line are produced, not the gapped strings.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int64_t row, col, len, match;
  int col_gap, trace;
} Ent;

typedef struct {
  Ent* v;
  int64_t n, cap;
} Stack;

static int push(Stack* s, Ent e) {
  if (s->n == s->cap) {
    int64_t cap = s->cap ? 2 * s->cap : 1024;
    Ent* v = (Ent*)realloc(s->v, (size_t)cap * sizeof(Ent));
    if (!v) return -1;
    s->v = v;
    s->cap = cap;
  }
  s->v[s->n++] = e;
  return 0;
}

static int64_t affine(int64_t length, int64_t open, int64_t extend) {
  /* calc_affine_penalty(length, open, extend, penalize_extend_when_opening=False) */
  if (length <= 0) return 0;
  return open + extend * length - extend;
}

/* out[0] score, out[1] alignment length, out[2] '|' count, out[3] stack pops.
 * line (optional, capacity la+lb): match line of the alignment ('|', '.', ' ').
 * Returns 0, -1 out of memory, -2 no alignment (empty input). */
int64_t pw2_first(const char* a, int64_t la, const char* b, int64_t lb, int match, int mismatch, int open,
                  int extend, int64_t* out, char* line) {
  if (la <= 0 || lb <= 0) return -2;
  const int64_t W = lb + 1;
  int32_t* S = (int32_t*)malloc((size_t)(la + 1) * W * sizeof(int32_t));
  uint8_t* T = (uint8_t*)calloc((size_t)(la + 1) * W, 1);
  int64_t* col_score = (int64_t*)malloc((size_t)W * sizeof(int64_t));
  if (!S || !T || !col_score) { free(S); free(T); free(col_score); return -1; }
  const int64_t first_gap = affine(1, open, extend);
  for (int64_t i = 0; i <= la; ++i) S[i * W] = (int32_t)affine(i, open, extend);
  for (int64_t j = 0; j <= lb; ++j) S[j] = (int32_t)affine(j, open, extend);
  col_score[0] = 0;
  for (int64_t j = 1; j <= lb; ++j) col_score[j] = affine(j, 2 * open, extend);
  for (int64_t r = 1; r <= la; ++r) {
    int64_t row_score = affine(r, 2 * open, extend);
    for (int64_t c = 1; c <= lb; ++c) {
      const int64_t nogap = S[(r - 1) * W + c - 1] + (a[r - 1] == b[c - 1] ? match : mismatch);
      const int64_t row_open = S[r * W + c - 1] + first_gap;
      const int64_t row_extend = row_score + extend;
      row_score = row_open > row_extend ? row_open : row_extend;
      const int64_t col_open = S[(r - 1) * W + c] + first_gap;
      const int64_t col_extend = col_score[c] + extend;
      col_score[c] = col_open > col_extend ? col_open : col_extend;
      int64_t best = nogap;
      if (col_score[c] > best) best = col_score[c];
      if (row_score > best) best = row_score;
      S[r * W + c] = (int32_t)best;
      int row_trace = 0, col_trace = 0, trace = 0;
      if (row_open == row_score) row_trace += 1;
      if (row_extend == row_score) row_trace += 8;
      if (col_open == col_score[c]) col_trace += 4;
      if (col_extend == col_score[c]) col_trace += 16;
      if (nogap == best) trace += 2;
      if (row_score == best) trace += row_trace;
      if (col_score[c] == best) trace += col_trace;
      T[r * W + c] = (uint8_t)trace;
    }
  }
  free(col_score);

  Stack st = {0, 0, 0};
  Ent e0 = {la, lb, 0, 0, 0, T[la * W + lb]};
  int64_t rc = push(&st, e0);
  int64_t pops = 0;
  int found = 0;
  Ent cur = e0;
  while (rc == 0 && st.n > 0 && !found) {
    cur = st.v[--st.n];
    ++pops;
    int dead = 0;
    int trace = cur.trace;
    while ((cur.row > 0 || cur.col > 0) && !dead) {
      Ent cache = cur;
      if (!trace) {
        if (cur.col && cur.col_gap) {
          dead = 1;
        } else {  /* _finish_backtrace: the rest of the longer sequence against gaps */
          /* one of row/col is 0 here; a literal '-' symbol against the gap character compares
           * equal in format_alignment and shows as '|' */
          while (cur.row > 0 || cur.col > 0) {
            const char sym = cur.row > 0 ? a[--cur.row] : b[--cur.col];
            if (line) line[cur.len] = sym == '-' ? '|' : ' ';
            cur.match += sym == '-';
            cur.len += 1;
          }
        }
        break;
      } else if (trace % 2 == 1) {
        trace -= 1;
        if (cur.col_gap) {
          dead = 1;
        } else {
          cur.col -= 1;
          if (line) line[cur.len] = b[cur.col] == '-' ? '|' : ' ';
          cur.match += b[cur.col] == '-';
          cur.len += 1;
          cur.col_gap = 0;
        }
      } else if (trace % 4 == 2) {
        trace -= 2;
        cur.row -= 1;
        cur.col -= 1;
        const int eq = a[cur.row] == b[cur.col];
        if (line) line[cur.len] = eq ? '|' : ((a[cur.row] == '-' || b[cur.col] == '-') ? ' ' : '.');
        cur.len += 1;
        cur.match += eq;
        cur.col_gap = 0;
      } else if (trace % 8 == 4) {
        trace -= 4;
        cur.row -= 1;
        if (line) line[cur.len] = a[cur.row] == '-' ? '|' : ' ';
        cur.match += a[cur.row] == '-';
        cur.len += 1;
        cur.col_gap = 1;
      } else if (trace == 8 || trace == 24 || trace == 16) {
        const int by_col = trace != 16;
        if (by_col) trace -= 8; else trace -= 16;
        if (by_col && cur.col_gap) {
          dead = 1;
        } else {  /* _find_gap_open */
          cur.col_gap = by_col ? 0 : 1;
          const int64_t target = by_col ? cur.col : cur.row;
          const int64_t target_score = S[cur.row * W + cur.col];
          for (int64_t n = 0; n < target; ++n) {
            if (by_col) cur.col -= 1; else cur.row -= 1;
            const char sym = by_col ? b[cur.col] : a[cur.row];
            if (line) line[cur.len] = sym == '-' ? '|' : ' ';
            cur.match += sym == '-';
            cur.len += 1;
            const int64_t actual = S[cur.row * W + cur.col] + affine(n + 1, open, extend);
            const int t_here = T[cur.row * W + cur.col];
            if (actual == target_score && n > 0) {
              if (!t_here) break;
              Ent alt = cur;
              alt.trace = t_here;
              if ((rc = push(&st, alt)) != 0) break;
            }
            if (!t_here) dead = 1;
          }
          if (rc) break;
        }
      }
      if (trace) {
        cache.trace = trace;
        if ((rc = push(&st, cache)) != 0) break;
      }
      trace = T[cur.row * W + cur.col];
    }
    if (!dead && rc == 0) found = 1;
  }
  if (rc == 0 && found) {
    out[0] = S[la * W + lb];
    out[1] = cur.len;
    out[2] = cur.match;
    out[3] = pops;
    if (line) {  /* built back to front */
      for (int64_t i = 0, j = cur.len - 1; i < j; ++i, --j) { char t = line[i]; line[i] = line[j]; line[j] = t; }
    }
  }
  free(st.v);
  free(S);
  free(T);
  if (rc) return -1;
  return found ? 0 : -2;
}
