// Sequence-to-graph alignment kernels for sm_100a (replace the dynamic programme and the
// traceback inside `spoa.poa(sequences, 1)`; reference call sites src/DataScanner.py:206,213
// and src/DecisionMaker.py:160,171).
//
// poa_dp_kernel   one CTA per (graph, read) alignment.  Rows (graph nodes in rank order) are
//                 swept top to bottom; the T threads of the CTA own 8 consecutive read
//                 columns each, so one pass covers a strip of 8*T columns and longer reads
//                 take several passes that hand the strip boundary (H, A, E, Q per row)
//                 through global memory.  Per row:
//                   phase 1  every thread folds the predecessor rows into its 8 cells
//                            (vertical / diagonal candidates, first-in-edge argmax);
//                   scan     the horizontal gap states E, Q are two max-plus prefix scans:
//                            8 cells in registers, warp shuffles, one shared-memory hop
//                            across warps, ONE __syncthreads per row;
//                   phase 2  H and the 16-bit traceback code of each cell; codes go to
//                            global memory (16 B per thread, coalesced), the packed row to
//                            the shared-memory ring and, if a far successor needs it, to
//                            global memory.
//                 Predecessor rows come from registers (rank-adjacent row), the ring of
//                 the last `ring_rows` packed rows in shared memory, or the exported rows.
// poa_tb_kernel   one thread per alignment walks the stored codes back to the origin.
//
// Integer arithmetic only; no tensor cores (nothing here is a dense contraction).
#include <cuda_runtime.h>

#include <cstdint>

#include "poa_cell.h"
#include "poa_kernels.h"
#include "poa_task.h"
#include "poa_window.h"

namespace svs {

namespace {

constexpr int kRowBatch = 32;    // rows whose metadata is staged in shared memory at once
constexpr int kPredCap = kRowBatch * 32;
constexpr int32_t kSrcRow0 = -1;
constexpr int32_t kSrcAdj = -2;
constexpr int32_t kSrcGlobal = 1 << 30;

struct WarpPub {  // what the last lanes of a warp publish for the warp to its right
  int32_t e31, e30, eloc7, q31, q30, qloc7, a7, pad;
};

struct Stage {
  uint32_t poff[kRowBatch + 1];
  int32_t psrc[kPredCap];
  int32_t pbh[kPredCap];
  int32_t xslot[kRowBatch];
  uint32_t single_before[kRowBatch];
  int16_t tlo[kRowBatch], thi[kRowBatch];   // active thread range of the row in this strip (empty: tlo > thi)
  int16_t ptlo[kPredCap], pthi[kPredCap];   // the same for every staged predecessor row
  int32_t cbase[kRowBatch];                 // first stored column - 1 of the row's code row (band-limited rows)
  int wlo, whi;                             // warps that own a band cell of some row of the batch (empty: wlo > whi)
  int32_t bA[kRowBatch], bE[kRowBatch], bQ[kRowBatch];
  uint8_t letter[kRowBatch];
  uint8_t flags[kRowBatch];
};

// The dynamic programme of one alignment, executed by the whole CTA.
// Band of a row = the read columns that are computed; everything outside counts as minus
// infinity (kNegBand).  Bands come from `band` ([row][lo,hi]; nullptr = all columns), see
// compute_bands and the "exact pruning" notes in poa_cell.h.
constexpr int kFull = 0;

template <int MODE>
__device__ __forceinline__ void row_band(const PoaTask& tk, const int32_t* band, uint32_t row, int32_t& lo, int32_t& hi) {
  if (band != nullptr) {
    const int2 b = __ldcg(reinterpret_cast<const int2*>(band) + row);
    lo = b.x; hi = b.y;
  } else {
    lo = 1; hi = static_cast<int32_t>(tk.L);
  }
}

// thread range [tlo, thi] of a band inside the strip [jb, je] (kC columns per thread)
template <int kC>
__device__ __forceinline__ void strip_threads(int32_t lo, int32_t hi, int32_t jb, int32_t je, int16_t& tlo, int16_t& thi) {
  if (lo > hi || hi < jb || lo > je) { tlo = 1; thi = 0; return; }
  tlo = static_cast<int16_t>((max(lo, jb) - jb) / kC);
  thi = static_cast<int16_t>((min(hi, je) - jb) / kC);
}

template <int T, int kC, int MODE>
__device__ __forceinline__ void dp_align(const PoaTask& tk, const Scores& s, const SingleTables& tabs,
                                         const int ring_rows, unsigned char* smem_raw, const int32_t* band) {
  static_assert(kC == 4 || kC == 8 || kC == 16, "columns per thread");
  const int32_t NEGW = pack_cell(kNegBand, kNeg, kNeg);
  // code row pitches: full-width rows, or (pruned) the widest band of the alignment (band[0..1])
  const uint64_t pitch1 = band != nullptr ? static_cast<uint32_t>(__ldcg(band)) : tk.w1;
  const uint64_t pitch2 = band != nullptr ? static_cast<uint32_t>(__ldcg(band + 1)) : tk.w2;
  constexpr int NW = T / 32;
  constexpr int WC = T * kC;
  int32_t* ring = reinterpret_cast<int32_t*>(smem_raw);
  WarpPub* pub = reinterpret_cast<WarpPub*>(ring + static_cast<size_t>(ring_rows) * WC);
  Stage& st = *reinterpret_cast<Stage*>(pub + 2 * NW);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t R = tk.R, L = tk.L;
  const uint64_t bstride = static_cast<uint64_t>(R) + 1;

  int32_t best = INT32_MIN;
  uint32_t best_row = 0;

  for (uint32_t pass = 0; pass < tk.npass; ++pass) {
    const uint32_t jb = 1 + pass * tk.strip;
    const uint32_t je = min(L, jb + tk.strip - 1);
    const uint32_t j0 = jb + kC * tid;
    const bool active = j0 <= je;
    const bool last_pass = (pass + 1 == tk.npass);
    const bool owns_end = last_pass && active && (L < j0 + kC);
    const int c_end = owns_end ? static_cast<int>(L - j0) : -1;
    const bool writes_bnd = !last_pass && (static_cast<uint32_t>(tid) == (je - jb) / kC);
    const int32_t* bin = tk.bnd + static_cast<uint64_t>(pass & 1) * 4 * bstride;
    int32_t* bout = tk.bnd + static_cast<uint64_t>((pass + 1) & 1) * 4 * bstride;

    int32_t rd[kC];
#pragma unroll
    for (int c = 0; c < kC; ++c) {
      const uint32_t j = j0 + c;
      rd[c] = (active && j <= L) ? static_cast<int32_t>(tk.read[j - 1]) : 0x100;
    }
    int32_t wprev[kC];       // packed cells of the previous row in my columns
    int32_t hleft_adj = 0;   // H[i-1][j0-1]
#pragma unroll
    for (int c = 0; c < kC; ++c) wprev[c] = 0;

    uint32_t slot = 1 % ring_rows;
    for (uint32_t i0 = 1; i0 <= R; i0 += kRowBatch) {
      const uint32_t nrows = min(static_cast<uint32_t>(kRowBatch), R - i0 + 1);
      __syncthreads();
      if (static_cast<uint32_t>(tid) < nrows) {
        const uint32_t i = i0 + tid;
        st.letter[tid] = tk.letter[i];
        st.flags[tid] = tk.flags[i];
        st.xslot[tid] = tk.xslot[i];
        st.single_before[tid] = tk.single_before[i];
        int32_t blo, bhi;
        row_band<MODE>(tk, band, i, blo, bhi);
        st.cbase[tid] = (band != nullptr && blo >= 1) ? ((blo - 1) / kC) * kC : 0;
        strip_threads<kC>(blo, bhi, static_cast<int32_t>(jb), static_cast<int32_t>(je), st.tlo[tid], st.thi[tid]);
        // the row's own left boundary: column 0 (exact) in the first strip, else the last
        // chunk of the previous strip if the band covered it
        const bool own_left = (pass == 0) || (blo <= static_cast<int32_t>(jb) - 1 && bhi >= static_cast<int32_t>(jb) - kC);
        if (pass == 0) {
          st.bA[tid] = tk.h0[i];
          st.bE[tid] = kNeg;
          st.bQ[tid] = kNeg;
        } else if (own_left) {
          st.bA[tid] = __ldcg(bin + bstride + i);
          st.bE[tid] = __ldcg(bin + 2 * bstride + i);
          st.bQ[tid] = __ldcg(bin + 3 * bstride + i);
        } else {
          st.bA[tid] = kNegBand;
          st.bE[tid] = kNeg;
          st.bQ[tid] = kNeg;
        }
        const uint32_t base = tk.pred_off[i0];
        const uint32_t pb = tk.pred_off[i], pe = tk.pred_off[i + 1];
        st.poff[tid] = pb - base;
        if (static_cast<uint32_t>(tid) == nrows - 1) st.poff[nrows] = pe - base;
        for (uint32_t e = pb; e < pe; ++e) {
          const uint32_t p = tk.preds[e];
          int32_t src, bh;
          int16_t ptlo = 0, pthi = static_cast<int16_t>(T);
          if (p == 0) {
            src = kSrcRow0;
            bh = row0_h(s, static_cast<int32_t>(jb) - 1);
          } else {
            int32_t plo, phi;
            row_band<MODE>(tk, band, p, plo, phi);
            strip_threads<kC>(plo, phi, static_cast<int32_t>(jb), static_cast<int32_t>(je), ptlo, pthi);
            const bool p_left = (pass == 0) || (plo <= static_cast<int32_t>(jb) - 1 && phi >= static_cast<int32_t>(jb) - kC);
            bh = (pass == 0) ? tk.h0[p] : (p_left ? __ldcg(bin + p) : kNegBand);
            if (p + 1 == i) src = kSrcAdj;
            else if (i - p <= static_cast<uint32_t>(ring_rows)) src = static_cast<int32_t>(p % ring_rows);
            else src = kSrcGlobal | tk.xslot[p];
          }
          st.psrc[e - base] = src;
          st.pbh[e - base] = bh;
          st.ptlo[e - base] = ptlo;
          st.pthi[e - base] = pthi;
        }
      }
      __syncthreads();
      if (tid == 0) {  // warps touched by the bands of this batch (bands are intervals => a contiguous range)
        int lo_t = T, hi_t = -1;
        for (uint32_t r = 0; r < nrows; ++r) {
          if (st.tlo[r] <= st.thi[r]) { lo_t = min(lo_t, static_cast<int>(st.tlo[r])); hi_t = max(hi_t, static_cast<int>(st.thi[r])); }
        }
        st.wlo = lo_t >> 5;
        st.whi = hi_t < 0 ? -1 : (hi_t >> 5);
      }
      __syncthreads();
      const int wlo = st.wlo, whi = st.whi;
      const int bar_threads = (whi - wlo + 1) * 32;
      // warps outside the range own no band cell in any row of the batch: they skip it; the
      // others synchronise among themselves with a named barrier
      if (warp < wlo || warp > whi) {
        for (uint32_t r = 0; r < nrows; ++r) slot = (slot + 1 == static_cast<uint32_t>(ring_rows)) ? 0 : slot + 1;
        continue;
      }

      for (uint32_t r = 0; r < nrows; ++r) {
        const uint32_t i = i0 + r;
        const uint32_t nb = st.poff[r], ne = st.poff[r + 1];
        const int32_t letter = st.letter[r];
        const bool single = (ne - nb == 1);
        if (st.tlo[r] > st.thi[r]) {   // the row has no cell in this strip (uniform)
          slot = (slot + 1 == static_cast<uint32_t>(ring_rows)) ? 0 : slot + 1;
          continue;
        }
        const bool t_active = active && tid >= st.tlo[r] && tid <= st.thi[r];
        CellAcc acc[kC];
#pragma unroll
        for (int c = 0; c < kC; ++c) { acc[c].Fm = 0; acc[c].Om = 0; acc[c].D = 0; acc[c].meta = 0; }

        // ---- phase 1: fold predecessor rows ------------------------------------------
        if (t_active) {
          for (uint32_t e = nb; e < ne; ++e) {
            const int32_t src = st.psrc[e];
            // cells of the predecessor row outside ITS band are minus infinity
            const bool chunk_ok = tid >= st.ptlo[e] && tid <= st.pthi[e];
            const bool left_ok = tid - 1 >= st.ptlo[e] && tid - 1 <= st.pthi[e];
            int32_t w[kC];
            int32_t hl;
            if (src != kSrcRow0 && !chunk_ok) {
#pragma unroll
              for (int c = 0; c < kC; ++c) w[c] = NEGW;
              hl = (tid == 0) ? st.pbh[e] : kNegBand;
              if (tid > 0 && left_ok) {
                if (src == kSrcAdj) hl = hleft_adj;
                else if (src & kSrcGlobal) hl = unpack_h(__ldcg(tk.xrows + static_cast<uint64_t>(src & ~kSrcGlobal) * tk.ldx + 3 + j0 - 1));
                else hl = unpack_h(ring[static_cast<size_t>(src) * WC + kC * tid - 1]);
              }
            } else if (src == kSrcAdj) {
#pragma unroll
              for (int c = 0; c < kC; ++c) w[c] = wprev[c];
              hl = (tid == 0) ? st.pbh[e] : (left_ok ? hleft_adj : kNegBand);
            } else if (src == kSrcRow0) {
#pragma unroll
              for (int c = 0; c < kC; ++c) w[c] = pack_cell(row0_h(s, static_cast<int32_t>(j0) + c), kNeg, kNeg);
              hl = row0_h(s, static_cast<int32_t>(j0) - 1);
            } else if (src & kSrcGlobal) {
              const int32_t* row = tk.xrows + static_cast<uint64_t>(src & ~kSrcGlobal) * tk.ldx + 3;
#pragma unroll
              for (int q = 0; q < kC / 4; ++q) {
                const int4 v = __ldcg(reinterpret_cast<const int4*>(row + j0 + 4 * q));
                w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
              }
              hl = (tid == 0 || left_ok) ? unpack_h(__ldcg(row + j0 - 1)) : kNegBand;
              if (tid == 0) hl = st.pbh[e];
            } else {
              const int32_t* row = ring + static_cast<size_t>(src) * WC;
#pragma unroll
              for (int q = 0; q < kC / 4; ++q) {
                const int4 v = *reinterpret_cast<const int4*>(row + kC * tid + 4 * q);
                w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
              }
              hl = (tid == 0) ? st.pbh[e] : (left_ok ? unpack_h(row[kC * tid - 1]) : kNegBand);
            }
            if (single) {
#pragma unroll
              for (int c = 0; c < kC; ++c) {
                cell_pred_single(acc[c], w[c], hl, (letter == rd[c]) ? s.m : s.n, s, tabs);
                hl = unpack_h(w[c]);
              }
            } else {
              const uint32_t k = e - nb;
#pragma unroll
              for (int c = 0; c < kC; ++c) {
                cell_pred_key(acc[c], k, w[c], hl, (letter == rd[c]) ? s.m : s.n, s, tabs);
                hl = unpack_h(w[c]);
              }
            }
          }
        }

        // ---- scan: horizontal gap states across the row ----------------------------------
        int32_t a7 = kNegBand;
        int32_t el = kNeg, ql = kNeg, eloc7 = kNeg, qloc7 = kNeg;
        if (t_active) {
#pragma unroll
          for (int c = 0; c < kC; ++c) {
            const int32_t A = single ? imax(acc[c].D, imax(acc[c].Fm, acc[c].Om))
                                     : imax(key_value(acc[c].D), key_value(static_cast<int32_t>(acc[c].meta)));
            if (c == kC - 1) { eloc7 = el; qloc7 = ql; a7 = imax(A, kNegBand); }
            el = imax(A + s.g, el + s.e);
            ql = imax(A + s.q, ql + s.c);
          }
        }
        const int32_t bA = st.bA[r], bE = st.bE[r], bQ = st.bQ[r];
        int32_t ein0 = 0, qin0 = 0;
        if (tid == 0) {  // the strip boundary enters through thread 0
          ein0 = imax(bA + s.g, bE + s.e);
          qin0 = imax(bA + s.q, bQ + s.c);
          el = imax(el, ein0 + kC * s.e);
          ql = imax(ql, qin0 + kC * s.c);
        }
        int32_t ve = el, vq = ql;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int32_t oe = __shfl_up_sync(0xffffffffu, ve, d);
          const int32_t oq = __shfl_up_sync(0xffffffffu, vq, d);
          if (lane >= d) {
            ve = imax(ve, oe + kC * s.e * d);
            vq = imax(vq, oq + kC * s.c * d);
          }
        }
        const int par = static_cast<int>(i & 1);
        {
          const int32_t e30 = __shfl_sync(0xffffffffu, ve, 30);
          const int32_t q30 = __shfl_sync(0xffffffffu, vq, 30);
          if (lane == 31) {
            WarpPub pb;
            pb.e31 = ve; pb.e30 = e30; pb.eloc7 = eloc7;
            pb.q31 = vq; pb.q30 = q30; pb.qloc7 = qloc7;
            pb.a7 = a7; pb.pad = 0;
            pub[par * NW + warp] = pb;
          }
        }
        asm volatile("bar.sync 1, %0;" ::"r"(bar_threads) : "memory");
        int32_t t1e = kNeg, t1q = kNeg, t2e = kNeg, t2q = kNeg;
        if (warp > wlo) {
          if (lane < warp && lane >= wlo) {
            const WarpPub pb = pub[par * NW + lane];
            const int d1 = warp - 1 - lane;
            t1e = pb.e31 + 32 * kC * s.e * d1;
            t1q = pb.q31 + 32 * kC * s.c * d1;
            if (lane < warp - 1) {
              t2e = pb.e31 + 32 * kC * s.e * (d1 - 1);
              t2q = pb.q31 + 32 * kC * s.c * (d1 - 1);
            }
          }
          t1e = __reduce_max_sync(0xffffffffu, t1e);
          t1q = __reduce_max_sync(0xffffffffu, t1q);
          t2e = __reduce_max_sync(0xffffffffu, t2e);
          t2q = __reduce_max_sync(0xffffffffu, t2q);
        }
        const int32_t vte = imax(ve, t1e + kC * s.e * (lane + 1));
        const int32_t vtq = imax(vq, t1q + kC * s.c * (lane + 1));
        int32_t ein = __shfl_up_sync(0xffffffffu, vte, 1);
        int32_t qin = __shfl_up_sync(0xffffffffu, vtq, 1);
        if (lane == 0) { ein = t1e; qin = t1q; }
        if (tid == 0) { ein = ein0; qin = qin0; }
        const int32_t se = imax(ein + (kC - 1) * s.e, eloc7);  // E, Q at my last column
        const int32_t sq = imax(qin + (kC - 1) * s.c, qloc7);
        RowCarry cy;
        cy.A = __shfl_up_sync(0xffffffffu, a7, 1);
        cy.E = __shfl_up_sync(0xffffffffu, se, 1);
        cy.Q = __shfl_up_sync(0xffffffffu, sq, 1);
        if (lane == 0) {
          if (warp == 0) {
            cy.A = bA; cy.E = bE; cy.Q = bQ;
          } else if (warp == wlo) {   // everything to the left is outside every band of the batch
            cy.A = kNegBand; cy.E = kNeg; cy.Q = kNeg;
          } else {
            const WarpPub pb = pub[par * NW + warp - 1];
            const int32_t einl = imax(pb.e30, t2e + kC * s.e * 31);
            const int32_t qinl = imax(pb.q30, t2q + kC * s.c * 31);
            cy.A = pb.a7;
            cy.E = imax(einl + (kC - 1) * s.e, pb.eloc7);
            cy.Q = imax(qinl + (kC - 1) * s.c, pb.qloc7);
          }
        }
        cy.H = imax(cy.A, imax(cy.E, cy.Q));
        hleft_adj = cy.H;

        // ---- phase 2: H, traceback codes, packed row --------------------------------------
        if (t_active) {
          uint32_t cw[kC / 2];
          int32_t hsel = INT32_MIN;
          if (single) {
#pragma unroll
            for (int c = 0; c < kC; ++c) {
              int32_t H;
              const uint32_t cd = cell_finish_single(acc[c], cy, s, H);
              H = imax(H, kNegBand); cy.H = H; cy.A = imax(cy.A, kNegBand);   // pruned neighbours must not drift
              wprev[c] = pack_cell(H, acc[c].Fm, acc[c].Om);
              if (c & 1) cw[c >> 1] |= cd << 16; else cw[c >> 1] = cd;
              if (c == c_end) hsel = H;
            }
          } else {
#pragma unroll
            for (int c = 0; c < kC; ++c) {
              int32_t H, Fv, Ov;
              const uint32_t cd = cell_finish_key(acc[c], cy, s, H, Fv, Ov);
              H = imax(H, kNegBand); cy.H = H; cy.A = imax(cy.A, kNegBand);
              wprev[c] = pack_cell(H, Fv, Ov);
              if (c & 1) cw[c >> 1] |= cd << 16; else cw[c >> 1] = cd;
              if (c == c_end) hsel = H;
            }
          }
          if (MODE == kFull) {
            const uint64_t n1 = st.single_before[r];
            uint8_t* crow = tk.codes + n1 * pitch1 + (static_cast<uint64_t>(i - 1) - n1) * pitch2;
            const uint64_t ccol = static_cast<uint64_t>(j0 - 1) - static_cast<uint32_t>(st.cbase[r]);
            if (single) {  // single predecessor: low bytes only
              uint32_t b[kC / 4];
#pragma unroll
              for (int q = 0; q < kC / 4; ++q)
                b[q] = (cw[2 * q] & 0xffu) | ((cw[2 * q] >> 8) & 0xff00u) | ((cw[2 * q + 1] & 0xffu) << 16) |
                       ((cw[2 * q + 1] & 0xff0000u) << 8);
              if (kC == 4) *reinterpret_cast<uint32_t*>(crow + ccol) = b[0];
              else if (kC == 8) *reinterpret_cast<uint2*>(crow + ccol) = make_uint2(b[0], b[kC / 4 - 1]);
              else *reinterpret_cast<uint4*>(crow + ccol) = make_uint4(b[0], b[1], b[kC / 4 - 2], b[kC / 4 - 1]);
            } else if (kC == 4) {
              *reinterpret_cast<uint2*>(crow + 2 * ccol) = make_uint2(cw[0], cw[1]);
            } else {
#pragma unroll
              for (int q = 0; q < kC / 8; ++q)
                *reinterpret_cast<uint4*>(crow + 2 * ccol + 16 * q) =
                    make_uint4(cw[4 * q], cw[4 * q + 1], cw[4 * q + 2], cw[4 * q + 3]);
            }
          }
          int32_t* rrow = ring + static_cast<size_t>(slot) * WC + kC * tid;
#pragma unroll
          for (int q = 0; q < kC / 4; ++q)
            *reinterpret_cast<int4*>(rrow + 4 * q) = make_int4(wprev[4 * q], wprev[4 * q + 1], wprev[4 * q + 2], wprev[4 * q + 3]);
          if (st.flags[r] & kFlagExport) {
            int32_t* xrow = tk.xrows + static_cast<uint64_t>(st.xslot[r]) * tk.ldx + 3;
#pragma unroll
            for (int q = 0; q < kC / 4; ++q)
              *reinterpret_cast<int4*>(xrow + j0 + 4 * q) = make_int4(wprev[4 * q], wprev[4 * q + 1], wprev[4 * q + 2], wprev[4 * q + 3]);
            if (tid == 0 && pass == 0) xrow[0] = pack_cell(bA, kNeg, kNeg);
          }
          if (writes_bnd) {
            bout[i] = cy.H;
            bout[bstride + i] = cy.A;
            bout[2 * bstride + i] = cy.E;
            bout[3 * bstride + i] = cy.Q;
          }
          if (owns_end && (st.flags[r] & kFlagSink) && hsel > best) {
            best = hsel;
            best_row = i;
          }
        }
        slot = (slot + 1 == static_cast<uint32_t>(ring_rows)) ? 0 : slot + 1;
      }
    }
    if (owns_end) {
      tk.result[0] = static_cast<int32_t>(best_row);
      tk.result[1] = best;
    }
    __syncthreads();
  }
}

// Band of every row: the columns whose upper bound (cell_bound, concave in the column) reaches
// the lower bound `lb` of the optimal score.
template <int T, int kC>
__device__ void compute_bands(const PoaTask& tk, const Scores& s, int32_t lb, bool have_lb, int32_t* band) {
  const int32_t L = static_cast<int32_t>(tk.L);
  __shared__ int s_widest;
  if (threadIdx.x == 0) s_widest = kC;
  __syncthreads();
  int widest = kC;
  for (uint32_t i = 1 + threadIdx.x; i <= tk.R; i += T) {
    int32_t lo = 1, hi = L;
    if (have_lb) {
      const int4 d = *(reinterpret_cast<const int4*>(tk.depth) + i);   // written by this CTA (device-resident graph): no read-only path
      auto ub = [&](int32_t j) { return cell_bound(s, d.x, d.y, d.z, d.w, j, L); };
      // maximiser: one of the breakpoints of the two concave pieces
      int32_t cand[6] = {1, L, d.x, d.y, L - d.w, L - d.z};
      int32_t jm = 1, best = INT32_MIN;
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const int32_t j = min(L, max(1, cand[k]));
        const int32_t v = ub(j);
        if (v > best) { best = v; jm = j; }
      }
      if (best < lb) {
        lo = 1; hi = 0;  // no cell of this row can be on a co-optimal path
      } else {
        int32_t a = 1, b = jm;      // first column with ub >= lb
        while (a < b) { const int32_t mid = (a + b) >> 1; if (ub(mid) >= lb) b = mid; else a = mid + 1; }
        lo = a;
        a = jm; b = L;              // last column with ub >= lb
        while (a < b) { const int32_t mid = (a + b + 1) >> 1; if (ub(mid) >= lb) a = mid; else b = mid - 1; }
        hi = a;
      }
    }
    band[2 * i] = lo;
    band[2 * i + 1] = hi;
    if (lo <= hi) widest = max(widest, ((hi - 1) / kC - (lo - 1) / kC + 1) * kC);   // stored cells of the row
  }
  atomicMax(&s_widest, widest);
  __syncthreads();
  if (threadIdx.x == 0) {   // row 0 has no codes: its two entries carry the row pitches (bytes)
    const int32_t p1 = (s_widest + 15) / 16 * 16;
    band[0] = p1;
    band[1] = 2 * p1;
  }
  __syncthreads();
}

// Traceback by one warp.  Long diagonal runs through chain rows (one predecessor = the
// previous row) are the common case: the 32 lanes look at the cells (i-k, j-k) in parallel
// and the walk advances by the number of leading lanes whose cell is such a diagonal move;
// everything else is one serial step of the reference walk (tb_step) by lane 0.
__device__ void tb_walk_warp(const PoaTask& tk, const Scores& s, const int32_t* band, int cols) {
  const uint32_t p1 = band != nullptr ? static_cast<uint32_t>(band[0]) : tk.w1;
  const uint32_t p2 = band != nullptr ? static_cast<uint32_t>(band[1]) : tk.w2;
  const TbView v{tk.codes, p1, p2, band, static_cast<uint32_t>(cols), tk.single_before, tk.col0code, tk.pred_off, tk.preds, tk.node_id};
  const int lane = threadIdx.x & 31;
  uint32_t i = static_cast<uint32_t>(tk.result[0]), j = tk.L;
  int32_t n = 0;
  const int32_t cap = static_cast<int32_t>(tk.path_cap);
  bool ok = true;
  while (ok && !(i == 0 && j == 0)) {
    bool mine = false;
    int32_t node = 0;
    if (i > static_cast<uint32_t>(lane) && j > static_cast<uint32_t>(lane)) {
      const uint32_t r = i - lane, c = j - lane;
      bool inside = true;
      uint32_t col = c - 1;
      if (band != nullptr) {   // a speculated cell may lie outside the stored band of its row
        const int32_t blo = band[2 * r], bhi = band[2 * r + 1];
        const uint32_t first = blo >= 1 ? (static_cast<uint32_t>(blo - 1) / cols) * cols : 0u;
        inside = blo <= bhi && c - 1 >= first && static_cast<int32_t>(c) <= bhi;
        col -= first;
      }
      if (inside && (tk.flags[r] & kFlagChain)) {
        const uint64_t n1 = tk.single_before[r];
        const uint32_t cd = tk.codes[n1 * p1 + (static_cast<uint64_t>(r - 1) - n1) * p2 + col];
        if ((cd & 3u) == kMoveDiag) {
          mine = true;
          node = static_cast<int32_t>(tk.node_id[r]);
        }
      }
    }
    const unsigned hit = __ballot_sync(0xffffffffu, mine);
    const int m = __ffs(~hit) - 1;   // leading lanes with a chain diagonal (32 when all)
    const int run = m < 0 ? 32 : m;
    if (run > 0) {
      if (n + run > cap) { ok = false; break; }
      if (lane < run) {
        tk.path[2 * (n + lane)] = node;
        tk.path[2 * (n + lane) + 1] = static_cast<int32_t>(j - lane - 1);
      }
      n += run;
      i -= run;
      j -= run;
    } else {
      if (lane == 0) ok = tb_step(v, s, i, j, n, tk.path, cap);
      i = __shfl_sync(0xffffffffu, i, 0);
      j = __shfl_sync(0xffffffffu, j, 0);
      n = __shfl_sync(0xffffffffu, n, 0);
      ok = __shfl_sync(0xffffffffu, static_cast<int>(ok), 0) != 0;
    }
  }
  if (lane == 0) tk.result[2] = ok ? n : -1;
}

// ---------------------------------------------------------------------------------------------
// Window kernel: the CTA owns a window (group of sequences) from its first to its last read.
// Per read: rank-ordered export of the graph (poa_dgraph.h) -> bands -> dynamic programme ->
// traceback -> merge of the path into the graph -> rank order; at the end MSA rows and the
// heaviest-bundle consensus are written to the output arena.  The graph never leaves the
// device; the host only sees the result record of the window.
struct CtaExec {
  uint32_t* warp_tot;   // shared memory, 32 words
  int32_t* sweep_ring;  // shared memory, 2 x 64 x 2 words (depth sweeps)
  template <class F> __device__ __forceinline__ void run(F f) { f(threadIdx.x, blockDim.x); __syncthreads(); }
  template <class F> __device__ __forceinline__ void one(F f) { if (threadIdx.x == 0) f(); __syncthreads(); }
  template <class F, class G> __device__ __forceinline__ void two(F f, G g) {
    if (threadIdx.x == 0) f(); else if (threadIdx.x == 32) g();
    __syncthreads();
  }
  __device__ __forceinline__ void atomic_max(uint32_t* p, uint32_t v) { atomicMax(p, v); }
  __device__ __forceinline__ void atomic_min(int32_t* p, int32_t v) { atomicMin(p, v); }
  __device__ __forceinline__ void atomic_add(uint32_t* p, uint32_t v) { atomicAdd(p, v); }
  __device__ void scan(uint32_t* a, uint32_t n) {
    const uint32_t nt = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    uint32_t carry = 0;
    for (uint32_t base = 0; base < n; base += nt) {
      const uint32_t i = base + tid;
      uint32_t v = i < n ? a[i] : 0;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= static_cast<uint32_t>(d)) v += o;
      }
      if (lane == 31) warp_tot[warp] = v;
      __syncthreads();
      uint32_t add = carry, total = 0;
      for (uint32_t w = 0; w < nw; ++w) {
        const uint32_t t = warp_tot[w];
        if (w < warp) add += t;
        total += t;
      }
      if (i < n) a[i] = v + add;
      carry += total;
      __syncthreads();
    }
  }
  // a[i] = min(a[i..n))
  __device__ void suffix_min(int32_t* a, uint32_t n) {
    const uint32_t nt = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    int32_t* tot = reinterpret_cast<int32_t*>(warp_tot);
    int32_t carry = INT32_MAX;
    if (n == 0) return;
    for (int64_t base = static_cast<int64_t>((n - 1) / nt) * nt; base >= 0; base -= nt) {
      const uint32_t i = static_cast<uint32_t>(base) + tid;
      int32_t v = i < n ? a[i] : INT32_MAX;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int32_t o = __shfl_down_sync(0xffffffffu, v, d);
        if (lane + d < 32) v = min(v, o);
      }
      if (lane == 0) tot[warp] = v;
      __syncthreads();
      int32_t add = carry, total = INT32_MAX;
      for (uint32_t w = 0; w < nw; ++w) {
        const int32_t t = tot[w];
        if (w > warp) add = min(add, t);
        total = min(total, t);
      }
      if (i < n) a[i] = min(v, add);
      carry = min(carry, total);
      __syncthreads();
    }
  }
  // Path-length intervals (dg_depth_forward / dg_depth_backward): warp 0 sweeps forward, warp 1
  // backward (a CTA of one warp does both in turn).  A row is one step of the warp: the lanes
  // fetch the row's predecessors in parallel; values of the last 64 rows live in a ring in
  // shared memory, older ones in global memory.
  __device__ void depth_sweeps(const WinMem& m, uint32_t R) {
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    int32_t* dp = m.depth;
    // partial values of the backward sweep start as "no successor seen"
    for (uint32_t i = tid; i <= R; i += blockDim.x) { dp[4 * i + 2] = INT32_MAX; dp[4 * i + 3] = -1; }
    __syncthreads();
    if (warp == 0) {
      int32_t* ring = sweep_ring;                       // [64][2]
      if (lane == 0) { ring[0] = 0; ring[1] = 0; }      // row 0
      __syncwarp();
      for (uint32_t i0 = 1; i0 <= R; i0 += 32) {
        const uint32_t nrows = min(32u, R - i0 + 1);
        const uint32_t pb = lane < nrows ? m.pred_off[i0 + lane] : 0;
        const uint32_t pe = lane < nrows ? m.pred_off[i0 + lane + 1] : 0;
        for (uint32_t r = 0; r < nrows; ++r) {
          const uint32_t i = i0 + r;
          const uint32_t b = __shfl_sync(0xffffffffu, pb, r), e = __shfl_sync(0xffffffffu, pe, r);
          int32_t lo = INT32_MAX, hi = 0;
          for (uint32_t k = b + lane; k < e; k += 32) {
            const uint32_t p = m.preds[k];
            int32_t a, c;
            if (i - p < 64) { a = ring[2 * (p & 63)]; c = ring[2 * (p & 63) + 1]; }
            else { a = dp[4 * p]; c = dp[4 * p + 1]; }
            lo = min(lo, a);
            hi = max(hi, c);
          }
          lo = __reduce_min_sync(0xffffffffu, lo) + 1;
          hi = __reduce_max_sync(0xffffffffu, hi) + 1;
          if (lane == 0) {
            ring[2 * (i & 63)] = lo; ring[2 * (i & 63) + 1] = hi;
            dp[4 * i] = lo; dp[4 * i + 1] = hi;
          }
          __syncwarp();
        }
      }
    }
    if (warp == (nw > 1 ? 1u : 0u)) {
      int32_t* ring = sweep_ring + 128;                 // [64][2] partial (smin, smax) of rows [i-63, i]
      for (uint32_t k = lane; k < 64; k += 32) { ring[2 * k] = INT32_MAX; ring[2 * k + 1] = -1; }
      __syncwarp();
      for (uint32_t i1 = R; i1 >= 1; i1 -= min(i1, 32u)) {
        // rows i1, i1-1, ... (up to 32), lane l holds the predecessor range of row i1 - l
        const uint32_t nrows = min(32u, i1);
        const uint32_t pb = lane < nrows ? m.pred_off[i1 - lane] : 0;
        const uint32_t pe = lane < nrows ? m.pred_off[i1 - lane + 1] : 0;
        for (uint32_t r = 0; r < nrows; ++r) {
          const uint32_t i = i1 - r;
          const uint32_t b = __shfl_sync(0xffffffffu, pb, r), e = __shfl_sync(0xffffffffu, pe, r);
          int32_t lo = ring[2 * (i & 63)], hi = ring[2 * (i & 63) + 1];
          if (lo == INT32_MAX) { lo = 0; hi = 0; }   // no successor
          const int32_t a = lo + 1, c = hi + 1;
          __syncwarp();
          if (lane == 0) { dp[4 * i + 2] = lo; dp[4 * i + 3] = hi; }
          for (uint32_t k = b + lane; k < e; k += 32) {   // the predecessors of a row are distinct rows
            const uint32_t p = m.preds[k];
            if (p == 0) continue;
            if (i - p < 64) {
              ring[2 * (p & 63)] = min(ring[2 * (p & 63)], a);
              ring[2 * (p & 63) + 1] = max(ring[2 * (p & 63) + 1], c);
            } else {
              dp[4 * p + 2] = min(dp[4 * p + 2], a);
              dp[4 * p + 3] = max(dp[4 * p + 3], c);
            }
          }
          __syncwarp();
          // the slot now belongs to row i - 64: start from what successors 64 or more rows away left for it
          if (lane == 0 && i > 64) { ring[2 * (i & 63)] = dp[4 * (i - 64) + 2]; ring[2 * (i & 63) + 1] = dp[4 * (i - 64) + 3]; }
          __syncwarp();
        }
        if (i1 <= 32) break;
      }
      if (lane == 0) { dp[2] = 0; dp[3] = 0; }
    }
    __syncthreads();
  }
};

#include "poa_dp2.cuh"

template <int T, int kC>
__global__ void __launch_bounds__(T, (T == 128 ? 4 : (T == 256 && kC == 8 ? 2 : 1)))
poa_window_kernel(const WinParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ WinState S;
  __shared__ WinMem m;
  __shared__ PoaTask tk;
  __shared__ WinCaps caps;
  __shared__ uint32_t warp_tot[32];
  __shared__ unsigned long long cyc[8];
  __shared__ int s_next, s_slot;
  __shared__ uint64_t s_cells, s_rows, s_exported, s_need, s_pairs, s_bases, s_steps, s_preds;
  __shared__ uint32_t s_nalign, s_retries;
  __shared__ unsigned long long s_need2, s_eval[5];
  __shared__ int32_t sweep_ring[256];
  CtaExec x{warp_tot, sweep_ring};
  const int tid = threadIdx.x;
  if (P.sm_limit > 0) {   // profiling aid: full per-SM occupancy on a few SMs only (small arena, short ncu replays)
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    if (smid >= static_cast<unsigned>(P.sm_limit)) return;
  }
  // scratch slot: any free one (at most n_slots CTAs of this kernel are resident at a time)
  if (tid == 0) {
    int k = static_cast<int>((blockIdx.x * 7919u) % static_cast<unsigned>(P.n_slots));
    while (atomicCAS(&P.slot_flags[k], 0, 1) != 0) k = (k + 1 == P.n_slots) ? 0 : k + 1;
    s_slot = k;
  }
  __syncthreads();
  uint8_t* const slot = P.slot_base + static_cast<uint64_t>(s_slot) * P.slot_bytes;
  const Scores s = P.s;
  while (true) {
    __syncthreads();
    if (tid == 0) s_next = atomicAdd(P.counter, 1);
    __syncthreads();
    if (s_next >= P.n_windows) break;
    const int widx = P.order[s_next];
    const WinDesc d = P.desc[widx];
    if (tid == 0) {
      caps = d.caps;
      const uint64_t fixed = win_layout(slot, P.slot_bytes, d.caps, &m);
      S.nv = S.ne = S.nseq = 0;
      S.path_off = 0;
      S.err = fixed + 4096 > P.slot_bytes ? kWinNodeCap : kWinOk;
      S.max_indeg = 1; S.n_export = S.n_single = S.n_new = S.msa_cols = 0;
      S.last_score = 0; S.last_len = 0;
      S.err_pending = 0; S.topo_serial = 0; S.topo_rounds = 0; S.chg[0] = S.chg[1] = 0;
      for (int k = 0; k < 8; ++k) cyc[k] = 0;
      s_cells = s_rows = s_exported = s_need = s_pairs = s_bases = s_steps = s_preds = 0;
      s_nalign = s_retries = 0;
      for (int k = 0; k < 5; ++k) s_eval[k] = 0;
    }
    __syncthreads();
    for (uint32_t q = 0; q < d.caps.nseq && S.err == kWinOk; ++q) {
      const int64_t id = P.members[d.member_begin + q];
      const int64_t o0 = P.read_off[id];
      const uint32_t L = static_cast<uint32_t>(P.read_off[id + 1] - o0);
      const uint8_t* seq = P.reads + o0;
      if (P.pair_cnt != nullptr && tid == 0) P.pair_cnt[d.member_begin + q] = 0;
      if (L == 0) continue;                       // empty sequences are ignored (no MSA row)
      if (S.nseq == 0) {
        dg_init_chain(x, m, caps, &S, seq, L);
        continue;
      }
      long long t0 = clock64();
      dg_export(x, m, caps, &S, s, static_cast<uint32_t>(P.ring_rows), false);
      if (tid == 0) {
        const long long t1 = clock64();
        cyc[0] += static_cast<unsigned long long>(t1 - t0);
        const uint32_t R = S.nv;
        tk.letter = m.r_letter; tk.pred_off = m.pred_off; tk.preds = m.preds; tk.flags = m.r_flags;
        tk.xslot = m.xslot; tk.h0 = m.h0; tk.col0code = m.col0code; tk.node_id = m.node_id;
        tk.single_before = m.single_before; tk.depth = m.depth; tk.read = seq;
        tk.R = R; tk.L = L;
        const uint32_t cpp = T * kC;
        tk.npass = (L + cpp - 1) / cpp;
        tk.strip = ((L + tk.npass - 1) / tk.npass + kC - 1) / kC * kC;
        tk.w1 = static_cast<uint32_t>((static_cast<uint64_t>(L) + kC - 1 + 15) / 16 * 16);
        tk.w2 = static_cast<uint32_t>((static_cast<uint64_t>(L) + kC - 1 + 7) / 8 * 8 * 2);
        tk.ldx = (static_cast<uint64_t>(L) + 3 + kC + 7) / 8 * 8;
        const uint64_t xbytes = dg_align(static_cast<uint64_t>(S.n_export) * tk.ldx * 4, 256);
        tk.xrows = reinterpret_cast<int32_t*>(m.dyn);
        tk.codes = m.dyn + xbytes;
        tk.codes_cap = m.dyn_bytes > xbytes ? m.dyn_bytes - xbytes : 0;
        tk.coff = m.coff;
        tk.bnd = m.bnd; tk.result = m.result; tk.path = m.path; tk.path_cap = m.path_cap;
        const int64_t span = 10ll * (static_cast<int64_t>(R) + L + 2);
        tk.prune = (P.prune && L >= 1024 && R >= 1024 && span < (1 << 21)) ? 1u : 0u;
        const double spb = S.last_len ? static_cast<double>(S.last_score) / S.last_len : 4.0;
        tk.lb_guess = static_cast<int32_t>((spb - P.prune_margin) * static_cast<double>(L)) - 40;
        const uint64_t n1 = S.n_single;
        const uint64_t full = n1 * tk.w1 + (static_cast<uint64_t>(R) - n1) * tk.w2 + 64;
        if (S.max_indeg > kMaxIndeg) S.err = kWinIndeg;
        else if (span >= kMaxKeySpan) S.err = kWinScoreSpan;
        else if (xbytes + 4096 > m.dyn_bytes || (!tk.prune && !(kC == 8 && P.dp_version == 2) && full > tk.codes_cap)) { S.err = kWinCodesCap; s_need = full + xbytes; }
        s_cells += (static_cast<uint64_t>(R) + 1) * (static_cast<uint64_t>(L) + 1);
        s_rows += R;
        s_exported += S.n_export;
        s_bases += L;
        s_preds += m.pred_off[R + 1];
        s_nalign += 1;
        m.result[3] = 0;
      }
      __syncthreads();
      if (S.err != kWinOk) break;
      t0 = clock64();
      int32_t* band = nullptr;
      bool v2_overflow = false;
      if (kC == 8 && P.dp_version == 2) {
        // warp-pipelined kernel: bands (pruned: from the guessed lower bound; else full rows) and
        // exact-size code rows; a result below the guess repeats the alignment with the score found
        int32_t lb = tk.lb_guess;
        bool have_lb = tk.prune != 0;
        for (int attempt = 0; attempt < 3; ++attempt) {
          if (tid == 0) { tk.result[0] = 0; tk.result[1] = INT32_MIN; }
          compute_bands2<T>(x, tk, s, lb, have_lb, m.band, m.coff, &s_need2);
          if (s_need2 > tk.codes_cap) { v2_overflow = true; break; }
          dp2_align<T>(tk, s, P.tabs, P.ring_rows, smem_raw, m.band, m.coff, s_eval);
          const int32_t found_row = tk.result[0], found = tk.result[1];
          if (!have_lb || (found_row > 0 && found >= lb)) break;
          if (tid == 0) s_retries += 1;
          have_lb = found_row > 0 && found > kNegBand / 2;
          lb = found;
          __syncthreads();
        }
        if (v2_overflow) {
          if (tid == 0) { S.err = kWinCodesCap; s_need = s_need2; }
          __syncthreads();
          break;
        }
      } else
      {
      if (tk.prune) {
        // exact pruning with a guessed lower bound (see poa_persistent_kernel)
        band = m.band;
        int32_t lb = tk.lb_guess;
        bool have_lb = true, overflow = false;
        for (int attempt = 0; attempt < 3; ++attempt) {
          if (tid == 0) { tk.result[0] = 0; tk.result[1] = INT32_MIN; }
          compute_bands<T, kC>(tk, s, lb, have_lb, band);
          {
            const uint64_t n1 = tk.single_before[tk.R + 1];
            const uint64_t need = n1 * static_cast<uint32_t>(band[0]) + (static_cast<uint64_t>(tk.R) - n1) * static_cast<uint32_t>(band[1]) + 64;
            if (need > tk.codes_cap) { overflow = true; if (tid == 0) s_need = need; break; }
          }
          dp_align<T, kC, kFull>(tk, s, P.tabs, P.ring_rows, smem_raw, band);
          const int32_t found_row = tk.result[0], found = tk.result[1];
          if (!have_lb || (found_row > 0 && found >= lb)) break;
          if (tid == 0) s_retries += 1;
          have_lb = found_row > 0 && found > kNegBand / 2;
          lb = found;
          __syncthreads();
        }
        if (overflow) {
          if (tid == 0) S.err = kWinCodesCap;
          __syncthreads();
          break;
        }
      } else {
        dp_align<T, kC, kFull>(tk, s, P.tabs, P.ring_rows, smem_raw, nullptr);
      }
      }
      long long t1 = clock64();
      if (tid < 32) {
        if (kC == 8 && P.dp_version == 2) tb2_walk_warp(tk, s, m.band, m.coff);
        else tb_walk_warp(tk, s, band, kC);
      }
      __syncthreads();
      long long t2 = clock64();
      if (tid == 0) {
        cyc[1] += static_cast<unsigned long long>(t1 - t0);
        cyc[2] += static_cast<unsigned long long>(t2 - t1);
        if (m.result[2] < 0 || m.result[0] <= 0) S.err = kWinTraceback;
        S.last_score = m.result[1];
        if (m.result[2] > 0) s_steps += static_cast<uint64_t>(m.result[2]);
        S.last_len = L;
      }
      __syncthreads();
      if (S.err != kWinOk) break;
      const int32_t np = m.result[2];
      if (P.pairs_out != nullptr && d.pairs_off >= 0) {   // debug: forward pairs of this sequence
        int32_t* dst = P.pairs_out + 2 * (d.pairs_off + static_cast<int64_t>(s_pairs));
        for (int32_t a = tid; a < np; a += T) {
          dst[2 * a] = m.path[2 * (np - 1 - a)];
          dst[2 * a + 1] = m.path[2 * (np - 1 - a) + 1];
        }
        __syncthreads();
        if (tid == 0) { P.pair_cnt[d.member_begin + q] = np; s_pairs += static_cast<uint64_t>(np); }
      }
      const uint32_t n_old = S.nv;
      dg_add_alignment(x, m, caps, &S, m.path, np, seq, L);
      long long t3 = clock64();
      if (S.err != kWinOk) break;
      dg_toposort(x, m, caps, &S, L, n_old);
      if (tid == 0) {
        const long long t4 = clock64();
        cyc[3] += static_cast<unsigned long long>(t3 - t2);
        cyc[4] += static_cast<unsigned long long>(t4 - t3);
      }
    }
    // ---- consensus, MSA, result record --------------------------------------------------------
    __syncthreads();
    const long long t5 = clock64();
    __shared__ uint32_t s_cons_len;
    __shared__ unsigned long long s_out_off;
    if (tid == 0) { s_cons_len = 0; s_out_off = 0; }
    __syncthreads();
    if (S.err == kWinOk && S.nseq > 0) {
      dg_export(x, m, caps, &S, s, static_cast<uint32_t>(P.ring_rows), true);
      uint8_t* cons_tmp = reinterpret_cast<uint8_t*>(m.band);
      const uint64_t V1 = static_cast<uint64_t>(caps.vcap) + 2;
      x.one([&]() { s_cons_len = dg_consensus_serial(m, &S, m.depth, m.depth + V1, cons_tmp); });
      uint32_t* head = m.single_before;
      uint32_t* col_of = reinterpret_cast<uint32_t*>(m.xslot);
      if (P.want_msa) dg_msa_columns(x, m, &S, head, col_of);
      x.one([&]() {
        const uint64_t bytes = (P.want_msa ? static_cast<uint64_t>(S.nseq) * S.msa_cols : 0) + s_cons_len;
        const unsigned long long need = dg_align(bytes, 16);
        const unsigned long long at = atomicAdd(P.out_cursor, need);
        s_out_off = at;
        if (at + need > P.out_cap) { S.err = kWinOutCap; s_need = need; }
      });
      if (S.err == kWinOk) {
        uint8_t* out = P.out_base + s_out_off;
        const uint64_t msa_bytes = P.want_msa ? static_cast<uint64_t>(S.nseq) * S.msa_cols : 0;
        if (P.want_msa) dg_msa_rows(x, m, &S, col_of, m.seq_len, S.nseq, out);
        for (uint32_t k = tid; k < s_cons_len; k += T) out[msa_bytes + k] = cons_tmp[k];
      }
    }
    __syncthreads();
    if (tid == 0) {
      WinResult r;
      r.status = S.err;
      r.msa_rows = (S.err == kWinOk && P.want_msa) ? S.nseq : 0;
      r.msa_cols = (S.err == kWinOk && P.want_msa) ? S.msa_cols : 0;
      r.cons_len = S.err == kWinOk ? s_cons_len : 0;
      r.out_off = s_out_off;
      r.n_align = s_nalign; r.retries = s_retries;
      r.nodes = S.nv; r.edges = S.ne;
      r.cells = s_cells; r.rows = s_rows; r.exported = s_exported; r.need_bytes = s_need;
      r.eval_cells = 8ull * s_eval[0];
      for (int k = 0; k < 4; ++k) r.warp_cyc[k] = s_eval[1 + k];
      r.read_bases = s_bases; r.path_steps = s_steps; r.pred_entries = s_preds;
      cyc[5] += static_cast<unsigned long long>(clock64() - t5);
      for (int k = 0; k < 8; ++k) r.cyc[k] = cyc[k];
      P.results[widx] = r;
    }
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    atomicExch(&P.slot_flags[s_slot], 0);
  }
}

}  // namespace

int poa_cols_per_thread(int threads, int cols) {
  if (cols == 16 && threads == 256) return 16;
  if (cols == 4 && threads == 512) return 4;
  return 8;
}

size_t poa_dp_smem_bytes(int threads, int ring_rows, int cols) {
  return static_cast<size_t>(ring_rows) * threads * poa_cols_per_thread(threads, cols) * sizeof(int32_t) +
         2 * (threads / 32) * sizeof(WarpPub) + sizeof(Stage);
}

int poa_dp_cols_per_pass(int threads, int cols) { return threads * poa_cols_per_thread(threads, cols); }

// ---- window kernel launchers ----------------------------------------------------------------
size_t poa_window_smem_bytes(int threads, int ring_rows, int cols) {
  const size_t v1 = poa_dp_smem_bytes(threads, ring_rows, cols);
  const size_t v2 = poa_cols_per_thread(threads, cols) == 8 ? dp2_smem_bytes(threads, ring_rows) : 0;
  return v1 > v2 ? v1 : v2;
}

int poa_window_ctas_per_sm(int threads, int ring_rows, int cols) {
  const size_t smem = poa_window_smem_bytes(threads, ring_rows, cols) + 2048;   // + static shared memory
  const int c = poa_cols_per_thread(threads, cols);
  if (smem > 227 * 1024) return 0;
  int by_smem = static_cast<int>((227 * 1024) / (smem + 1024));
  int cap = 1;
  if (threads == 128 && c == 8) cap = 4;
  else if (threads == 256 && c == 8) cap = 2;
  else if (threads == 256 && c == 16) cap = 1;
  else if (threads == 512) cap = 1;
  else return 0;
  return by_smem < cap ? by_smem : cap;
}

template <int T, int kC>
static cudaError_t window_cfg(int bytes) {
  return cudaFuncSetAttribute(poa_window_kernel<T, kC>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

cudaError_t poa_window_configure(int threads, int ring_rows, int cols) {
  if (poa_window_ctas_per_sm(threads, ring_rows, cols) <= 0) return cudaErrorInvalidValue;
  const int bytes = static_cast<int>(poa_window_smem_bytes(threads, ring_rows, cols));
  const int c = poa_cols_per_thread(threads, cols);
  if (threads == 128) return window_cfg<128, 8>(bytes);
  if (threads == 256 && c == 16) return window_cfg<256, 16>(bytes);
  if (threads == 256) return window_cfg<256, 8>(bytes);
  if (threads == 512 && c == 4) return window_cfg<512, 4>(bytes);
  if (threads == 512) return window_cfg<512, 8>(bytes);
  return cudaErrorInvalidValue;
}

cudaError_t poa_window_launch(const WinParams& p, int grid, int threads, int cols, cudaStream_t stream) {
  if (grid <= 0) return cudaSuccess;
  const size_t smem = poa_window_smem_bytes(threads, p.ring_rows, cols);
  const int c = poa_cols_per_thread(threads, cols);
  if (threads == 128) poa_window_kernel<128, 8><<<grid, 128, smem, stream>>>(p);
  else if (threads == 256 && c == 16) poa_window_kernel<256, 16><<<grid, 256, smem, stream>>>(p);
  else if (threads == 256) poa_window_kernel<256, 8><<<grid, 256, smem, stream>>>(p);
  else if (threads == 512 && c == 4) poa_window_kernel<512, 4><<<grid, 512, smem, stream>>>(p);
  else if (threads == 512) poa_window_kernel<512, 8><<<grid, 512, smem, stream>>>(p);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

}  // namespace svs
