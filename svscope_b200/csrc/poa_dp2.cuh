// Warp-pipelined dynamic programme of one sequence-to-graph alignment (included by
// poa_kernels.cu inside its anonymous namespace).
//
// Geometry: a pass covers a strip of T*8 read columns; warp w owns the 256 columns
// [jb + 256 w, jb + 256 (w + 1)), lane l the 8 columns from jb + 256 w + 8 l on ("chunk").
// Every warp sweeps the graph rows top to bottom ON ITS OWN: there is no CTA-wide barrier
// inside a pass.  The only coupling between warps is the horizontal dependency across the
// 256-column boundary: after the in-warp prefix scans of row i, lane 31 of warp w publishes the
// state of its last column (A, E, Q, H) in a 64-deep ring in shared memory and then row index i
// in prog[w]; warp w + 1 starts row i when prog[w] >= i, so the warps run as a skewed
// pipeline over the rows (anti-diagonal wavefront at warp granularity).  Warp w waits for
// warp w + 1 only when it is more than 32 rows ahead (ring reuse).  Predecessor rows come
// from registers (previous row), the warp's private ring of packed rows in shared memory
// (ring_rows deep, 1 KB per row) or, for rows with a far successor, global memory.
//
// Traceback codes: row i stores the chunks [clo_i, chi_i] its band touches, 1 byte per cell
// (rows with one predecessor) or 2 bytes per cell, at byte offset 8 * coff[i] (exclusive prefix
// sums, compute_bands2), so the scratch holds exactly the evaluated cells.
//
// Cell arithmetic, packed rows and code format: poa_cell.h (unchanged, proven against the
// five-matrix oracle).

// Profile builds (-DSVS_DP_PROFILE, scripts/dp_phase_profile.sh): lane 0 of every warp accumulates the cycles
// between marks in the row loop; the totals of every alignment of CTA 0 are printed.  Off in the product.
#ifdef SVS_DP_PROFILE
#define DP_MARK(k) do { if (lane == 0) { const long long t_now = clock64(); prof_acc[k] += t_now - t_mark; t_mark = t_now; } } while (0)
#else
#define DP_MARK(k) do { } while (0)
#endif

constexpr int kCarryDepth = 64;
constexpr int kStageCap = 160;   // predecessor entries staged per 32-row batch and warp

struct __align__(16) Carry {
  int32_t A, E, Q, H;
};

__host__ __device__ inline size_t dp2_smem_bytes(int threads, int ring_rows) {
  const int nw = threads / 32;
  return static_cast<size_t>(threads) * 8 * 4 * (ring_rows + 2)   // packed-row rings + the source row + a scratch row
         + static_cast<size_t>(nw) * kCarryDepth * sizeof(Carry)
         + 128                                                  // prog[], fprog[]
         + static_cast<size_t>(nw) * kStageCap * 12             // staged predecessor entries
         + static_cast<size_t>(nw) * 32 * 16;                   // row records of the running 32-row batch
}

struct __align__(16) TbRow {
  uint32_t first_flags;   // first stored column - 1 (a multiple of 8) | bit 0: one predecessor, bit 1: chain row, bit 2: has cells
  uint32_t coff;          // offset of the row's codes, units of 8 bytes
  uint32_t pred_off;      // offset of the row's predecessor list
  uint32_t pred0;         // first predecessor row
  uint32_t ncols;         // stored cells
  uint32_t node_id;
  uint32_t pad0, pad1;
};

// Band, chunk range and code offset of every row.  band[2i], band[2i+1] = first / last
// column (1-based; lo > hi: no cell); coff[i] = offset of the row's codes in units of 8 bytes.
// Returns (in *need_bytes, written by thread 0 to shared memory by the caller) the total.
template <int T>
__device__ void compute_bands2(CtaExec& x, const PoaTask& tk, const Scores& s, int32_t lb, bool have_lb, int32_t* band,
                               uint32_t* coff, TbRow* tbrows, unsigned long long* need_bytes) {
  const int32_t L = static_cast<int32_t>(tk.L);
  for (uint32_t i = threadIdx.x; i <= tk.R; i += T) {
    int32_t lo = 1, hi = L;
    if (i == 0) { lo = 1; hi = 0; }
    else if (have_lb) {
      const int4 d = *(reinterpret_cast<const int4*>(tk.depth) + i);
      auto ub = [&](int32_t j) { return cell_bound(s, d.x, d.y, d.z, d.w, j, L); };
      int32_t cand[6] = {1, L, d.x, d.y, L - d.w, L - d.z};
      int32_t jm = 1, best = INT32_MIN;
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const int32_t j = min(L, max(1, cand[k]));
        const int32_t v = ub(j);
        if (v > best) { best = v; jm = j; }
      }
      if (best < lb) {
        lo = 1; hi = 0;
      } else {
        int32_t a = 1, b = jm;
        while (a < b) { const int32_t mid = (a + b) >> 1; if (ub(mid) >= lb) b = mid; else a = mid + 1; }
        lo = a;
        a = jm; b = L;
        while (a < b) { const int32_t mid = (a + b + 1) >> 1; if (ub(mid) >= lb) a = mid; else b = mid - 1; }
        hi = a;
      }
    }
    band[2 * i] = lo;
    band[2 * i + 1] = hi;
    uint32_t units = 0;   // 8-byte units, rounded to 16 bytes
    if (lo <= hi) {
      const uint32_t chunks = static_cast<uint32_t>(((hi - 1) >> 3) - ((lo - 1) >> 3) + 1);
      const bool single = (tk.pred_off[i + 1] - tk.pred_off[i] == 1);
      units = single ? ((chunks + 1) & ~1u) : 2 * chunks;
    }
    coff[i + 1] = units;
  }
  if (threadIdx.x == 0) coff[0] = 0;
  __syncthreads();
  x.scan(coff + 1, tk.R + 1);     // coff[i+1] = units of rows 0..i  =>  coff[i] = offset of row i
  if (threadIdx.x == 0) *need_bytes = 8ull * coff[tk.R + 1] + 64;
  // one record per row for the traceback walk
  for (uint32_t i = threadIdx.x; i <= tk.R; i += T) {
    TbRow r;
    const int32_t lo = band[2 * i], hi = band[2 * i + 1];
    const bool cells = i > 0 && lo <= hi;
    const uint32_t po = tk.pred_off[i], npred = tk.pred_off[i + 1] - po;
    r.first_flags = (cells ? ((static_cast<uint32_t>(lo - 1) >> 3) << 3) : 0u) | (npred == 1 ? 1u : 0u) |
                    ((i > 0 && (tk.flags[i] & kFlagChain)) ? 2u : 0u) | (cells ? 4u : 0u);
    r.coff = coff[i];
    r.pred_off = po;
    r.pred0 = (i > 0 && npred > 0) ? tk.preds[po] : 0u;
    r.ncols = cells ? static_cast<uint32_t>((((hi - 1) >> 3) - ((lo - 1) >> 3) + 1) * 8) : 0u;
    r.node_id = tk.node_id[i];
    r.pad0 = 0; r.pad1 = 0;
    int4* dst = reinterpret_cast<int4*>(tbrows + i);
    dst[0] = make_int4(static_cast<int>(r.first_flags), static_cast<int>(r.coff), static_cast<int>(r.pred_off), static_cast<int>(r.pred0));
    dst[1] = make_int4(static_cast<int>(r.ncols), static_cast<int>(r.node_id), 0, 0);
  }
  __syncthreads();
}

__device__ __forceinline__ int ld_prog(const volatile int* p) { return *p; }

// Rare path of the predecessor fold: the row lives in global memory (it has a successor more than
// ring_rows rows away).  Copies my chunk into the warp's scratch row in shared memory, so that the
// fold itself has a single source, and returns lane 0's left neighbour.  Called by the whole warp.
__device__ __noinline__ int32_t dp2_stage_global(const PoaTask& tk, int32_t xslot, uint32_t j0, bool load_chunk, bool left_ok,
                                                 bool warp0, int32_t bh, const volatile int* fprog_left, int pbase_prog,
                                                 int32_t* scratch, long long* t_wait) {
  const int lane = threadIdx.x & 31;
  const int32_t* row = tk.xrows + static_cast<uint64_t>(xslot) * tk.ldx + 3;
  __syncwarp();   // the previous user of the scratch row is done
  if (load_chunk) {
    const int4 v0 = __ldcg(reinterpret_cast<const int4*>(row + j0));
    const int4 v1 = __ldcg(reinterpret_cast<const int4*>(row + j0 + 4));
    *reinterpret_cast<int4*>(scratch + 8 * lane) = v0;
    *reinterpret_cast<int4*>(scratch + 8 * lane + 4) = v1;
  }
  int32_t hl0 = warp0 ? bh : kNegBand;
  if (!warp0 && __shfl_sync(0xffffffffu, static_cast<int>(left_ok), 0)) {
    // lane 0's left column was written by the left warp at least ring_rows rows ago: wait for its fence
    if (lane == 0) {
      const long long t0 = clock64();
      while (*fprog_left < pbase_prog + bh) { }
      *t_wait += clock64() - t0;
      hl0 = unpack_h(__ldcg(row + j0 - 1));
    }
  }
  __syncwarp();
  return hl0;
}

// Source, chunk range and left-edge value of every predecessor of the lane's row, staged in the
// warp's shared-memory arrays for the row loop.  Returns whether the row reads anything the warp
// on the left produces.
__device__ __noinline__ bool dp2_stage_preds(const PoaTask& tk, const int32_t* band, const int32_t* bin, bool m_inter, uint32_t mi,
                                             uint32_t m_poff, uint32_t m_pend, uint32_t pbase, int32_t m_clo, int32_t wc0,
                                             int32_t sc0, int warp, uint32_t pass, int ring_rows, int32_t h_row0_left,
                                             int32_t* psrc, uint32_t* pchk, int32_t* pbh) {
  bool m_wait = m_inter && warp > 0 && m_clo < wc0;
  if (m_inter) {
    for (uint32_t e = m_poff; e < m_pend; ++e) {
      const uint32_t p = tk.preds[e];
      int32_t src, bh = kNegBand;
      uint32_t chk;
      if (p == 0) {
        src = ring_rows;          // the source-row slot
        chk = 0xffff0000u;
        bh = h_row0_left;
      } else {
        const int2 pb = *reinterpret_cast<const int2*>(band + 2 * p);
        const int32_t pclo = (pb.x - 1) >> 3, pchi = (pb.y - 1) >> 3;
        chk = pb.x <= pb.y ? (static_cast<uint32_t>(pclo) | (static_cast<uint32_t>(pchi) << 16)) : 1u;   // 1: lo = 1 > hi = 0
        if (mi - p <= static_cast<uint32_t>(ring_rows)) src = static_cast<int32_t>(p % ring_rows);
        else src = kSrcGlobal | tk.xslot[p];
        if (warp == 0) {
          if (pass == 0) bh = tk.h0[p];
          else if (pb.x <= pb.y && pclo <= sc0 - 1 && pchi >= sc0 - 1) bh = __ldcg(bin + p);
        } else {
          bh = static_cast<int32_t>(p);   // warps > 0 look the value up in the left warp's carry ring
          if (pb.x <= pb.y && pclo <= wc0 - 1 && pchi >= wc0 - 1) m_wait = true;
        }
      }
      psrc[e - pbase] = src;
      pchk[e - pbase] = chk;
      pbh[e - pbase] = bh;
    }
  }
  return m_wait;
}

template <int T>
__device__ __forceinline__ void dp2_align(const PoaTask& tk, const Scores& s, const int ring_rows,
                                          unsigned char* smem_raw, const int32_t* __restrict__ band,
                                          const uint32_t* __restrict__ coff, unsigned long long* eval_chunks) {
  constexpr int kC = 8;
  constexpr int NW = T / 32;
  const int32_t NEGW = pack_cell(kNegBand, kNeg, kNeg);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // entry `lane` of the table a successor looks up by the five low bits of a predecessor's packed word
  // (poa_cell.h): a look-up is one shuffle whose source lane is the packed word itself
  const PredLutEntry lut = pred_lut_entry(s, lane);
  // per warp: ring_rows packed rows (row r in slot r % ring_rows), slot ring_rows = the virtual source row
  // in my columns, slot ring_rows + 1 = scratch for a predecessor row fetched from global memory: every
  // predecessor is read from shared memory by ONE copy of the cell code
  int32_t* ring = reinterpret_cast<int32_t*>(smem_raw) + static_cast<size_t>(warp) * (ring_rows + 2) * 256;
  Carry* carry_all = reinterpret_cast<Carry*>(smem_raw + static_cast<size_t>(T) * 32 * (ring_rows + 2));
  volatile int* prog = reinterpret_cast<volatile int*>(carry_all + NW * kCarryDepth);
  // fprog[w]: rows of warp w whose GLOBAL stores (exported rows, strip boundary) are fenced; a fence
  // waits for every outstanding store of the warp (the traceback codes too), so it is issued every 8
  // rows, not per row: the per-row hand-over between neighbouring warps is shared memory only
  volatile int* fprog = prog + 16;
  int32_t* stage = reinterpret_cast<int32_t*>(const_cast<int*>(prog) + 32) + warp * (kStageCap * 3);
  // one 16-byte record per row of the running batch: chunk range, letter | flags | ring slot | need-left,
  // predecessor entries, code offset; the row loop gets a row's metadata with one broadcast load
  int4* mrec = reinterpret_cast<int4*>(reinterpret_cast<int32_t*>(const_cast<int*>(prog) + 32) + NW * (kStageCap * 3)) + warp * 32;
  int32_t* psrc = stage;                                        // source of the predecessor row
  uint32_t* pchk = reinterpret_cast<uint32_t*>(stage + kStageCap);   // its chunk range, lo | hi << 16
  int32_t* pbh = stage + 2 * kStageCap;                         // warp 0: H of the predecessor left of the strip
  Carry* carry_mine = carry_all + warp * kCarryDepth;
  const Carry* carry_left = carry_all + (warp > 0 ? warp - 1 : 0) * kCarryDepth;

  const uint32_t R = tk.R, L = tk.L;
  const uint64_t bstride = static_cast<uint64_t>(R) + 1;
  int32_t best = INT32_MIN;
  uint32_t best_row = 0;
  uint32_t n_chunks = 0;   // evaluated 8-cell chunks of this thread

  // progress of a warp = pass * (R + 1) + last row done: the passes of different warps overlap
  // (warp 0 starts strip p + 1 while the other warps are still in strip p); the strip boundary
  // travels through global memory (bnd), guarded by the progress of the warp that owns the last
  // chunk of the strip
  if (lane == 0) { prog[warp] = 0; fprog[warp] = 0; }
  __syncthreads();
  int next_fence = 0;
  bool unfenced = false;   // rows processed since my last fence
  const int wlast = static_cast<int>(((tk.strip >> 3) - 1) >> 5);   // warp that owns the last chunk of a full strip
  long long t_wait_left = 0;   // cycles lane 0 spent polling for the state its left neighbour hands over (per warp)
  long long t_wait_pipe = 0;   // ... and polling at the start of a row: left neighbour one row behind, right neighbour 32 rows behind
  const long long t_begin = clock64();
#ifdef SVS_DP_PROFILE
  __shared__ long long prof_all[16 * 10];
  long long* prof_acc = prof_all + warp * 10;   // 0 batch staging, 1 early polls, 2 fold, 3 pre-scan, 4 wait for the left carry, 5 scan + publish, 6 finish, 7 skipped batches, 8 rows, 9 in-edges
  if (lane == 0) for (int k = 0; k < 10; ++k) prof_acc[k] = 0;
  long long t_mark = t_begin;
#endif
  int next_check = 40;   // (absolute progress) next row at which I make sure not to lap the consumer of my carry ring

  for (uint32_t pass = 0; pass < tk.npass; ++pass) {
    const int pbase_prog = static_cast<int>(pass * (R + 1));
    const uint32_t jb = 1 + pass * tk.strip;
    const uint32_t je = min(L, jb + tk.strip - 1);
    const uint32_t j0 = jb + 256u * warp + kC * lane;
    const bool active = j0 <= je;
    const int32_t gc = static_cast<int32_t>((j0 - 1) >> 3);                 // my chunk
    const int32_t wc0 = static_cast<int32_t>((jb - 1) >> 3) + 32 * warp;    // first chunk of my warp
    const int32_t wc1 = min(wc0 + 31, static_cast<int32_t>((je - 1) >> 3)); // last chunk of my warp inside the strip
    const int32_t sc0 = static_cast<int32_t>((jb - 1) >> 3);                // first chunk of the strip
    const bool last_pass = (pass + 1 == tk.npass);
    const bool owns_end = last_pass && active && (L < j0 + kC);
    const int c_end = owns_end ? static_cast<int>(L - j0) : -1;
    const bool writes_bnd = !last_pass && active && (gc == static_cast<int32_t>((je - 1) >> 3));
    const int32_t* bin = tk.bnd + static_cast<uint64_t>(pass & 1) * 4 * bstride;
    int32_t* bout = tk.bnd + static_cast<uint64_t>((pass + 1) & 1) * 4 * bstride;

    int32_t rd[kC];
#pragma unroll
    for (int c = 0; c < kC; ++c) {
      const uint32_t j = j0 + c;
      rd[c] = (active && j <= L) ? static_cast<int32_t>(tk.read[j - 1]) : 0x100;
    }
    {   // the virtual source row in my columns (H of a leading gap; F, O = -inf)
      int32_t* r0 = ring + static_cast<size_t>(ring_rows) * 256 + kC * lane;
#pragma unroll
      for (int c = 0; c < kC; ++c) r0[c] = pack_cell(row0_h(s, static_cast<int32_t>(j0) + c), kNeg, kNeg);
      __syncwarp();
    }
    const int32_t h_row0_left = row0_h(s, static_cast<int32_t>(jb + 256u * warp) - 1);   // source row, column left of my warp

    uint32_t i0 = 1;
    while (i0 <= R) {
      // ---- metadata of up to 32 rows, one row per lane ---------------------------------------
      uint32_t nrows = min(32u, R - i0 + 1);
      const uint32_t mi = i0 + lane;
      const bool mvalid = static_cast<uint32_t>(lane) < nrows;
      int32_t m_lo = 1, m_hi = 0;
      uint32_t m_poff = 0, m_pend = 0, m_coff = 0, m_info = 0;
      int32_t m_bA = kNegBand, m_bE = kNeg, m_bQ = kNeg;
      if (mvalid) {
        const int2 b = *reinterpret_cast<const int2*>(band + 2 * mi);
        m_lo = b.x; m_hi = b.y;
        m_poff = tk.pred_off[mi];
        m_pend = tk.pred_off[mi + 1];
        m_coff = coff[mi];
        m_info = static_cast<uint32_t>(tk.letter[mi]) | (static_cast<uint32_t>(tk.flags[mi]) << 8);
      }
      const uint32_t pbase = __shfl_sync(0xffffffffu, m_poff, 0);
      {   // keep the staged predecessor entries of the batch within kStageCap
        const unsigned fits = __ballot_sync(0xffffffffu, mvalid && (m_pend - pbase) <= static_cast<uint32_t>(kStageCap));
        nrows = min(nrows, static_cast<uint32_t>(__popc(fits)));   // m_pend is increasing: a prefix of the lanes
      }
      const bool mine = static_cast<uint32_t>(lane) < nrows;
      const int32_t m_clo = (m_lo - 1) >> 3, m_chi = (m_hi - 1) >> 3;
      const bool m_inter = mine && m_lo <= m_hi && m_chi >= wc0 && m_clo <= wc1;
      const unsigned any = __ballot_sync(0xffffffffu, m_inter);
      if (any == 0) {   // no row of the batch has a cell in my 256 columns
        if (unfenced) { __threadfence_block(); unfenced = false; }
        if (lane == 31) { prog[warp] = pbase_prog + static_cast<int>(i0 + nrows - 1); fprog[warp] = pbase_prog + static_cast<int>(i0 + nrows - 1); }
        i0 += nrows;
        DP_MARK(7);
        continue;
      }
      // left boundary of the strip (warp 0): column 0 in the first strip, else the state the
      // previous strip left behind if the band covered its last chunk
      if (warp == 0 && pass > 0) {   // the previous strip must have passed the rows of this batch
        const int need = static_cast<int>((pass - 1) * (R + 1) + i0 + nrows - 1);
        if (lane == 0) { const long long t0 = clock64(); while (ld_prog(fprog + wlast) < need) { } t_wait_left += clock64() - t0; }
        __syncwarp();
        asm volatile("" ::: "memory");
      }
      if (warp == 0 && mine) {
        if (pass == 0) {
          m_bA = tk.h0[mi];
        } else if (m_lo <= m_hi && m_clo <= sc0 - 1 && m_chi >= sc0 - 1) {
          m_bA = __ldcg(bin + bstride + mi);
          m_bE = __ldcg(bin + 2 * bstride + mi);
          m_bQ = __ldcg(bin + 3 * bstride + mi);
        }
      }
      // ---- stage the predecessor entries of my row (once per 32 rows: kept out of the row loop's code) ----
      __syncwarp();
      const bool m_wait = dp2_stage_preds(tk, band, bin, m_inter, mi, m_poff, m_pend, pbase, m_clo, wc0, sc0, warp, pass, ring_rows,
                                          h_row0_left, psrc, pchk, pbh);
      if (m_inter) {   // only rows with a cell in my columns are visited below
        mrec[lane] = make_int4(static_cast<int>(static_cast<uint32_t>(m_clo) | (static_cast<uint32_t>(m_chi) << 16)),
                               static_cast<int>(m_info | ((mi % static_cast<uint32_t>(ring_rows)) << 16) | (m_wait ? 1u << 24 : 0u)),
                               static_cast<int>((m_poff - pbase) | ((m_pend - pbase) << 16)),
                               static_cast<int>(m_coff));
      }
      __syncwarp();
      DP_MARK(0);

      for (uint32_t r = 0; r < nrows; ++r) {
        if (!((any >> r) & 1u)) continue;
        const uint32_t i = i0 + r;
        const int4 rec = mrec[r];
        const int32_t clo = static_cast<int32_t>(static_cast<uint32_t>(rec.x) & 0xffffu);
        const int32_t chi = static_cast<int32_t>(static_cast<uint32_t>(rec.x) >> 16);
        const uint32_t info = static_cast<uint32_t>(rec.y);
        const uint32_t nb = static_cast<uint32_t>(rec.z) & 0xffffu;
        const uint32_t ne = static_cast<uint32_t>(rec.z) >> 16;
        const uint32_t row_coff = static_cast<uint32_t>(rec.w);
        const bool need_left = (info >> 24) != 0;
        const int32_t letter = static_cast<int32_t>(info & 0xffu);
        const uint32_t rflags = (info >> 8) & 0xffu;
        const uint32_t rslot = (info >> 16) & 0xffu;   // i % ring_rows
        const bool single = (ne - nb == 1);
        const bool t_active = active && gc >= clo && gc <= chi;
        __syncwarp();   // the packed row my neighbours stored last is visible

        // ---- pipeline control -----------------------------------------------------------------
        Carry cin;
        cin.A = kNegBand; cin.E = kNeg; cin.Q = kNeg; cin.H = kNegBand;
        if (warp == 0) {
          cin.A = __shfl_sync(0xffffffffu, m_bA, r);
          cin.E = __shfl_sync(0xffffffffu, m_bE, r);
          cin.Q = __shfl_sync(0xffffffffu, m_bQ, r);
        } else if (need_left) {   // predecessor rows left of my span: the left warp must have done row i - 1
          if (lane == 0) { const long long t0 = clock64(); while (ld_prog(prog + warp - 1) < pbase_prog + static_cast<int>(i) - 1) { } t_wait_pipe += clock64() - t0; }
          __syncwarp();
          asm volatile("" ::: "memory");
        }
        const int abs_i = pbase_prog + static_cast<int>(i);   // carry-ring slots are indexed by absolute progress
        if (NW > 1 && warp + 1 < NW && abs_i >= next_check) {
          if (lane == 0) { const long long t0 = clock64(); while (ld_prog(prog + warp + 1) < abs_i - 32) { } t_wait_pipe += clock64() - t0; }
          __syncwarp();
          next_check = abs_i + 8;
        }
        DP_MARK(1);
#ifdef SVS_DP_PROFILE
        if (lane == 0) { prof_acc[8] += 1; prof_acc[9] += ne - nb; }
#endif

        // ---- phase 1: fold predecessor rows ---------------------------------------------------
        // Executed by every lane of the warp (the table look-ups are shuffles); lanes without a cell
        // in this row fold band-edge values that nothing reads.
        CellAcc acc[kC];
        if (!single) {
#pragma unroll
          for (int c = 0; c < kC; ++c) cell_key_init(acc[c]);
        }
        for (uint32_t e = nb; e < ne; ++e) {
          const int32_t src = psrc[e];
          const uint32_t chk = pchk[e];
          const int32_t pclo = static_cast<int32_t>(chk & 0xffffu), pchi = static_cast<int32_t>(chk >> 16);
          const bool chunk_ok = gc >= pclo && gc <= pchi;
          const bool left_ok = gc - 1 >= pclo && gc - 1 <= pchi;
          int32_t hl0;   // lane 0: H of the predecessor row in the column left of my warp (or of the strip)
          const int32_t* row;
          if (src & kSrcGlobal) {   // rare: a row with a far successor, kept in global memory
            hl0 = dp2_stage_global(tk, src & ~kSrcGlobal, j0, chunk_ok && active, left_ok, warp == 0, pbh[e],
                                   fprog + (warp > 0 ? warp - 1 : 0), pbase_prog, ring + static_cast<size_t>(ring_rows + 1) * 256,
                                   &t_wait_left);
            row = ring + static_cast<size_t>(ring_rows + 1) * 256;
          } else {
            row = ring + static_cast<size_t>(src) * 256;
            const int32_t bh = pbh[e];
            hl0 = (warp == 0 || src == ring_rows) ? bh
                                                   : (left_ok ? carry_left[static_cast<uint32_t>(pbase_prog + bh) & (kCarryDepth - 1)].H : kNegBand);
          }
          int32_t w[kC];
          {
            const int4 v0 = *reinterpret_cast<const int4*>(row + kC * lane);
            const int4 v1 = *reinterpret_cast<const int4*>(row + kC * lane + 4);
            w[0] = v0.x; w[1] = v0.y; w[2] = v0.z; w[3] = v0.w; w[4] = v1.x; w[5] = v1.y; w[6] = v1.z; w[7] = v1.w;
          }
          if (!chunk_ok) {
#pragma unroll
            for (int c = 0; c < kC; ++c) w[c] = NEGW;
          }
          // packed word of the column on my left (only its H matters)
          int32_t wl = (lane == 0) ? hl0 * 32 : (left_ok ? row[kC * lane - 1] : NEGW);
          if (single) {   // one predecessor: no argmax bookkeeping
#pragma unroll
            for (int c = 0; c < kC; ++c) {
              const int32_t sfv = __shfl_sync(0xffffffffu, lut.sf, w[c]);
              const int32_t sov = __shfl_sync(0xffffffffu, lut.so, w[c]);
              const int32_t smv = __shfl_sync(0xffffffffu, lut.sm, w[c]);
              cell_pred_single(acc[c], w[c], unpack_h(wl), (letter == rd[c]) ? s.m : s.n, sfv, sov, smv);
              wl = w[c];
            }
          } else {
            const int32_t rk = 31 - static_cast<int32_t>(e - nb);
#pragma unroll
            for (int c = 0; c < kC; ++c) {
              const int32_t tfv = __shfl_sync(0xffffffffu, lut.tf, w[c]);
              const int32_t tov = __shfl_sync(0xffffffffu, lut.to, w[c]);
              const int32_t tvv = __shfl_sync(0xffffffffu, lut.tv, w[c]);
              cell_pred_key(acc[c], rk, w[c], wl, tfv, tov, tvv);
              wl = w[c];
            }
          }
        }
        if (!single) {
#pragma unroll
          for (int c = 0; c < kC; ++c) cell_key_add_sub(acc[c], (letter == rd[c]) ? s.m : s.n);
        }

        DP_MARK(2);
        // ---- scan: horizontal gap states across the 256 columns of the warp ---------------------
        int32_t a7 = kNegBand;
        int32_t el = kNeg, ql = kNeg, eloc7 = kNeg, qloc7 = kNeg;
        if (t_active) {
#pragma unroll
          for (int c = 0; c < kC; ++c) {
            const int32_t A = single ? imax(acc[c].D, imax(acc[c].Fm, acc[c].Om))
                                     : imax(key_value_diag(acc[c].D), key_value(static_cast<int32_t>(acc[c].meta)));
            if (c == kC - 1) { eloc7 = el; qloc7 = ql; a7 = A; }
            el = imax(A + s.g, el + s.e);
            ql = imax(A + s.q, ql + s.c);
          }
        }
        DP_MARK(3);
        if (warp > 0 && need_left) {   // the scan needs the left warp's state of THIS row
          if (lane == 0) { const long long t0 = clock64(); while (ld_prog(prog + warp - 1) < pbase_prog + static_cast<int>(i)) { } t_wait_left += clock64() - t0; }
          __syncwarp();
          asm volatile("" ::: "memory");
          if (clo < wc0) cin = carry_left[abs_i & (kCarryDepth - 1)];
        }
        DP_MARK(4);
        int32_t ein0 = 0, qin0 = 0;
        if (lane == 0) {   // the state left of the warp enters through lane 0
          ein0 = imax(cin.A + s.g, cin.E + s.e);
          qin0 = imax(cin.A + s.q, cin.Q + s.c);
          el = imax(el, ein0 + kC * s.e);
          ql = imax(ql, qin0 + kC * s.c);
        }
        int32_t ve = el, vq = ql;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          // lanes below d get their own value back from the shuffle: own + (negative gap extension) never wins
          const int32_t oe = __shfl_up_sync(0xffffffffu, ve, d);
          const int32_t oq = __shfl_up_sync(0xffffffffu, vq, d);
          ve = imax(ve, oe + kC * s.e * d);
          vq = imax(vq, oq + kC * s.c * d);
        }
        int32_t ein = __shfl_up_sync(0xffffffffu, ve, 1);
        int32_t qin = __shfl_up_sync(0xffffffffu, vq, 1);
        if (lane == 0) { ein = ein0; qin = qin0; }
        const int32_t se = imax(ein + (kC - 1) * s.e, eloc7);   // E, Q at my last column
        const int32_t sq = imax(qin + (kC - 1) * s.c, qloc7);
        RowCarry cy;
        cy.A = __shfl_up_sync(0xffffffffu, a7, 1);
        cy.E = __shfl_up_sync(0xffffffffu, se, 1);
        cy.Q = __shfl_up_sync(0xffffffffu, sq, 1);
        if (lane == 0) { cy.A = cin.A; cy.E = cin.E; cy.Q = cin.Q; }
        cy.H = imax(cy.A, imax(cy.E, cy.Q));

        // ---- publish the state of my warp's last column (or of the strip's last column, through
        //      global memory, for the next strip), then the row index ---------------------------------
        if (writes_bnd && t_active) {
          bout[i] = imax(a7, imax(se, sq));
          bout[bstride + i] = a7;
          bout[2 * bstride + i] = se;
          bout[3 * bstride + i] = sq;
        }
        if (abs_i >= next_fence) {   // rows < i (and the boundary of row i) are now visible to the whole CTA
          __threadfence_block();
          if (lane == 31) fprog[warp] = abs_i - 1;
          next_fence = abs_i + 8;
        }
        unfenced = true;
        if (lane == 31) {
          if (warp + 1 < NW) {
            Carry out;
            out.A = a7; out.E = se; out.Q = sq; out.H = imax(a7, imax(se, sq));
            carry_mine[abs_i & (kCarryDepth - 1)] = out;
          }
          asm volatile("" ::: "memory");   // same-thread shared-memory stores are performed in order
          prog[warp] = abs_i;
        }

        DP_MARK(5);
        // ---- phase 2: H, traceback codes, packed row ---------------------------------------------
        if (t_active) {
          ++n_chunks;
          uint8_t* crow = tk.codes + 8ull * row_coff;
          const uint32_t cidx = static_cast<uint32_t>(gc - clo);
          int32_t* rrow = ring + rslot * 256 + kC * lane;
          int32_t* xrow = nullptr;
          if (rflags & kFlagExport) xrow = tk.xrows + static_cast<uint64_t>(tk.xslot[i]) * tk.ldx + 3 + j0;
          // two halves of four cells through ONE copy of the cell code (instruction-cache footprint):
          // the second half's accumulators move into the first half's registers
#pragma unroll 1
          for (int h = 0; h < 2; ++h) {
            uint32_t cw0 = 0, cw1 = 0;
            int32_t wp[4];
            if (single) {
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                int32_t H;
                const uint32_t cd = cell_finish_single(acc[c], cy, s, H);
                wp[c] = pack_cell(H, acc[c].Fm, acc[c].Om);
                if (c == 0) cw0 = cd; else if (c == 1) cw0 |= cd << 16; else if (c == 2) cw1 = cd; else cw1 |= cd << 16;
              }
            } else {
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                int32_t H, Fv, Ov;
                const uint32_t cd = cell_finish_key(acc[c], cy, s, H, Fv, Ov);
                wp[c] = pack_cell(H, Fv, Ov);
                if (c == 0) cw0 = cd; else if (c == 1) cw0 |= cd << 16; else if (c == 2) cw1 = cd; else cw1 |= cd << 16;
              }
            }
            if (single) {   // low bytes only
              const uint32_t b = (cw0 & 0xffu) | ((cw0 >> 8) & 0xff00u) | ((cw1 & 0xffu) << 16) | ((cw1 & 0xff0000u) << 8);
              *reinterpret_cast<uint32_t*>(crow + 8ull * cidx + 4 * h) = b;
            } else {
              *reinterpret_cast<uint2*>(crow + 16ull * cidx + 8 * h) = make_uint2(cw0, cw1);
            }
            *reinterpret_cast<int4*>(rrow + 4 * h) = make_int4(wp[0], wp[1], wp[2], wp[3]);
            if (xrow != nullptr) *reinterpret_cast<int4*>(xrow + 4 * h) = make_int4(wp[0], wp[1], wp[2], wp[3]);
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[c] = acc[c + 4];
          }
          if (owns_end && (rflags & kFlagSink)) {   // H of the last read column: my own packed cell, just stored
            const int32_t hsel = unpack_h(rrow[c_end]);
            if (hsel > best) { best = hsel; best_row = i; }
          }
        }
        DP_MARK(6);
      }
      // End of the batch: rows at its end that were skipped count as done, and everything the warp stored in
      // the batch is fenced and published in fprog as well.  Without this a warp that visits no row for more
      // than 32 rows after an exported row p (a pruned or far-away branch in between) would wait for its right
      // neighbour before its next fence, while that neighbour waits for the fence of p (dp2_stage_global): a
      // cyclic wait.  With it fprog[w] >= the end of warp w's last finished batch; when warp w waits for warp
      // w + 1 (more than 32 rows behind), w + 1 stands in an earlier batch (batch boundaries depend on the graph
      // only), so every row it can ask for is fenced.
      if (unfenced) { __threadfence_block(); unfenced = false; }
      __syncwarp();
      if (lane == 31) {
        fprog[warp] = pbase_prog + static_cast<int>(i0 + nrows - 1);
        prog[warp] = pbase_prog + static_cast<int>(i0 + nrows - 1);
      }
      i0 += nrows;
    }
    if (owns_end) {
      tk.result[0] = static_cast<int32_t>(best_row);
      tk.result[1] = best;
    }
  }
  const long long t_done = clock64();
  __syncthreads();
#ifdef SVS_DP_PROFILE
  if (blockIdx.x == 0 && lane == 0) {
    printf("dpprof R %u L %u warp %d total %lld : stage %lld poll0 %lld fold %lld prescan %lld waitleft %lld scan %lld finish %lld skipped %lld | rows %lld inedges %lld\n",
           tk.R, tk.L, warp, t_done - t_begin, prof_acc[0], prof_acc[1], prof_acc[2], prof_acc[3], prof_acc[4], prof_acc[5], prof_acc[6],
           prof_acc[7], prof_acc[8], prof_acc[9]);
  }
  __syncthreads();
#endif
  n_chunks = __reduce_add_sync(0xffffffffu, n_chunks);
  if (lane == 0 && eval_chunks != nullptr) {
    atomicAdd(eval_chunks, static_cast<unsigned long long>(n_chunks));
    atomicAdd(eval_chunks + 1, static_cast<unsigned long long>(t_done - t_begin));        // busy + polling, per warp
    atomicAdd(eval_chunks + 2, static_cast<unsigned long long>(t_wait_left));
    atomicAdd(eval_chunks + 3, static_cast<unsigned long long>(t_wait_pipe));
    atomicAdd(eval_chunks + 4, static_cast<unsigned long long>(clock64() - t_done));        // waiting for the slowest warp at the end
  }
}

// ---- traceback ---------------------------------------------------------------------------------
// The walk is latency-bound: every step needs the row's code layout, the code itself and the
// predecessor row.  All a step needs about a row sits in one 32-byte record (written by
// compute_bands2), including the row's FIRST predecessor, so that
//   * the usual step (move through in-edge 0) needs no look-up in the predecessor list, and
//   * the record of that predecessor is requested before the code of the current cell arrives:
//     in steady state a step costs one dependent global load (the code) instead of three.
__device__ __forceinline__ TbRow tb_load(const TbRow* rows, uint32_t i) {
  const int4 a = __ldcg(reinterpret_cast<const int4*>(rows + i));
  const int4 b = __ldcg(reinterpret_cast<const int4*>(rows + i) + 1);
  TbRow r;
  r.first_flags = static_cast<uint32_t>(a.x); r.coff = static_cast<uint32_t>(a.y);
  r.pred_off = static_cast<uint32_t>(a.z); r.pred0 = static_cast<uint32_t>(a.w);
  r.ncols = static_cast<uint32_t>(b.x); r.node_id = static_cast<uint32_t>(b.y); r.pad0 = 0; r.pad1 = 0;
  return r;
}

struct TbCtx {
  const uint8_t* codes;
  const TbRow* rows;
  const uint16_t* col0code;
  const uint32_t* preds;
};

__device__ __forceinline__ uint32_t tb3_code_at(const TbCtx& v, const Scores& s, const TbRow& R, uint32_t ii, uint32_t jj) {
  if (ii == 0) return jj == 0 ? 0u : row0_code(s, static_cast<int32_t>(jj));
  if (jj == 0) return v.col0code[ii];
  const uint8_t* row = v.codes + 8ull * R.coff;
  const uint32_t col = jj - 1 - (R.first_flags & ~7u);
  if (R.first_flags & 1u) return code_of_single_byte(row[col]);
  return reinterpret_cast<const uint16_t*>(row)[col];
}

__device__ __forceinline__ uint32_t tb3_pred_row(const TbCtx& v, const TbRow& R, uint32_t k) {
  if (k == kNoPred) return 0u;
  if (k == 0) return R.pred0;
  return v.preds[R.pred_off + k];
}

// Moves the walk from row `i` (record cur, speculative record spec of cur.pred0) to row `to`.
__device__ __forceinline__ void tb3_goto(const TbCtx& v, uint32_t& i, TbRow& cur, TbRow& spec, uint32_t to) {
  if (to == i) return;
  if (to == 0) { i = 0; return; }
  cur = (to == cur.pred0) ? spec : tb_load(v.rows, to);
  i = to;
  if (cur.pred0 != 0) spec = tb_load(v.rows, cur.pred0);   // requested now, needed at the next row change
}

// One iteration of the reference's traceback loop at (i, j) != (0, 0): same decisions as poa_cell.h tb_step.
__device__ bool tb3_step(const TbCtx& v, const Scores& s, uint32_t& i, uint32_t& j, TbRow& cur, TbRow& spec, int32_t& n,
                         int32_t* out_pairs, int32_t cap) {
  const uint32_t cd = tb3_code_at(v, s, cur, i, j);
  const uint32_t move = cd & 3, ext = (cd >> 2) & 1, km = code_kmove(cd);
  uint32_t pi, pj;
  if (move == kMoveDiag) { pi = tb3_pred_row(v, cur, km); pj = j - 1; }
  else if (move == kMoveVert) { pi = tb3_pred_row(v, cur, km); pj = j; }
  else { pi = i; pj = j - 1; }
  if (n >= cap) return false;
  out_pairs[2 * n] = (i == pi) ? -1 : static_cast<int32_t>(cur.node_id);
  out_pairs[2 * n + 1] = (j == pj) ? -1 : static_cast<int32_t>(j - 1);
  ++n;
  tb3_goto(v, i, cur, spec, pi);
  j = pj;
  if (move == kMoveHorz && ext) {
    while (true) {
      if (n >= cap) return false;
      out_pairs[2 * n] = -1;
      out_pairs[2 * n + 1] = static_cast<int32_t>(j - 1);
      ++n;
      --j;
      if (j == 0 || !((tb3_code_at(v, s, cur, i, j) >> 3) & 1)) break;
    }
  } else if (move == kMoveVert && ext) {
    while (i != 0) {
      const uint32_t c2 = tb3_code_at(v, s, cur, i, j);
      const uint32_t stop = code_stop(c2), ku = code_kup(c2);
      const uint32_t up = tb3_pred_row(v, cur, ku);
      if (n >= cap) return false;
      out_pairs[2 * n] = static_cast<int32_t>(cur.node_id);
      out_pairs[2 * n + 1] = -1;
      ++n;
      tb3_goto(v, i, cur, spec, up);
      if (stop || i == 0) break;
    }
  }
  return true;
}

// Traceback by one warp.  While the walk is on a chain row the 32 lanes test the cells (i-k, j-k)
// for "chain row + diagonal move" and the walk advances by the number of leading hits; anything
// else is one serial step (every lane executes it on identical values, lane 0 stores).
__device__ void tb3_walk_warp(const PoaTask& tk, const Scores& s, const TbRow* rows) {
  const TbCtx v{tk.codes, rows, tk.col0code, tk.preds};
  const int lane = threadIdx.x & 31;
  uint32_t i = static_cast<uint32_t>(tk.result[0]), j = tk.L;
  int32_t n = 0;
  const int32_t cap = static_cast<int32_t>(tk.path_cap);
  bool ok = true;
  TbRow cur = tb_load(rows, i), spec = cur;
  if (i != 0 && cur.pred0 != 0) spec = tb_load(rows, cur.pred0);
  int32_t* path = tk.path;
  while (ok && !(i == 0 && j == 0)) {
    int run = 0;
    if (i != 0 && (cur.first_flags & 2u)) {   // chain row: try a diagonal run
      bool mine = false;
      int32_t node = 0;
      if (i > static_cast<uint32_t>(lane) && j > static_cast<uint32_t>(lane)) {
        const uint32_t r = i - lane, c = j - lane;
        const TbRow R = lane == 0 ? cur : tb_load(rows, r);
        const uint32_t first = R.first_flags & ~7u;
        const bool inside = (R.first_flags & 4u) && c - 1 >= first && c - 1 < first + R.ncols;
        if (inside && (R.first_flags & 2u)) {
          const uint32_t cd = tk.codes[8ull * R.coff + (c - 1 - first)];
          if ((cd & 3u) == kMoveDiag) {
            mine = true;
            node = static_cast<int32_t>(R.node_id);
          }
        }
      }
      const unsigned hit = __ballot_sync(0xffffffffu, mine);
      const int m = __ffs(~hit) - 1;
      run = m < 0 ? 32 : m;
      if (run > 0) {
        if (n + run > cap) { ok = false; break; }
        if (lane < run) {
          path[2 * (n + lane)] = node;
          path[2 * (n + lane) + 1] = static_cast<int32_t>(j - lane - 1);
        }
        n += run;
        i -= run;
        j -= run;
        if (i != 0) {
          cur = tb_load(rows, i);
          if (cur.pred0 != 0) spec = tb_load(rows, cur.pred0);
        }
      }
    }
    if (run == 0) {
      // every lane walks the same step on the same values (uniform loads); only lane 0 stores
      ok = tb3_step(v, s, i, j, cur, spec, n, path, cap);
    }
  }
  if (lane == 0) tk.result[2] = ok ? n : -1;
}
