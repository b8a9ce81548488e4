// CPU emulation of the MisScore kernel (test infrastructure): the same per-thread row passes
// (misscore_cell.h), strip plan, nibble layout and traceback (misscore_tb.h) as
// svscope_b200/csrc/misscore.cu, with the CTA's threads run one after the other.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../svscope_b200/csrc/misscore_cell.h"

using namespace svs;

template <int kC>
static int run(const uint8_t* a, int la, const uint8_t* b, int lb, int s_match, int s_mis, int gap, int threads,
               int32_t* out, uint8_t* line) {
  const int64_t pitch = mis_trace_pitch(lb);
  std::vector<uint8_t> trace(static_cast<size_t>(pitch) * la, 0);
  const int strip = mis_strip_cols(lb, threads, kC);
  const int nstrips = (lb + strip - 1) / strip;
  std::vector<int32_t> bnd[2];
  bnd[0].assign(la + 1, 0);
  bnd[1].assign(la + 1, 0);
  struct Th { int up[kC]; int upleft; int loc[kC]; uint32_t eq; };
  std::vector<Th> th(threads);
  std::vector<int> x(threads);
  int score = 0;
  for (int s = 0; s < nstrips; ++s) {
    const int col0 = s * strip;
    const int32_t* bin = bnd[(s & 1) ^ 1].data();
    int32_t* bout = bnd[s & 1].data();
    const int nthr = (strip + kC - 1) / kC;
    for (int t = 0; t < nthr; ++t) {
      const int j0 = col0 + t * kC + 1;
      for (int k = 0; k < kC; ++k) th[t].up[k] = -gap * (j0 + k);
      th[t].upleft = -gap * (j0 - 1);
    }
    for (int r = 1; r <= la; ++r) {
      const int lin0 = s == 0 ? -gap * r : bin[r];
      for (int t = 0; t < nthr; ++t) {
        const int j0 = col0 + t * kC + 1;
        uint32_t eq = 0;
        for (int k = 0; k < kC; ++k)
          if (j0 + k <= lb && a[r - 1] == b[j0 + k - 1]) eq |= 1u << k;
        th[t].eq = eq;
        const int p = mis_pass1<kC>(th[t].up, th[t].upleft, eq, s_match, s_mis, gap, t == 0 ? lin0 : kMisNeg, th[t].loc);
        x[t] = p + gap * kC * (t + 1);
      }
      int run_max = kMisNeg;
      for (int t = 0; t < nthr; ++t) {
        const int lin = t == 0 ? lin0 : run_max - gap * kC * t;
        run_max = x[t] > run_max ? x[t] : run_max;
        const int j0 = col0 + t * kC + 1;
        const uint64_t nibs = mis_pass2<kC>(th[t].up, th[t].upleft, th[t].eq, s_match, s_mis, gap, lin, th[t].loc);
        if (j0 <= lb) {
          // kC nibbles = kC/2 bytes at byte (j0-1)/2 of the row
          uint8_t* dst = trace.data() + static_cast<int64_t>(r - 1) * pitch + (j0 - 1) / 2;
          for (int k = 0; k < kC / 2; ++k) dst[k] = static_cast<uint8_t>(nibs >> (8 * k));
        }
        if (t == nthr - 1) bout[r] = th[t].up[kC - 1];
        const int jl = lb - j0;  // index of the last column inside this thread, if any
        if (r == la && jl >= 0 && jl < kC) score = th[t].up[jl];
      }
    }
  }
  MisTrace T{trace.data(), pitch};
  int32_t res[2] = {0, 0};
  const int rc = misscore_traceback(T, a, la, b, lb, line, res);
  out[0] = score; out[1] = res[0]; out[2] = res[1]; out[3] = rc;
  if (line && rc == 0)
    for (int i = 0, j = res[0] - 1; i < j; ++i, --j) { uint8_t t = line[i]; line[i] = line[j]; line[j] = t; }
  return rc;
}

extern "C" int mis_emul(const uint8_t* a, int la, const uint8_t* b, int lb, int s_match, int s_mis, int gap,
                        int threads, int cols, int32_t* out, uint8_t* line) {
  if (cols == 16) return run<16>(a, la, b, lb, s_match, s_mis, gap, threads, out, line);
  if (cols == 4) return run<4>(a, la, b, lb, s_match, s_mis, gap, threads, out, line);
  return -1;
}
