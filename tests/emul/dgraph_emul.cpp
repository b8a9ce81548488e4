// TEST INFRASTRUCTURE: runs the device-resident graph code of the product (poa_dgraph.h: merge
// of an alignment, rank order, rank-ordered export, MSA, consensus) on the CPU with a
// sequential execution policy and compares every array it produces, read after read, with the
// host graph class (poa_graph.cpp), which the oracle tests pin.  The alignments come from the
// CPU emulation of the kernel arithmetic (poa_emul.cpp).  Not shipped.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../svscope_b200/csrc/poa_dgraph.h"
#include "poa_graph.h"

using namespace svs;

extern "C" {
void* emu_new(int ring_rows);
void emu_free(void* h);
int64_t emu_add(void* h, const uint8_t* seq, int64_t len);
int64_t emu_last_alignment(void* h, int32_t* nodes, int32_t* pos, int64_t cap);
}

namespace {

struct SeqExec {
  uint32_t nt = 128;
  template <class F> void run(F f) { for (uint32_t t = 0; t < nt; ++t) f(t, nt); }
  template <class F> void one(F f) { f(); }
  template <class F, class G> void two(F f, G g) { g(); f(); }
  void scan(uint32_t* a, uint32_t n) { for (uint32_t i = 1; i < n; ++i) a[i] += a[i - 1]; }
  void depth_sweeps(const WinMem& m, uint32_t R) { dg_depth_backward(m, R); dg_depth_forward(m, R); }
  void atomic_max(uint32_t* p, uint32_t v) { if (v > *p) *p = v; }
  void atomic_min(int32_t* p, int32_t v) { if (v < *p) *p = v; }
  void atomic_add(uint32_t* p, uint32_t v) { *p += v; }
  void suffix_min(int32_t* a, uint32_t n) { for (uint32_t i = n; i-- > 1;) if (a[i] < a[i - 1]) a[i - 1] = a[i]; }
};

template <class T>
bool same(const char* what, const T* a, const T* b, size_t n, std::string* msg, int seq) {
  for (size_t k = 0; k < n; ++k) {
    if (a[k] != b[k]) {
      char buf[256];
      std::snprintf(buf, sizeof(buf), "%s differs at %zu after sequence %d: got %lld want %lld", what, k, seq,
                    static_cast<long long>(a[k]), static_cast<long long>(b[k]));
      *msg = buf;
      return false;
    }
  }
  return true;
}

}  // namespace

extern "C" int dgraph_emul_check(const uint8_t* seqs, const int64_t* off, int n_seqs, int ring_rows, int n_threads,
                                 int tight_caps, char* msg_out, int msg_cap) {
  std::string msg;
  WinCaps caps;
  caps.nseq = static_cast<uint32_t>(n_seqs);
  for (int k = 0; k < n_seqs; ++k) {
    const uint32_t len = static_cast<uint32_t>(off[k + 1] - off[k]);
    caps.sumlen += len;
    if (len > caps.lmax) caps.lmax = len;
  }
  caps.vcap = static_cast<uint32_t>(caps.sumlen) + 1;
  caps.ecap = static_cast<uint32_t>(caps.sumlen) + 1;
  WinMem m;
  const uint64_t fixed = win_layout(nullptr, 0, caps, &m);
  std::vector<uint8_t> slot(fixed + 4096, 0xCD);
  win_layout(slot.data(), slot.size(), caps, &m);
  WinState S;
  std::memset(&S, 0, sizeof(S));
  SeqExec x;
  x.nt = static_cast<uint32_t>(n_threads);
  const Scores sc{5, -4, -8, -6, -10, -4};
  const PoaScoring psc;
  PoaGraph host;
  void* emu = emu_new(ring_rows);
  std::vector<uint32_t> merged_len;
  bool ok = true;
  const bool serial_rank = tight_caps != 0;   // flag reused: 1 = literal one-thread rank order
  for (int k = 0; k < n_seqs && ok; ++k) {
    const uint8_t* seq = seqs + off[k];
    const uint32_t L = static_cast<uint32_t>(off[k + 1] - off[k]);
    const int64_t np = emu_add(emu, seq, L);
    if (L == 0) continue;
    std::vector<int32_t> nodes(np), pos(np);
    emu_last_alignment(emu, nodes.data(), pos.data(), np);
    host.add_alignment(nodes.data(), pos.data(), static_cast<size_t>(np), seq, L);
    if (S.nseq == 0) {
      dg_init_chain(x, m, caps, &S, seq, L);
    } else {
      std::vector<int32_t> rev(2 * static_cast<size_t>(np) + 2);
      for (int64_t a = 0; a < np; ++a) {
        rev[2 * a] = nodes[np - 1 - a];
        rev[2 * a + 1] = pos[np - 1 - a];
      }
      const uint32_t n_old = S.nv;
      dg_add_alignment(x, m, caps, &S, rev.data(), static_cast<int32_t>(np), seq, L);
      if (!S.err) {
        if (serial_rank) dg_toposort_serial(m, caps, &S);
        else dg_toposort(x, m, caps, &S, L, n_old);
      }
    }
    merged_len.push_back(L);
    if (S.err) { msg = "device graph error " + std::to_string(S.err); ok = false; break; }
    // ---- compare with the host graph ------------------------------------------------------
    if (S.nv != host.num_nodes()) { msg = "node count differs"; ok = false; break; }
    const auto& r2n = host.rank_to_node();
    ok = ok && same("rank_to_node", m.node_id + 1, r2n.data(), r2n.size(), &msg, k);
    if (!ok) break;
    RankedGraph rg;
    host.export_ranked(psc, static_cast<uint32_t>(ring_rows), &rg);
    dg_export(x, m, caps, &S, sc, static_cast<uint32_t>(ring_rows), (k & 1) != 0);
    const size_t R1 = static_cast<size_t>(rg.R) + 1;
    ok = ok && same("letter", m.r_letter + 1, rg.letter.data() + 1, rg.R, &msg, k);
    ok = ok && same("pred_off", m.pred_off, rg.pred_off.data(), R1 + 1, &msg, k);
    ok = ok && same("preds", m.preds, rg.preds.data(), rg.preds.size(), &msg, k);
    ok = ok && same("flags", m.r_flags + 1, rg.flags.data() + 1, rg.R, &msg, k);
    ok = ok && same("xslot", m.xslot + 1, rg.xslot.data() + 1, rg.R, &msg, k);
    ok = ok && same("h0", m.h0 + 1, rg.h0.data() + 1, rg.R, &msg, k);
    ok = ok && same("col0code", m.col0code + 1, rg.col0code.data() + 1, rg.R, &msg, k);
    ok = ok && same("node_id", m.node_id + 1, rg.node_id.data() + 1, rg.R, &msg, k);
    ok = ok && same("single_before", m.single_before, rg.single_before.data(), R1 + 1, &msg, k);
    ok = ok && same("depth", m.depth + 4, rg.depth.data() + 4, 4 * static_cast<size_t>(rg.R), &msg, k);
    if (ok && (S.n_export != rg.n_export || S.max_indeg != std::max<uint32_t>(1, rg.max_indeg))) {
      msg = "n_export / max_indeg differ";
      ok = false;
    }
  }
  if (ok && S.nseq > 0) {
    dg_export(x, m, caps, &S, sc, static_cast<uint32_t>(ring_rows), true);
    // consensus
    std::vector<int32_t> score(S.nv + 2), pred(S.nv + 2);
    std::vector<uint8_t> cons(S.nv + 2);
    const uint32_t cl = dg_consensus_serial(m, &S, score.data(), pred.data(), cons.data());
    const std::string want = host.consensus();
    if (want.size() != cl || std::memcmp(want.data(), cons.data(), cl) != 0) {
      msg = "consensus differs: got " + std::string(reinterpret_cast<char*>(cons.data()), cl) + " want " + want;
      ok = false;
    }
    // MSA
    if (ok) {
      std::vector<uint32_t> head(S.nv + 2), col(S.nv + 2);
      dg_msa_columns(x, m, &S, head.data(), col.data());
      const auto rows = host.msa();
      const uint32_t W = S.msa_cols;
      if (rows.size() != merged_len.size() || (rows.size() && rows[0].size() != W)) {
        msg = "msa shape differs";
        ok = false;
      } else {
        std::vector<uint8_t> out(static_cast<size_t>(W) * rows.size() + 1);
        dg_msa_rows(x, m, &S, col.data(), merged_len.data(), static_cast<uint32_t>(rows.size()), out.data());
        for (size_t r = 0; r < rows.size() && ok; ++r) {
          if (std::memcmp(rows[r].data(), out.data() + r * W, W) != 0) {
            msg = "msa row " + std::to_string(r) + " differs";
            ok = false;
          }
        }
      }
    }
  }
  if (getenv("DGRAPH_DEBUG")) std::fprintf(stderr, "nodes %u serial fallbacks %u last rounds %u\n", S.nv, S.topo_serial, S.topo_rounds);
  emu_free(emu);
  if (msg_out && msg_cap > 0) {
    std::snprintf(msg_out, static_cast<size_t>(msg_cap), "%s", msg.c_str());
  }
  return ok ? 0 : 1;
}
