// Partial-order graph of one window, resident in DEVICE memory, and every graph step of
// `spoa.poa(sequences, 1)` between two alignments (reference call sites src/DataScanner.py:206,213
// and src/DecisionMaker.py:160,171; behaviour of spoa's Graph as restated in SURVEY.md
// Appendix B): merge of an alignment path (AddAlignment), rank
// order (TopologicalSort, depth-first over node ids), export of the rank-ordered arrays the
// DP kernel consumes, MSA rows and heaviest-bundle consensus.
//
// One CTA owns one window from its first to its last read, so nothing of this ever visits
// the host.  The code is written against an execution policy X:
//     x.run(f)     every thread of the CTA calls f(tid, n_threads); barrier afterwards
//     x.one(f)     thread 0 calls f(); barrier afterwards
//     x.two(f, g)  two threads of different warps call f() and g() concurrently; barrier
//     x.scan(a, n) in-place inclusive prefix sum of a[0..n); barrier afterwards
// On the device X is CtaExec (poa_window.cu); tests/emul runs the same functions with a
// sequential policy on the CPU and compares every array with the host graph (poa_graph.cpp).
// Rule for the code below: values shared between threads live in WinState or in the slot
// arrays, never in locals that outlive one run()/one() call.
#pragma once
#include <cstdint>

#include "poa_cell.h"
#include "poa_task.h"

namespace svs {

constexpr int kMaxAligned = 7;   // an aligned group holds at most 8 nodes (distinct letters)

// per-window status (result record of the window kernel); anything but kWinOk fails THAT window only
enum WinStatus : int32_t {
  kWinOk = 0,
  kWinNodeCap = 1,      // more graph nodes than the slot was laid out for
  kWinEdgeCap = 2,
  kWinAlignedCap = 3,   // aligned group larger than kMaxAligned + 1
  kWinStackCap = 4,
  kWinCodesCap = 5,     // traceback codes of one alignment do not fit the slot
  kWinTraceback = 6,
  kWinScoreSpan = 7,    // |V| + L too large for the packed cell format
  kWinOutCap = 8,       // output arena exhausted
  kWinPending = 9,
  kWinIndeg = 10,       // a node with more than 31 in-edges (index field of the traceback codes)
};

struct WinCaps {
  uint32_t vcap = 0;     // graph nodes
  uint32_t ecap = 0;     // graph edges
  uint32_t lmax = 0;     // longest sequence
  uint32_t nseq = 0;
  uint64_t sumlen = 0;   // all sequences
};

// Pointers into the scratch slot of the CTA (laid out by win_layout).
struct WinMem {
  // graph, indexed by node id (creation order)
  uint8_t* letter;
  uint8_t* has_out;
  uint8_t* n_al;
  uint8_t* state;        // toposort: 0 unseen, 1 open, 2 emitted
  uint8_t* as_al;        // toposort: reached through an aligned link
  int32_t* in_head;      // first / last in-edge (linked in creation order), -1 = none
  int32_t* in_tail;
  int32_t* al;           // [vcap][kMaxAligned] aligned alternatives, in creation order
  int32_t* e_tail;       // edges
  int32_t* e_next;
  int32_t* e_w;
  uint32_t* row_of;      // node -> row (rank + 1)
  uint32_t* stack;       // toposort stack pool [ecap + 2 vcap]
  int32_t* f;            // [vcap] smallest node id among the node's descendants-or-self (aligned links included):
                         //        the node is emitted by the depth-first walk that starts at node f
  uint32_t* cnt;         // [vcap+2] rank-order scratch: nodes per walk -> first rank of the walk
  uint32_t* soff;        // [vcap+2] rank-order scratch: stack need per walk -> first stack slot
  uint32_t* indeg;       // [vcap] number of in-edges
  uint8_t* stamp;        // [vcap] rank-order scratch: propagation round in which the node is due
  int32_t* at;           // [lmax] merge scratch: node aligned to / chosen for each read position
  int32_t* flag;         // [lmax] merge scratch
  int32_t* path_node;    // [sumlen] node of every base of every merged sequence
  uint32_t* seq_len;     // [nseq] lengths of the merged (non-empty) sequences, in merge order
  // rank-ordered view (the arrays of PoaTask), indexed by row 0..R
  uint8_t* r_letter;
  uint8_t* r_flags;
  uint32_t* pred_off;    // [vcap+2]
  uint32_t* preds;       // [ecap + vcap]
  int32_t* pred_w;       // [ecap + vcap] edge weights in the same order (consensus)
  int32_t* xslot;
  int32_t* h0;
  uint16_t* col0code;
  uint32_t* node_id;     // row -> node
  uint32_t* single_before;
  int32_t* depth;        // [vcap+1][4]
  int32_t* band;         // [vcap+1][2]
  uint32_t* coff;        // [vcap+3] code-row offsets of the running alignment
  void* tbrow;           // [vcap+2] 32-byte traceback records of the running alignment (poa_dp2.cuh TbRow)
  int32_t* bnd;          // [2][4][vcap+1]
  int32_t* result;       // [4]
  int32_t* path;         // [2 * path_cap]
  uint32_t path_cap;
  uint8_t* dyn;          // rest of the slot: exported rows + traceback codes of the running alignment
  uint64_t dyn_bytes;
};

SVS_HD uint64_t dg_align(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

// Carves the slot; returns the bytes of the fixed part (everything before `dyn`).
SVS_HD uint64_t win_layout(uint8_t* base, uint64_t slot_bytes, const WinCaps& c, WinMem* m) {
  uint64_t off = 0;
  auto take = [&](uint64_t bytes) -> uint8_t* {
    uint8_t* p = base + off;
    off += dg_align(bytes, 256);
    return p;
  };
  const uint64_t V = c.vcap, V1 = V + 2, E = c.ecap, L = static_cast<uint64_t>(c.lmax) + 8;
  m->letter = take(V);
  m->has_out = take(V);
  m->n_al = take(V);
  m->state = take(V + 2);
  m->as_al = take(V + 2);
  m->in_head = reinterpret_cast<int32_t*>(take(4 * V));
  m->in_tail = reinterpret_cast<int32_t*>(take(4 * V));
  m->al = reinterpret_cast<int32_t*>(take(4 * V * kMaxAligned));
  m->e_tail = reinterpret_cast<int32_t*>(take(4 * E));
  m->e_next = reinterpret_cast<int32_t*>(take(4 * E));
  m->e_w = reinterpret_cast<int32_t*>(take(4 * E));
  m->row_of = reinterpret_cast<uint32_t*>(take(4 * V));
  m->stack = reinterpret_cast<uint32_t*>(take(4 * (E + 2 * V)));
  m->f = reinterpret_cast<int32_t*>(take(4 * V));
  m->cnt = reinterpret_cast<uint32_t*>(take(4 * (V + 2)));
  m->soff = reinterpret_cast<uint32_t*>(take(4 * (V + 2)));
  m->indeg = reinterpret_cast<uint32_t*>(take(4 * V));
  m->stamp = take(V + 2);
  m->at = reinterpret_cast<int32_t*>(take(4 * L));
  m->flag = reinterpret_cast<int32_t*>(take(4 * L));
  m->path_node = reinterpret_cast<int32_t*>(take(4 * (c.sumlen + 8)));
  m->seq_len = reinterpret_cast<uint32_t*>(take(4 * (static_cast<uint64_t>(c.nseq) + 8)));
  m->r_letter = take(V1);
  m->r_flags = take(V1);
  m->pred_off = reinterpret_cast<uint32_t*>(take(4 * (V1 + 1)));
  m->preds = reinterpret_cast<uint32_t*>(take(4 * (E + V1)));
  m->pred_w = reinterpret_cast<int32_t*>(take(4 * (E + V1)));
  m->xslot = reinterpret_cast<int32_t*>(take(4 * V1));
  m->h0 = reinterpret_cast<int32_t*>(take(4 * V1));
  m->col0code = reinterpret_cast<uint16_t*>(take(2 * V1));
  m->node_id = reinterpret_cast<uint32_t*>(take(4 * V1));
  m->single_before = reinterpret_cast<uint32_t*>(take(4 * (V1 + 1)));
  m->depth = reinterpret_cast<int32_t*>(take(16 * V1));
  m->band = reinterpret_cast<int32_t*>(take(8 * V1));
  m->coff = reinterpret_cast<uint32_t*>(take(4 * (V1 + 2)));
  m->tbrow = take(32 * V1);
  m->bnd = reinterpret_cast<int32_t*>(take(32 * V1));
  m->result = reinterpret_cast<int32_t*>(take(64));
  m->path_cap = static_cast<uint32_t>(V + c.lmax + 2);
  m->path = reinterpret_cast<int32_t*>(take(8ull * m->path_cap));
  m->dyn = base + off;
  m->dyn_bytes = slot_bytes > off ? slot_bytes - off : 0;
  return off;
}

// Shared by the threads of the CTA (shared memory on the device).
struct WinState {
  uint32_t nv, ne;        // nodes, edges
  uint32_t nseq;          // non-empty sequences merged so far
  uint64_t path_off;      // bases of the merged sequences (offset of the next path in path_node)
  int32_t err;
  uint32_t max_indeg, n_export, n_single;
  uint32_t n_new;         // merge: new nodes / edges of the running read
  uint32_t msa_cols;
  uint32_t chg[2];        // rank order: a propagation round lowered some f (alternating slots)
  int32_t err_pending;    // error raised inside a parallel phase, filed into err after its barrier
  uint32_t topo_rounds, topo_serial;   // statistics
  int32_t last_score;
  uint32_t last_len;
};

// ---------------------------------------------------------------------------------------------
// first sequence: a chain (spoa: empty alignment -> AddSequence; rank order of a chain is
// its node order)
template <class X>
SVS_HD void dg_init_chain(X& x, const WinMem& m, const WinCaps& c, WinState* S, const uint8_t* seq, uint32_t L) {
  x.one([&]() {
    S->err = (L > c.vcap) ? kWinNodeCap : ((L > 0 && L - 1 > c.ecap) ? kWinEdgeCap : S->err);
  });
  if (S->err) return;
  x.run([&](uint32_t tid, uint32_t nt) {
    for (uint32_t p = tid; p < L; p += nt) {
      m.letter[p] = seq[p];
      m.has_out[p] = (p + 1 < L) ? 1 : 0;
      m.n_al[p] = 0;
      m.in_head[p] = m.in_tail[p] = static_cast<int32_t>(p) - 1;   // edge p-1 is (p-1 -> p); -1 for p = 0
      if (p > 0) {
        m.e_tail[p - 1] = static_cast<int32_t>(p - 1);
        m.e_next[p - 1] = -1;
        m.e_w[p - 1] = 2;   // unit weights: 1 + 1
      }
      m.node_id[p + 1] = p;
      m.row_of[p] = p + 1;
      m.f[p] = static_cast<int32_t>(p);   // all descendants have larger ids
      m.indeg[p] = p > 0 ? 1u : 0u;
      m.path_node[S->path_off + p] = static_cast<int32_t>(p);
    }
  });
  x.one([&]() {
    S->nv = L;
    S->ne = L - 1;
    m.seq_len[0] = L;
    S->nseq = 1;
    S->path_off += L;
  });
}

// ---------------------------------------------------------------------------------------------
// AddAlignment for a global alignment (every read position appears in exactly one pair).
// `pairs` = (node id | -1, read position | -1) in REVERSE order, as the traceback writes them.
template <class X>
SVS_HD void dg_add_alignment(X& x, const WinMem& m, const WinCaps& c, WinState* S, const int32_t* pairs,
                             int32_t n_pairs, const uint8_t* seq, uint32_t L) {
  // 1. node the read position is aligned to
  x.run([&](uint32_t tid, uint32_t nt) {
    for (uint32_t p = tid; p < L; p += nt) m.at[p] = -1;
  });
  x.run([&](uint32_t tid, uint32_t nt) {
    for (int32_t k = static_cast<int32_t>(tid); k < n_pairs; k += static_cast<int32_t>(nt)) {
      const int32_t pos = pairs[2 * k + 1];
      if (pos >= 0 && static_cast<uint32_t>(pos) < L) m.at[pos] = pairs[2 * k];
    }
  });
  // 2. same letter -> that node; an aligned alternative with the letter -> that one; else a
  //    new node (flag = 1), cross-linked into the aligned group of at[p] if there is one
  x.run([&](uint32_t tid, uint32_t nt) {
    for (uint32_t p = tid; p < L; p += nt) {
      const int32_t a = m.at[p];
      const uint8_t ch = seq[p];
      int32_t need = 1;
      if (a >= 0) {
        if (m.letter[a] == ch) {
          need = 0;
        } else {
          const uint32_t n = m.n_al[a];
          for (uint32_t k = 0; k < n; ++k) {
            const int32_t b = m.al[static_cast<uint64_t>(a) * kMaxAligned + k];
            if (m.letter[b] == ch) { m.at[p] = b; need = 0; break; }
          }
        }
      }
      m.flag[p] = need;
    }
  });
  x.scan(reinterpret_cast<uint32_t*>(m.flag), L);
  x.one([&]() {
    const uint32_t n_new = L ? static_cast<uint32_t>(m.flag[L - 1]) : 0;
    S->n_new = n_new;
    if (S->nv + n_new > c.vcap) S->err = kWinNodeCap;
  });
  if (S->err) return;
  // 3. create the nodes (ids in read order) and extend the aligned groups
  x.run([&](uint32_t tid, uint32_t nt) {
    for (uint32_t p = tid; p < L; p += nt) {
      const uint32_t incl = static_cast<uint32_t>(m.flag[p]);
      const uint32_t before = p ? static_cast<uint32_t>(m.flag[p - 1]) : 0;
      if (incl == before) continue;
      const uint32_t id = S->nv + before;
      const int32_t a = m.at[p];
      m.letter[id] = seq[p];
      m.has_out[id] = 0;
      m.indeg[id] = 0;
      m.in_head[id] = m.in_tail[id] = -1;
      uint32_t n_mine = 0;
      if (a >= 0) {
        const uint32_t n = m.n_al[a];
        if (n + 1 > static_cast<uint32_t>(kMaxAligned)) {
          S->err_pending = kWinAlignedCap;
        } else {
          for (uint32_t k = 0; k < n; ++k) {
            const int32_t b = m.al[static_cast<uint64_t>(a) * kMaxAligned + k];
            m.al[static_cast<uint64_t>(b) * kMaxAligned + m.n_al[b]] = static_cast<int32_t>(id);
            m.n_al[b] = static_cast<uint8_t>(m.n_al[b] + 1);
            m.al[static_cast<uint64_t>(id) * kMaxAligned + n_mine++] = b;
          }
          m.al[static_cast<uint64_t>(a) * kMaxAligned + n] = static_cast<int32_t>(id);
          m.n_al[a] = static_cast<uint8_t>(n + 1);
          m.al[static_cast<uint64_t>(id) * kMaxAligned + n_mine++] = a;
        }
      }
      m.n_al[id] = static_cast<uint8_t>(n_mine);
      m.at[p] = -2 - static_cast<int32_t>(id);   // resolved below (neighbours still read flag[])
    }
  });
  x.one([&]() { if (S->err_pending) { S->err = S->err_pending; S->err_pending = 0; } });
  if (S->err) return;
  x.run([&](uint32_t tid, uint32_t nt) {
    for (uint32_t p = tid; p < L; p += nt) {
      const int32_t a = m.at[p];
      if (a <= -2) m.at[p] = -2 - a;
      m.path_node[S->path_off + p] = m.at[p];
    }
  });
  // 4. edges between consecutive read nodes: an existing edge gains weight, a new one is
  //    appended to the in-list of its head (= in-edge order of first traversal)
  x.run([&](uint32_t tid, uint32_t nt) {
    for (uint32_t p = tid; p < L; p += nt) {
      int32_t need = 0;
      if (p > 0) {
        const int32_t tail = m.at[p - 1], head = m.at[p];
        need = 1;
        for (int32_t e = m.in_head[head]; e >= 0; e = m.e_next[e]) {
          if (m.e_tail[e] == tail) { m.e_w[e] += 2; need = 0; break; }
        }
      }
      m.flag[p] = need;
    }
  });
  x.scan(reinterpret_cast<uint32_t*>(m.flag), L);
  x.one([&]() {
    const uint32_t n_new = L ? static_cast<uint32_t>(m.flag[L - 1]) : 0;
    if (S->ne + n_new > c.ecap) S->err = kWinEdgeCap;
  });
  if (S->err) return;
  x.run([&](uint32_t tid, uint32_t nt) {
    for (uint32_t p = tid + 1; p < L; p += nt) {
      const uint32_t incl = static_cast<uint32_t>(m.flag[p]), before = static_cast<uint32_t>(m.flag[p - 1]);
      if (incl == before) continue;
      const int32_t e = static_cast<int32_t>(S->ne + before);
      const int32_t tail = m.at[p - 1], head = m.at[p];
      m.e_tail[e] = tail;
      m.e_next[e] = -1;
      m.e_w[e] = 2;
      if (m.in_tail[head] < 0) m.in_head[head] = e; else m.e_next[m.in_tail[head]] = e;
      m.in_tail[head] = e;
      m.indeg[head] += 1;
      m.has_out[tail] = 1;
    }
  });
  x.one([&]() {
    S->nv += S->n_new;
    S->ne += L ? static_cast<uint32_t>(m.flag[L - 1]) : 0;
    m.seq_len[S->nseq] = L;
    S->nseq += 1;
    S->path_off += L;
  });
}

// ---------------------------------------------------------------------------------------------
// Rank order: depth-first emission over node ids in creation order.  A node is emitted once
// every in-edge tail and (unless it was itself reached as an aligned alternative) every
// aligned node has been emitted; its aligned alternatives follow it immediately.  Literal
// transcription of the stack discipline ("push all unfinished, then re-examine"), one thread.
SVS_HD void dg_toposort_serial(const WinMem& m, const WinCaps& c, WinState* S) {
  const uint32_t n = S->nv;
  for (uint32_t v = 0; v < n; ++v) { m.state[v] = 0; m.as_al[v] = 0; }
  uint32_t rank = 0, top = 0;
  const uint32_t cap = c.vcap;
  for (uint32_t root = 0; root < n; ++root) {
    if (m.state[root] != 0) continue;
    m.stack[top++] = root;
    while (top > 0) {
      const uint32_t cur = m.stack[top - 1];
      bool ready = true;
      if (m.state[cur] != 2) {
        for (int32_t e = m.in_head[cur]; e >= 0; e = m.e_next[e]) {
          const uint32_t t = static_cast<uint32_t>(m.e_tail[e]);
          if (m.state[t] != 2) {
            if (top >= cap) { S->err = kWinStackCap; return; }
            m.stack[top++] = t;
            ready = false;
          }
        }
        const uint32_t na = m.n_al[cur];
        if (!m.as_al[cur]) {
          for (uint32_t k = 0; k < na; ++k) {
            const uint32_t a = static_cast<uint32_t>(m.al[static_cast<uint64_t>(cur) * kMaxAligned + k]);
            if (m.state[a] != 2) {
              if (top >= cap) { S->err = kWinStackCap; return; }
              m.stack[top++] = a;
              m.as_al[a] = 1;
              ready = false;
            }
          }
        }
        if (ready) {
          m.state[cur] = 2;
          m.f[cur] = static_cast<int32_t>(root);
          if (!m.as_al[cur]) {
            m.node_id[++rank] = cur;
            m.row_of[cur] = rank;
            for (uint32_t k = 0; k < na; ++k) {
              const uint32_t a = static_cast<uint32_t>(m.al[static_cast<uint64_t>(cur) * kMaxAligned + k]);
              m.node_id[++rank] = a;
              m.row_of[a] = rank;
            }
          }
        } else {
          m.state[cur] = 1;
        }
      }
      if (ready) --top;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// The same rank order, computed by the whole CTA.
//
// The depth-first walk that starts at node u (outer loop over node ids) emits exactly the nodes
// whose smallest descendant-or-self id (in-edges reversed, aligned links both ways) is u:
// every ancestor of u that an earlier walk has not emitted.  With f(v) = that id,
//   * walks are independent: walk u only needs to know which nodes belong to earlier walks
//     (f < u), so all walks run in parallel, one thread each, with the literal stack
//     discipline restricted to the nodes with f == u;
//   * walk u writes ranks [sum of |walk u'| for u' < u, ...): one prefix sum;
//   * f is maintained incrementally: new descendants only ever lower it.  After a merge the
//     read's own path gives f(path[p]) <= min over q >= p of f(path[q]) (a suffix minimum over
//     the read positions), new nodes also take the f of the group they were aligned into, and
//     the remaining decreases are propagated to in-edge tails and aligned nodes round by
//     round until nothing changes (they stay local: a lowered f stops at the first ancestor
//     that already had a smaller one).
// `path` = node of every position of the read just merged (m.at), n_old = nodes before it.
// Falls back to the one-thread walk if the propagation does not settle or the stack pool is short.
template <class X>
SVS_HD void dg_toposort(X& x, const WinMem& m, const WinCaps& c, WinState* S, uint32_t L, uint32_t n_old) {
  const uint32_t n = S->nv;
  const uint32_t pool = c.ecap + 2 * c.vcap;
  // ---- f of the path nodes ---------------------------------------------------------------------
  x.run([&](uint32_t tid, uint32_t nt) {
    for (uint32_t v = tid; v < n; v += nt) m.stamp[v] = 0;
    for (uint32_t p = tid; p < L; p += nt) {
      const uint32_t v = static_cast<uint32_t>(m.at[p]);
      int32_t b;
      if (v >= n_old) {
        b = static_cast<int32_t>(v);
        const uint32_t na = m.n_al[v];
        for (uint32_t k = 0; k < na; ++k) {
          const uint32_t a = static_cast<uint32_t>(m.al[static_cast<uint64_t>(v) * kMaxAligned + k]);
          if (a < n_old && m.f[a] < b) b = m.f[a];   // (aligned nodes created by this read lie elsewhere on the path)
        }
      } else {
        b = m.f[v];
      }
      m.flag[p] = b;
    }
  });
  x.suffix_min(m.flag, L);
  x.run([&](uint32_t tid, uint32_t nt) {
    for (uint32_t p = tid; p < L; p += nt) {
      const uint32_t v = static_cast<uint32_t>(m.at[p]);
      const int32_t g = m.flag[p];
      if (v >= n_old) { m.f[v] = g; m.stamp[v] = 1; }
      else if (g < m.f[v]) { m.f[v] = g; m.stamp[v] = 1; }
    }
    if (tid == 0) { S->chg[1] = 1; S->chg[0] = 0; S->topo_rounds = 0; }
  });
  // ---- propagate the decreases -----------------------------------------------------------------
  // (round r reads chg[r & 1], which round r - 1 wrote, and writes the other slot)
  uint32_t round = 1;
  while (S->chg[round & 1] && round < 250) {
    x.one([&]() { S->chg[(round + 1) & 1] = 0; });
    const uint8_t due = static_cast<uint8_t>(round), next = static_cast<uint8_t>(round + 1);
    x.run([&](uint32_t tid, uint32_t nt) {
      bool any = false;
      for (uint32_t v = tid; v < n; v += nt) {
        if (m.stamp[v] != due) continue;
        const int32_t fv = m.f[v];
        for (int32_t e = m.in_head[v]; e >= 0; e = m.e_next[e]) {
          const uint32_t t = static_cast<uint32_t>(m.e_tail[e]);
          if (m.f[t] > fv) { x.atomic_min(&m.f[t], fv); m.stamp[t] = next; any = true; }
        }
        const uint32_t na = m.n_al[v];
        for (uint32_t k = 0; k < na; ++k) {
          const uint32_t a = static_cast<uint32_t>(m.al[static_cast<uint64_t>(v) * kMaxAligned + k]);
          if (m.f[a] > fv) { x.atomic_min(&m.f[a], fv); m.stamp[a] = next; any = true; }
        }
      }
      if (any) S->chg[(round + 1) & 1] = 1;
    });
    ++round;
  }
  bool serial = S->chg[round & 1] != 0;
  // ---- size and stack need of every walk -------------------------------------------------------
  if (!serial) {
    x.run([&](uint32_t tid, uint32_t nt) {
      for (uint32_t v = tid; v <= n; v += nt) { m.cnt[v + 1] = 0; m.soff[v + 1] = 0; }
      for (uint32_t v = tid; v < n; v += nt) { m.state[v] = 0; m.as_al[v] = 0; }
      if (tid == 0) { m.cnt[0] = 0; m.soff[0] = 0; }
    });
    x.run([&](uint32_t tid, uint32_t nt) {
      for (uint32_t v = tid; v < n; v += nt) {
        const uint32_t u = static_cast<uint32_t>(m.f[v]);
        x.atomic_add(&m.cnt[u + 1], 1u);
        x.atomic_add(&m.soff[u + 1], m.indeg[v] + m.n_al[v] + (u == v ? 1u : 0u));
      }
    });
    x.scan(m.cnt + 1, n);      // cnt[u]  = nodes of the walks before u  = first rank of walk u
    x.scan(m.soff + 1, n);     // soff[u] = first stack slot of walk u
    serial = m.soff[n] > pool;
  }
  if (serial) {
    x.one([&]() { S->topo_serial += 1; dg_toposort_serial(m, c, S); });
    return;
  }
  // ---- the walks ---------------------------------------------------------------------------------
  x.run([&](uint32_t tid, uint32_t nt) {
    if (tid == 0) S->topo_rounds = round;
    for (uint32_t root = tid; root < n; root += nt) {
      if (static_cast<uint32_t>(m.f[root]) != root) continue;
      uint32_t rank = m.cnt[root];
      if (m.cnt[root + 1] - rank == 1) {   // the walk is the node itself
        m.node_id[rank + 1] = root;
        continue;
      }
      uint32_t* stack = m.stack + m.soff[root];
      uint32_t top = 0;
      auto done = [&](uint32_t t) -> bool { return static_cast<uint32_t>(m.f[t]) != root || m.state[t] == 2; };
      stack[top++] = root;
      while (top > 0) {
        const uint32_t cur = stack[top - 1];
        bool ready = true;
        if (m.state[cur] != 2) {
          for (int32_t e = m.in_head[cur]; e >= 0; e = m.e_next[e]) {
            const uint32_t t = static_cast<uint32_t>(m.e_tail[e]);
            if (!done(t)) { stack[top++] = t; ready = false; }
          }
          const uint32_t na = m.n_al[cur];
          if (!m.as_al[cur]) {
            for (uint32_t k = 0; k < na; ++k) {
              const uint32_t a = static_cast<uint32_t>(m.al[static_cast<uint64_t>(cur) * kMaxAligned + k]);
              if (!done(a)) { stack[top++] = a; m.as_al[a] = 1; ready = false; }
            }
          }
          if (ready) {
            m.state[cur] = 2;
            if (!m.as_al[cur]) {
              m.node_id[++rank] = cur;
              for (uint32_t k = 0; k < na; ++k)
                m.node_id[++rank] = static_cast<uint32_t>(m.al[static_cast<uint64_t>(cur) * kMaxAligned + k]);
            }
          } else {
            m.state[cur] = 1;
          }
        }
        if (ready) --top;
      }
    }
  });
  x.run([&](uint32_t tid, uint32_t nt) {
    for (uint32_t r = 1 + tid; r <= n; r += nt) m.row_of[m.node_id[r]] = r;
  });
}

// Path-length intervals of every row (nodes on source->row paths, row included: dmin, dmax;
// nodes on row->sink paths, row excluded: smin, smax), literal one-thread sweeps.
SVS_HD void dg_depth_forward(const WinMem& m, uint32_t R) {
  int32_t* dp = m.depth;
  for (uint32_t i = 1; i <= R; ++i) {
    int32_t lo = INT32_MAX, hi = 0;
    for (uint32_t k = m.pred_off[i]; k < m.pred_off[i + 1]; ++k) {
      const uint32_t p = m.preds[k];
      const int32_t a = dp[4 * p], b = dp[4 * p + 1];
      lo = a < lo ? a : lo;
      hi = b > hi ? b : hi;
    }
    dp[4 * i] = lo + 1;
    dp[4 * i + 1] = hi + 1;
  }
}

SVS_HD void dg_depth_backward(const WinMem& m, uint32_t R) {
  int32_t* dp = m.depth;
  uint8_t* seen = m.state;   // rank-order scratch is free here
  for (uint32_t i = 0; i <= R; ++i) { seen[i] = 0; dp[4 * i + 2] = 0; dp[4 * i + 3] = 0; }
  for (uint32_t i = R; i >= 1; --i) {
    const int32_t a = dp[4 * i + 2] + 1, b = dp[4 * i + 3] + 1;
    for (uint32_t k = m.pred_off[i]; k < m.pred_off[i + 1]; ++k) {
      const uint32_t p = m.preds[k];
      if (p == 0) continue;
      if (!seen[p]) { dp[4 * p + 2] = a; dp[4 * p + 3] = b; seen[p] = 1; }
      else {
        if (a < dp[4 * p + 2]) dp[4 * p + 2] = a;
        if (b > dp[4 * p + 3]) dp[4 * p + 3] = b;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Rank-ordered arrays for the DP kernel (same content as PoaGraph::export_ranked).
// Row r (1..R) is rank r-1; row 0 is the virtual source row.
template <class X>
SVS_HD void dg_export(X& x, const WinMem& m, const WinCaps& c, WinState* S, const Scores& sc, uint32_t ring_rows,
                      bool with_weights) {
  const uint32_t R = S->nv;
  // in-degree of every row (a node without in-edge hangs off the virtual source row)
  x.run([&](uint32_t tid, uint32_t nt) {
    if (tid == 0) {
      m.pred_off[0] = 0; m.pred_off[1] = 0;
      m.r_letter[0] = 0; m.r_flags[0] = 0; m.xslot[0] = -1; m.h0[0] = 0; m.col0code[0] = 0; m.node_id[0] = 0;
      m.depth[0] = m.depth[1] = m.depth[2] = m.depth[3] = 0;
      S->max_indeg = 1;
    }
    for (uint32_t i = 1 + tid; i <= R; i += nt) {
      const uint32_t node = m.node_id[i];
      const uint32_t d = m.indeg[node];
      m.pred_off[i + 1] = d ? d : 1;
      m.r_letter[i] = m.letter[node];
      m.r_flags[i] = m.has_out[node] ? 0 : kFlagSink;
    }
  });
  x.scan(m.pred_off + 2, R);
  // predecessor rows in stored in-edge order; chain flag; single-predecessor marker
  x.run([&](uint32_t tid, uint32_t nt) {
    uint32_t my_max = 1;
    for (uint32_t i = 1 + tid; i <= R; i += nt) {
      const uint32_t node = m.node_id[i];
      uint32_t k = m.pred_off[i];
      const uint32_t d = m.pred_off[i + 1] - k;
      if (m.in_head[node] < 0) {
        m.preds[k] = 0;
        if (with_weights) m.pred_w[k] = 0;
      } else {
        for (int32_t e = m.in_head[node]; e >= 0; e = m.e_next[e], ++k) {
          m.preds[k] = m.row_of[m.e_tail[e]];
          if (with_weights) m.pred_w[k] = m.e_w[e];
        }
      }
      if (d > my_max) my_max = d;
      const bool single = (d == 1);
      m.single_before[i + 1] = single ? 1u : 0u;
      if (single && m.preds[m.pred_off[i]] + 1 == i) m.r_flags[i] |= kFlagChain;
      m.xslot[i] = 0;
    }
    if (my_max > 1) x.atomic_max(&S->max_indeg, my_max);
  });
  // rows with a successor further away than the on-chip ring are exported to global memory
  x.run([&](uint32_t tid, uint32_t nt) {
    for (uint32_t i = 1 + tid; i <= R; i += nt) {
      for (uint32_t k = m.pred_off[i]; k < m.pred_off[i + 1]; ++k) {
        const uint32_t p = m.preds[k];
        if (p != 0 && i - p > ring_rows) m.xslot[p] = 1;
      }
    }
    if (tid == 0) { m.single_before[0] = 0; m.single_before[1] = 0; }
  });
  x.run([&](uint32_t tid, uint32_t nt) {
    for (uint32_t i = 1 + tid; i <= R; i += nt) {
      if (m.xslot[i]) m.r_flags[i] |= kFlagExport;
    }
  });
  x.scan(reinterpret_cast<uint32_t*>(m.xslot) + 1, R);         // inclusive count of exported rows
  x.scan(m.single_before + 2, R);                               // single_before[i+1] = singles among rows 1..i
  x.run([&](uint32_t tid, uint32_t nt) {
    if (tid == 0) {
      S->n_export = R ? static_cast<uint32_t>(m.xslot[R]) : 0;
      S->n_single = m.single_before[R + 1];
    }
  });
  x.run([&](uint32_t tid, uint32_t nt) {   // inclusive count -> slot index (or -1); back to front within a thread is not needed:
    for (uint32_t i = 1 + tid; i <= R; i += nt) {   // every thread only rewrites its own rows, reading flags
      m.xslot[i] = (m.r_flags[i] & kFlagExport) ? m.xslot[i] - 1 : -1;
    }
  });
  // path-length intervals (pruning bounds, poa_cell.h cell_bound): two sequential sweeps over
  // the rank order (dg_depth_forward / dg_depth_backward below); the policy decides how to
  // run them (CPU: one after the other; device: two warps, each with a ring of recent rows in
  // shared memory)
  x.depth_sweeps(m, R);
  // column 0: per gap piece the best in-edge tail plus one extension, sources open a fresh
  // gap, i.e. F0 = g + (dmin-1) e, O0 = q + (dmin-1) c with dmin = fewest nodes on a path from
  // a source to the row; the traceback code repeats the engine's equality tests there
  x.run([&](uint32_t tid, uint32_t nt) {
    auto f0 = [&](uint32_t r) -> int32_t { return r == 0 ? 0 : sc.g + (m.depth[4 * r] - 1) * sc.e; };
    auto o0 = [&](uint32_t r) -> int32_t { return r == 0 ? 0 : sc.q + (m.depth[4 * r] - 1) * sc.c; };
    auto hh = [&](uint32_t r) -> int32_t { return r == 0 ? 0 : imax(f0(r), o0(r)); };
    for (uint32_t i = 1 + tid; i <= R; i += nt) {
      const uint32_t b = m.pred_off[i], e = m.pred_off[i + 1];
      const int32_t H = hh(i), Fi = f0(i), Oi = o0(i);
      m.h0[i] = H;
      uint32_t kH = 0, ext = 0;
      bool found = false;
      for (uint32_t k = b; k < e && !found; ++k) {
        const uint32_t p = m.preds[k];
        const int32_t Hp = hh(p), Fp = f0(p), Op = o0(p);
        if (H == Fp + sc.e) { found = true; ext = 1; }
        else if (H == Hp + sc.g) { found = true; }
        else if (H == Op + sc.c) { found = true; ext = 1; }
        else if (H == Hp + sc.q) { found = true; }
        if (found) kH = k - b;
      }
      uint32_t kU = 0, stop = 1;
      if (m.preds[b] != 0) {   // the row has real in-edges
        bool hit = false;
        stop = 0;
        for (uint32_t k = b; k < e && !hit; ++k) {
          const uint32_t p = m.preds[k];
          if (Fi == hh(p) + sc.g) { hit = true; stop = 1; }
          else if (Fi == f0(p) + sc.e) { hit = true; stop = 0; }
          else if (Oi == hh(p) + sc.q) { hit = true; stop = 1; }
          else if (Oi == o0(p) + sc.c) { hit = true; stop = 0; }
          if (hit) kU = k - b;
        }
        if (!hit) { kU = kNoPred; stop = 0; }
      }
      m.col0code[i] = make_code(kMoveVert, ext, 0, stop, kH, kU);
    }
  });
}

// ---------------------------------------------------------------------------------------------
// MSA: one column per aligned group in rank order; every merged sequence writes its letters
// into the columns of its nodes, '-' elsewhere.  `col_of` [vcap] and `out` (rows x cols) are
// caller-provided; dg_msa_columns returns the column count in S->msa_cols.
template <class X>
SVS_HD void dg_msa_columns(X& x, const WinMem& m, WinState* S, uint32_t* head_flag /*[R+1]*/, uint32_t* col_of) {
  const uint32_t R = S->nv;
  x.run([&](uint32_t tid, uint32_t nt) {
    for (uint32_t i = 1 + tid; i <= R; i += nt) {
      const uint32_t node = m.node_id[i];
      const uint32_t na = m.n_al[node];
      uint32_t head = 1;
      for (uint32_t k = 0; k < na; ++k) {
        if (m.row_of[m.al[static_cast<uint64_t>(node) * kMaxAligned + k]] < i) head = 0;
      }
      head_flag[i] = head;
    }
  });
  x.scan(head_flag + 1, R);
  x.run([&](uint32_t tid, uint32_t nt) {
    for (uint32_t i = 1 + tid; i <= R; i += nt) col_of[m.node_id[i]] = head_flag[i] - 1;
    if (tid == 0) S->msa_cols = R ? head_flag[R] : 0;
  });
}

template <class X>
SVS_HD void dg_msa_rows(X& x, const WinMem& m, WinState* S, const uint32_t* col_of, const uint32_t* seq_len /*per merged sequence*/,
                        uint32_t n_rows, uint8_t* out) {
  const uint64_t W = S->msa_cols;
  x.run([&](uint32_t tid, uint32_t nt) {
    const uint64_t total = W * n_rows;
    for (uint64_t k = tid; k < total; k += nt) out[k] = '-';
  });
  x.run([&](uint32_t tid, uint32_t nt) {
    uint64_t off = 0;
    for (uint32_t r = 0; r < n_rows; ++r) {
      const uint32_t len = seq_len[r];
      for (uint32_t p = tid; p < len; p += nt) {
        const uint32_t node = static_cast<uint32_t>(m.path_node[off + p]);
        out[static_cast<uint64_t>(r) * W + col_of[node]] = m.letter[node];
      }
      off += len;
    }
  });
}

// ---------------------------------------------------------------------------------------------
// Heaviest-bundle consensus (one thread): per node the heaviest in-edge (ties: the tail with
// the larger-or-equal running score, i.e. the later edge wins), running score = edge weight +
// score of the chosen tail; the best-scoring node is extended to a sink by branch completion,
// then traced back.  Needs the export with weights.  score / pred are indexed by ROW.
// Returns the consensus length; letters are written to out[0..len) (cap >= R).
SVS_HD uint32_t dg_consensus_serial(const WinMem& m, WinState* S, int32_t* score, int32_t* pred, uint8_t* out) {
  const uint32_t R = S->nv;
  if (R == 0) return 0;
  score[0] = -1; pred[0] = -1;
  int32_t best = -1;
  auto relax = [&](uint32_t i, bool skip_dead) {
    score[i] = -1;
    pred[i] = -1;
    if (m.preds[m.pred_off[i]] != 0) {
      for (uint32_t k = m.pred_off[i]; k < m.pred_off[i + 1]; ++k) {
        const int32_t t = static_cast<int32_t>(m.preds[k]);
        if (skip_dead && score[t] == -1) continue;
        const int32_t w = m.pred_w[k];
        if (score[i] < w || (score[i] == w && score[pred[i]] <= score[t])) {
          score[i] = w;
          pred[i] = t;
        }
      }
    }
    if (pred[i] >= 0) score[i] += score[pred[i]];
  };
  for (uint32_t i = 1; i <= R; ++i) {
    relax(i, false);
    if (best < 0 || score[best] < score[i]) best = static_cast<int32_t>(i);
  }
  // branch completion while the best node is not a sink
  while (!(m.r_flags[best] & kFlagSink)) {
    const uint32_t start = static_cast<uint32_t>(best);
    // competing tails of every head of `start` lose their score
    for (uint32_t h = start + 1; h <= R; ++h) {
      bool is_head = false;
      for (uint32_t k = m.pred_off[h]; k < m.pred_off[h + 1]; ++k) is_head |= (m.preds[k] == start);
      if (!is_head) continue;
      for (uint32_t k = m.pred_off[h]; k < m.pred_off[h + 1]; ++k) {
        if (m.preds[k] != start) score[m.preds[k]] = -1;
      }
    }
    int32_t nb = -1;
    for (uint32_t i = start + 1; i <= R; ++i) {
      relax(i, true);
      if (nb < 0 || score[nb] < score[i]) nb = static_cast<int32_t>(i);
    }
    best = nb;
  }
  uint32_t n = 0;
  for (int32_t v = best; v > 0; v = pred[v]) out[n++] = m.r_letter[v];
  for (uint32_t a = 0, b = n ? n - 1 : 0; a < b; ++a, --b) { const uint8_t t = out[a]; out[a] = out[b]; out[b] = t; }
  return n;
}

}  // namespace svs
