// ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported, linked or executed by the product
// path (svscope_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may use it, and only as the checker / the CPU baseline.
//
// PARITY UNPINNED: this is a scalar CPU restatement of the partial-order-alignment routine
// that the reference calls as `spoa.poa(sequences, 1)` (reference call sites:
// src/DataScanner.py:142,206,213 and src/DecisionMaker.py:160,171).  The arithmetic lives in
// the third-party wheel pyspoa 0.2.1 (README.md:19; wraps rvaser/spoa 4.x), which is neither
// vendored under /root/reference nor installed in this image, and the reference ships no
// tests or golden vectors for it.  What follows restates the published
// algorithm of spoa (Graph::AddAlignment / TopologicalSort / GenerateMultipleSequenceAlignment
// / GenerateConsensus and the scalar "sisd" convex-gap engine) from the survey's Appendix B.
// Tie-break precedence, node-id order, in-edge order and the DFS visiting order are the
// parts kernels must reproduce, so they are kept in ONE place here.
//
// Two engines share the graph code and ONE traceback (trace_back): the flat five-matrix engine (Engine) and a
// bounded-memory row-checkpoint engine (BlockedEngine, spo_set_blocked) for windows whose matrices do not fit
// in memory; tests/test_oracle_golden.py holds them equal, read after read.
//
// Build: see oracle/Makefile (g++ -O3 -shared -fPIC).  Interface: plain C, used via ctypes
// from oracle/oracle.py.

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

namespace {

constexpr int32_t kNegInf = std::numeric_limits<int32_t>::min() + 1024;

enum AlnType { kSW = 0, kNW = 1, kOV = 2 };
enum GapMode { kLinear = 0, kAffine = 1, kConvex = 2 };

struct Edge;

struct Node {
  uint32_t id;
  uint32_t code;
  std::vector<Edge*> in;       // creation order
  std::vector<Edge*> out;      // creation order
  std::vector<Node*> aligned;  // mutually aligned alternatives (one MSA column)
};

struct Edge {
  Node* tail;
  Node* head;
  std::vector<uint32_t> labels;  // sequence indices that traverse the edge
  int64_t weight;
};

using Alignment = std::vector<std::pair<int32_t, int32_t>>;  // (node id | -1, read pos | -1)

struct Graph {
  uint32_t num_codes = 0;
  int32_t coder[256];
  int32_t decoder[256];
  std::vector<Node*> sequences;  // first node of every added (non-empty) sequence
  std::vector<std::unique_ptr<Node>> nodes;
  std::vector<std::unique_ptr<Edge>> edges;
  std::vector<Node*> rank_to_node;
  std::vector<Node*> consensus;

  Graph() {
    std::fill(coder, coder + 256, -1);
    std::fill(decoder, decoder + 256, -1);
  }

  Node* add_node(uint32_t code) {
    nodes.emplace_back(new Node{static_cast<uint32_t>(nodes.size()), code, {}, {}, {}});
    return nodes.back().get();
  }

  // An existing tail->head edge is augmented; otherwise a new edge is appended to both
  // adjacency lists (this fixes in-edge order = order of first traversal).
  void add_edge(Node* tail, Node* head, int64_t w) {
    const uint32_t label = static_cast<uint32_t>(sequences.size());
    for (Edge* e : tail->out) {
      if (e->head == head) {
        e->labels.push_back(label);
        e->weight += w;
        return;
      }
    }
    edges.emplace_back(new Edge{tail, head, {label}, w});
    tail->out.push_back(edges.back().get());
    head->in.push_back(edges.back().get());
  }

  // Fresh chain for seq[begin,end); returns its first node (nullptr when empty).
  Node* add_chain(const char* seq, const std::vector<uint32_t>& w, uint32_t begin, uint32_t end) {
    if (begin == end) return nullptr;
    Node* prev = nullptr;
    for (uint32_t i = begin; i < end; ++i) {
      Node* cur = add_node(coder[static_cast<uint8_t>(seq[i])]);
      if (prev) add_edge(prev, cur, static_cast<int64_t>(w[i - 1]) + w[i]);
      prev = cur;
    }
    return nodes[nodes.size() - (end - begin)].get();
  }

  void add_alignment(const Alignment& aln, const char* seq, uint32_t len) {
    if (len == 0) return;  // empty sequences are ignored (no MSA row)
    std::vector<uint32_t> w(len, 1);
    for (uint32_t i = 0; i < len; ++i) {
      uint8_t ch = static_cast<uint8_t>(seq[i]);
      if (coder[ch] == -1) {
        coder[ch] = num_codes;
        decoder[num_codes++] = ch;
      }
    }
    if (aln.empty()) {
      sequences.push_back(add_chain(seq, w, 0, len));
      topological_sort();
      return;
    }
    std::vector<uint32_t> valid;
    for (const auto& p : aln) {
      if (p.second != -1) {
        if (p.second < 0 || p.second >= static_cast<int32_t>(len))
          throw std::invalid_argument("alignment position out of range");
        valid.push_back(p.second);
      }
    }
    if (valid.empty()) throw std::invalid_argument("alignment has no sequence positions");

    // unaligned prefix and suffix become fresh chains, created BEFORE the aligned part
    Node* begin = add_chain(seq, w, 0, valid.front());
    Node* prev = begin ? nodes.back().get() : nullptr;
    Node* last = add_chain(seq, w, valid.back() + 1, len);

    for (const auto& p : aln) {
      if (p.second == -1) continue;
      uint32_t code = coder[static_cast<uint8_t>(seq[p.second])];
      Node* cur = nullptr;
      if (p.first == -1) {
        cur = add_node(code);
      } else {
        Node* at = nodes[p.first].get();
        if (at->code == code) {
          cur = at;
        } else {
          for (Node* a : at->aligned) {
            if (a->code == code) { cur = a; break; }
          }
          if (!cur) {
            cur = add_node(code);
            for (Node* a : at->aligned) {
              a->aligned.push_back(cur);
              cur->aligned.push_back(a);
            }
            at->aligned.push_back(cur);
            cur->aligned.push_back(at);
          }
        }
      }
      if (!begin) begin = cur;
      if (prev) add_edge(prev, cur, static_cast<int64_t>(w[p.second - 1]) + w[p.second]);
      prev = cur;
    }
    if (last) add_edge(prev, last, static_cast<int64_t>(w[valid.back()]) + w[valid.back() + 1]);
    sequences.push_back(begin);
    topological_sort();
  }

  // Iterative DFS over nodes in id order; in-edge tails first (stored order), then aligned
  // nodes; an aligned group is emitted as consecutive ranks led by the node that reached it.
  void topological_sort() {
    rank_to_node.clear();
    std::vector<uint8_t> mark(nodes.size(), 0);  // 0 new, 1 open, 2 done
    std::vector<uint8_t> ignored(nodes.size(), 0);
    std::vector<Node*> stack;
    for (const auto& root : nodes) {
      if (mark[root->id] != 0) continue;
      stack.push_back(root.get());
      while (!stack.empty()) {
        Node* cur = stack.back();
        bool ready = true;
        if (mark[cur->id] != 2) {
          for (Edge* e : cur->in) {
            if (mark[e->tail->id] != 2) {
              stack.push_back(e->tail);
              ready = false;
            }
          }
          if (!ignored[cur->id]) {
            for (Node* a : cur->aligned) {
              if (mark[a->id] != 2) {
                stack.push_back(a);
                ignored[a->id] = 1;
                ready = false;
              }
            }
          }
          if (ready) {
            mark[cur->id] = 2;
            if (!ignored[cur->id]) {
              rank_to_node.push_back(cur);
              for (Node* a : cur->aligned) rank_to_node.push_back(a);
            }
          } else {
            mark[cur->id] = 1;
          }
        }
        if (ready) stack.pop_back();
      }
    }
  }

  Node* successor(const Node* n, uint32_t label) const {
    for (Edge* e : n->out) {
      if (std::find(e->labels.begin(), e->labels.end(), label) != e->labels.end()) return e->head;
    }
    return nullptr;
  }

  std::vector<uint32_t> msa_columns(uint32_t* ncols) const {
    std::vector<uint32_t> col(nodes.size());
    uint32_t j = 0;
    for (uint32_t i = 0; i < rank_to_node.size(); ++i, ++j) {
      Node* n = rank_to_node[i];
      col[n->id] = j;
      for (Node* a : n->aligned) {
        col[a->id] = j;
        ++i;
      }
    }
    *ncols = j;
    return col;
  }

  std::vector<std::string> msa() const {
    uint32_t ncols = 0;
    auto col = msa_columns(&ncols);
    std::vector<std::string> rows;
    for (uint32_t s = 0; s < sequences.size(); ++s) {
      std::string row(ncols, '-');
      Node* n = sequences[s];
      while (n) {
        row[col[n->id]] = static_cast<char>(decoder[n->code]);
        n = successor(n, s);
      }
      rows.push_back(std::move(row));
    }
    return rows;
  }

  Node* branch_completion(uint32_t rank, std::vector<int64_t>& score, std::vector<Node*>& pred) {
    Node* start = rank_to_node[rank];
    for (Edge* o : start->out) {
      for (Edge* i : o->head->in) {
        if (i->tail != start) score[i->tail->id] = -1;
      }
    }
    Node* best = nullptr;
    for (uint32_t r = rank + 1; r < rank_to_node.size(); ++r) {
      Node* n = rank_to_node[r];
      score[n->id] = -1;
      pred[n->id] = nullptr;
      for (Edge* e : n->in) {
        if (score[e->tail->id] == -1) continue;
        if (score[n->id] < e->weight ||
            (score[n->id] == e->weight && score[pred[n->id]->id] <= score[e->tail->id])) {
          score[n->id] = e->weight;
          pred[n->id] = e->tail;
        }
      }
      if (pred[n->id]) score[n->id] += score[pred[n->id]->id];
      if (!best || score[best->id] < score[n->id]) best = n;
    }
    return best;
  }

  std::string generate_consensus() {
    consensus.clear();
    if (rank_to_node.empty()) return std::string();
    std::vector<int64_t> score(nodes.size(), -1);
    std::vector<Node*> pred(nodes.size(), nullptr);
    Node* best = nullptr;
    for (Node* n : rank_to_node) {
      for (Edge* e : n->in) {
        if (score[n->id] < e->weight ||
            (score[n->id] == e->weight && score[pred[n->id]->id] <= score[e->tail->id])) {
          score[n->id] = e->weight;
          pred[n->id] = e->tail;
        }
      }
      if (pred[n->id]) score[n->id] += score[pred[n->id]->id];
      if (!best || score[best->id] < score[n->id]) best = n;
    }
    if (!best->out.empty()) {
      std::vector<uint32_t> rank_of(nodes.size(), 0);
      for (uint32_t r = 0; r < rank_to_node.size(); ++r) rank_of[rank_to_node[r]->id] = r;
      while (!best->out.empty()) best = branch_completion(rank_of[best->id], score, pred);
    }
    while (pred[best->id]) {
      consensus.push_back(best);
      best = pred[best->id];
    }
    consensus.push_back(best);
    std::reverse(consensus.begin(), consensus.end());
    std::string s;
    for (Node* n : consensus) s.push_back(static_cast<char>(decoder[n->code]));
    return s;
  }
};

// Matrix access of the flat five-matrix engine.
struct FlatMatrices {
  const int32_t *Hm, *Fm, *Em, *Om, *Qm;
  uint64_t W;
  void ensure(uint64_t) {}
  int32_t H(uint64_t i, uint64_t j) const { return Hm[i * W + j]; }
  int32_t F(uint64_t i, uint64_t j) const { return Fm[i * W + j]; }
  int32_t E(uint64_t i, uint64_t j) const { return Em[i * W + j]; }
  int32_t O(uint64_t i, uint64_t j) const { return Om[i * W + j]; }
  int32_t Q(uint64_t i, uint64_t j) const { return Qm[i * W + j]; }
};

// Traceback by ordered equality tests on the filled matrices: the ONE statement of the tie-break
// precedence (diagonal over the in-edges in stored order, then vertical, then horizontal; gap
// extension walks).  M gives the matrix values; M.ensure(i) is called before row i becomes the
// row the walk stands on (a no-op for flat matrices; the row-checkpoint engine recomputes the
// block of rows that holds i).
template <class Mat>
Alignment trace_back(Mat& M, AlnType type, int32_t e, int32_t g, int32_t q, int32_t c,
                     const std::vector<int32_t>& profile, uint64_t W, const std::vector<Node*>& rn,
                     const std::vector<uint32_t>& rank_of, uint32_t best_i, uint32_t best_j) {
  Alignment aln;
  uint64_t i = best_i, j = best_j;
  auto going = [&]() {
    M.ensure(i);
    if (type == kSW) return M.H(i, j) != 0;
    if (type == kNW) return !(i == 0 && j == 0);
    return !(i == 0 || j == 0);
  };
  uint64_t prev_i = 0, prev_j = 0;
  while (going()) {
    const int32_t Hij = M.H(i, j);
    bool found = false, ext_left = false, ext_up = false;
    if (i != 0 && j != 0) {
      Node* node = rn[i - 1];
      const int32_t s = profile[node->code * W + j];
      uint64_t pi = node->in.empty() ? 0 : rank_of[node->in[0]->tail->id] + 1;
      if (Hij == M.H(pi, j - 1) + s) {
        prev_i = pi; prev_j = j - 1; found = true;
      } else {
        for (size_t p = 1; p < node->in.size(); ++p) {
          pi = rank_of[node->in[p]->tail->id] + 1;
          if (Hij == M.H(pi, j - 1) + s) {
            prev_i = pi; prev_j = j - 1; found = true;
            break;
          }
        }
      }
    }
    if (!found && i != 0) {
      Node* node = rn[i - 1];
      auto vertical = [&](uint64_t pi) {
        return (ext_up |= (Hij == M.F(pi, j) + e)) || Hij == M.H(pi, j) + g ||
               (ext_up |= (Hij == M.O(pi, j) + c)) || Hij == M.H(pi, j) + q;
      };
      uint64_t pi = node->in.empty() ? 0 : rank_of[node->in[0]->tail->id] + 1;
      if (vertical(pi)) {
        prev_i = pi; prev_j = j; found = true;
      } else {
        for (size_t p = 1; p < node->in.size(); ++p) {
          pi = rank_of[node->in[p]->tail->id] + 1;
          if (vertical(pi)) {
            prev_i = pi; prev_j = j; found = true;
            break;
          }
        }
      }
    }
    if (!found && j != 0) {
      if ((ext_left |= (Hij == M.E(i, j - 1) + e)) || Hij == M.H(i, j - 1) + g ||
          (ext_left |= (Hij == M.Q(i, j - 1) + c)) || Hij == M.H(i, j - 1) + q) {
        prev_i = i; prev_j = j - 1; found = true;
      }
    }
    aln.emplace_back(i == prev_i ? -1 : static_cast<int32_t>(rn[i - 1]->id),
                     j == prev_j ? -1 : static_cast<int32_t>(j - 1));
    i = prev_i;
    j = prev_j;
    if (ext_left) {
      M.ensure(i);
      while (true) {
        aln.emplace_back(-1, static_cast<int32_t>(j - 1));
        --j;
        if (M.E(i, j) + e != M.E(i, j + 1) && M.Q(i, j) + c != M.Q(i, j + 1)) break;
      }
    } else if (ext_up) {
      while (true) {
        M.ensure(i);
        bool stop = true;
        prev_i = 0;
        for (Edge* ed : rn[i - 1]->in) {
          uint64_t pi = rank_of[ed->tail->id] + 1;
          if ((stop = (M.F(i, j) == M.H(pi, j) + g)) || M.F(i, j) == M.F(pi, j) + e ||
              (stop = (M.O(i, j) == M.H(pi, j) + q)) || M.O(i, j) == M.O(pi, j) + c) {
            prev_i = pi;
            break;
          }
        }
        aln.emplace_back(static_cast<int32_t>(rn[i - 1]->id), -1);
        i = prev_i;
        if (stop || i == 0) break;
      }
    }
  }
  std::reverse(aln.begin(), aln.end());
  return aln;
}

struct Engine {
  AlnType type;
  GapMode mode;
  int32_t m, n, g, e, q, c;
  std::vector<uint32_t> rank_of;
  std::vector<int32_t> profile;  // num_codes x width
  // score matrices: uninitialised storage, grown geometrically without copying
  std::unique_ptr<int32_t[]> Hs, Fs, Es, Os, Qs;
  int32_t *H = nullptr, *F = nullptr, *E = nullptr, *O = nullptr, *Q = nullptr;
  uint64_t capacity = 0;
  int64_t last_cells = 0;
  int32_t last_score = 0;

  Engine(int type_, int m_, int n_, int g_, int e_, int q_, int c_)
      : type(static_cast<AlnType>(type_)), m(m_), n(n_), g(g_), e(e_), q(q_), c(c_) {
    if (g >= e) mode = kLinear;
    else if (g <= q || e >= c) mode = kAffine;
    else mode = kConvex;
    if (mode != kConvex)
      throw std::invalid_argument("oracle restates the convex (two-piece) gap mode only");
  }

  void init(const char* seq, uint32_t len, const Graph& gr) {
    const uint64_t W = static_cast<uint64_t>(len) + 1;
    const uint64_t Hh = gr.nodes.size() + 1;
    const uint64_t cells = W * Hh;
    if (capacity < cells) {
      capacity = cells + cells / 2;
      Hs.reset(); Fs.reset(); Es.reset(); Os.reset(); Qs.reset();
      Hs.reset(new int32_t[capacity]); Fs.reset(new int32_t[capacity]);
      Es.reset(new int32_t[capacity]); Os.reset(new int32_t[capacity]);
      Qs.reset(new int32_t[capacity]);
      H = Hs.get(); F = Fs.get(); E = Es.get(); O = Os.get(); Q = Qs.get();
    }
    if (profile.size() < gr.num_codes * W) profile.resize(gr.num_codes * W);
    if (rank_of.size() < gr.nodes.size()) rank_of.resize(gr.nodes.size());
    for (uint32_t k = 0; k < gr.num_codes; ++k) {
      char ch = static_cast<char>(gr.decoder[k]);
      profile[k * W] = 0;
      for (uint32_t j = 0; j < len; ++j) profile[k * W + j + 1] = (ch == seq[j]) ? m : n;
    }
    const auto& rn = gr.rank_to_node;
    for (uint32_t r = 0; r < rn.size(); ++r) rank_of[rn[r]->id] = r;

    // second gap piece (O vertical, Q horizontal)
    O[0] = 0; Q[0] = 0;
    for (uint64_t j = 1; j < W; ++j) { O[j] = kNegInf; Q[j] = q + static_cast<int32_t>(j - 1) * c; }
    for (uint64_t i = 1; i < Hh; ++i) {
      const auto& in = rn[i - 1]->in;
      int32_t pen = in.empty() ? q - c : kNegInf;
      for (Edge* ed : in) pen = std::max(pen, O[(rank_of[ed->tail->id] + 1) * W]);
      O[i * W] = pen + c;
      Q[i * W] = kNegInf;
    }
    // first gap piece (F vertical, E horizontal)
    F[0] = 0; E[0] = 0;
    for (uint64_t j = 1; j < W; ++j) { F[j] = kNegInf; E[j] = g + static_cast<int32_t>(j - 1) * e; }
    for (uint64_t i = 1; i < Hh; ++i) {
      const auto& in = rn[i - 1]->in;
      int32_t pen = in.empty() ? g - e : kNegInf;
      for (Edge* ed : in) pen = std::max(pen, F[(rank_of[ed->tail->id] + 1) * W]);
      F[i * W] = pen + e;
      E[i * W] = kNegInf;
    }
    H[0] = 0;
    switch (type) {
      case kSW:
        for (uint64_t j = 1; j < W; ++j) H[j] = 0;
        for (uint64_t i = 1; i < Hh; ++i) H[i * W] = 0;
        break;
      case kNW:
        for (uint64_t j = 1; j < W; ++j) H[j] = std::max(Q[j], E[j]);
        for (uint64_t i = 1; i < Hh; ++i) H[i * W] = std::max(O[i * W], F[i * W]);
        break;
      case kOV:
        for (uint64_t j = 1; j < W; ++j) H[j] = std::max(Q[j], E[j]);
        for (uint64_t i = 1; i < Hh; ++i) H[i * W] = 0;
        break;
    }
  }

  Alignment align(const char* seq, uint32_t len, const Graph& gr) {
    last_cells = 0;
    if (gr.nodes.empty() || len == 0) return Alignment();
    {  // worst-case score must stay above the sentinel
      int64_t worst = static_cast<int64_t>(std::min(std::min(g, q), std::min(e, c))) *
                      (static_cast<int64_t>(len) + static_cast<int64_t>(gr.nodes.size()) + 2);
      if (worst < kNegInf) throw std::invalid_argument("possible score overflow");
    }
    init(seq, len, gr);
    const uint64_t W = static_cast<uint64_t>(len) + 1;
    const auto& rn = gr.rank_to_node;
    last_cells = static_cast<int64_t>(W) * static_cast<int64_t>(rn.size() + 1);

    int32_t best = (type == kSW) ? 0 : kNegInf;
    uint32_t best_i = 0, best_j = 0;

    // local copies so that stores into the score rows cannot alias the parameters
    const int32_t g_ = g, e_ = e, q_ = q, c_ = c;
    const AlnType type_ = type;
    for (Node* node : rn) {
      const int32_t* __restrict prof = &profile[node->code * W];
      const uint64_t i = rank_of[node->id] + 1;
      uint64_t pi = node->in.empty() ? 0 : rank_of[node->in[0]->tail->id] + 1;
      int32_t* __restrict Hr = H + i * W;
      int32_t* __restrict Fr = F + i * W;
      int32_t* __restrict Or = O + i * W;
      {
        const int32_t* __restrict Hp = H + pi * W;
        const int32_t* __restrict Fp = F + pi * W;
        const int32_t* __restrict Op = O + pi * W;
        for (uint64_t j = 1; j < W; ++j) {
          Fr[j] = std::max(Hp[j] + g_, Fp[j] + e_);
          Or[j] = std::max(Hp[j] + q_, Op[j] + c_);
          Hr[j] = Hp[j - 1] + prof[j];
        }
      }
      for (size_t p = 1; p < node->in.size(); ++p) {
        pi = rank_of[node->in[p]->tail->id] + 1;
        const int32_t* __restrict Hp = H + pi * W;
        const int32_t* __restrict Fp = F + pi * W;
        const int32_t* __restrict Op = O + pi * W;
        for (uint64_t j = 1; j < W; ++j) {
          Fr[j] = std::max(Fr[j], std::max(Hp[j] + g_, Fp[j] + e_));
          Or[j] = std::max(Or[j], std::max(Hp[j] + q_, Op[j] + c_));
          Hr[j] = std::max(Hr[j], Hp[j - 1] + prof[j]);
        }
      }
      int32_t* __restrict Er = E + i * W;
      int32_t* __restrict Qr = Q + i * W;
      const bool sink = node->out.empty();
      // horizontal pieces: E[j] = max(H[j-1]+g, E[j-1]+e), Q likewise, then the 5-way max
      int32_t hl = Hr[0], el = Er[0], ql = Qr[0];
      for (uint64_t j = 1; j < W; ++j) {
        el = std::max(hl + g_, el + e_);
        ql = std::max(hl + q_, ql + c_);
        hl = std::max(Hr[j], std::max(std::max(Fr[j], el), std::max(Or[j], ql)));
        if (type_ == kSW) hl = std::max(hl, 0);
        Er[j] = el;
        Qr[j] = ql;
        Hr[j] = hl;
      }
      // end-cell candidates (first strictly greater wins, rank order then column order)
      if (type_ == kSW) {
        for (uint64_t j = 1; j < W; ++j) {
          if (best < Hr[j]) { best = Hr[j]; best_i = static_cast<uint32_t>(i); best_j = static_cast<uint32_t>(j); }
        }
      } else if (type_ == kNW) {
        if (sink && W > 1 && best < Hr[W - 1]) {
          best = Hr[W - 1]; best_i = static_cast<uint32_t>(i); best_j = static_cast<uint32_t>(W - 1);
        }
      } else if (sink) {
        for (uint64_t j = 1; j < W; ++j) {
          if (best < Hr[j]) { best = Hr[j]; best_i = static_cast<uint32_t>(i); best_j = static_cast<uint32_t>(j); }
        }
      }
    }
    if (best_i == 0 && best_j == 0) return Alignment();
    last_score = best;

    // traceback by ordered equality tests (shared with the row-checkpoint engine below)
    FlatMatrices mats{H, F, E, O, Q, W};
    Alignment aln = trace_back(mats, type, e, g, q, c, profile, W, rn, rank_of, best_i, best_j);
    return aln;
  }
};

// Row-checkpoint engine: the same recurrences and the same traceback (trace_back above), with
// bounded memory.  Rows are filled in blocks of `block_rows` ranks; a block keeps its five matrices
// only while it is the current block.  A row with a successor in a LATER block is kept (H, F, O)
// for the rest of the alignment, so every block can be recomputed from kept rows alone; the
// traceback recomputes the block that holds the row it stands on (each block at most once, since
// the walk only moves to lower ranks).  Used for windows whose five full matrices do not fit
// in memory (BASELINE configs[2] at full size: ~150k x 20k cells); tests/test_oracle_blocked.py
// checks it against the flat engine, alignment by alignment, with blocks of a few rows.
struct BlockedEngine {
  int32_t m, n, g, e, q, c;
  int64_t block_rows_opt = -1;   // -1: sized for ~2 GB per block
  std::vector<uint32_t> rank_of;
  std::vector<int32_t> profile;
  std::vector<int32_t> row0[5];        // H F E O Q of the virtual source row
  std::vector<int32_t> H0, F0, O0;     // column 0 of every row
  std::vector<int32_t> blk[5];         // H F E O Q of the current block
  uint64_t blk_first = 0, blk_count = 0, W = 0, B = 0;
  std::unordered_map<uint64_t, std::unique_ptr<int32_t[]>> kept;   // row -> H, F, O (3W)
  const Graph* gr = nullptr;
  int64_t last_cells = 0, kept_rows = 0, recomputed_blocks = 0;
  int32_t last_score = 0;
  int32_t best = 0;
  uint32_t best_i = 0, best_j = 0;

  BlockedEngine(int m_, int n_, int g_, int e_, int q_, int c_) : m(m_), n(n_), g(g_), e(e_), q(q_), c(c_) {}

  struct RowPtr { const int32_t *H, *F, *O; };
  RowPtr pred_row(uint64_t pi) const {
    if (pi == 0) return {row0[0].data(), row0[1].data(), row0[3].data()};
    if (pi >= blk_first && pi < blk_first + blk_count) {
      const uint64_t o = (pi - blk_first) * W;
      return {blk[0].data() + o, blk[1].data() + o, blk[3].data() + o};
    }
    auto it = kept.find(pi);
    if (it == kept.end()) throw std::logic_error("row-checkpoint engine: predecessor row neither in the block nor kept");
    return {it->second.get(), it->second.get() + W, it->second.get() + 2 * W};
  }

  // fills rows [first, first+count); `forward` = first visit (keeps rows, tracks the end cell)
  void compute_block(uint64_t first, uint64_t count, bool forward) {
    blk_first = first;
    blk_count = 0;          // rows become visible to pred_row one by one
    const auto& rn = gr->rank_to_node;
    for (uint64_t i = first; i < first + count; ++i) {
      Node* node = rn[i - 1];
      const int32_t* __restrict prof = &profile[node->code * W];
      const uint64_t o = (i - first) * W;
      int32_t* __restrict Hr = blk[0].data() + o;
      int32_t* __restrict Fr = blk[1].data() + o;
      int32_t* __restrict Er = blk[2].data() + o;
      int32_t* __restrict Or = blk[3].data() + o;
      int32_t* __restrict Qr = blk[4].data() + o;
      Hr[0] = H0[i]; Fr[0] = F0[i]; Or[0] = O0[i]; Er[0] = kNegInf; Qr[0] = kNegInf;
      const size_t np = node->in.size();
      for (size_t p = 0; p < std::max<size_t>(np, 1); ++p) {
        const uint64_t pi = np == 0 ? 0 : rank_of[node->in[p]->tail->id] + 1;
        const RowPtr pr = pred_row(pi);
        if (p == 0) {
          for (uint64_t j = 1; j < W; ++j) {
            Fr[j] = std::max(pr.H[j] + g, pr.F[j] + e);
            Or[j] = std::max(pr.H[j] + q, pr.O[j] + c);
            Hr[j] = pr.H[j - 1] + prof[j];
          }
        } else {
          for (uint64_t j = 1; j < W; ++j) {
            Fr[j] = std::max(Fr[j], std::max(pr.H[j] + g, pr.F[j] + e));
            Or[j] = std::max(Or[j], std::max(pr.H[j] + q, pr.O[j] + c));
            Hr[j] = std::max(Hr[j], pr.H[j - 1] + prof[j]);
          }
        }
      }
      int32_t hl = Hr[0], el = Er[0], ql = Qr[0];
      for (uint64_t j = 1; j < W; ++j) {
        el = std::max(hl + g, el + e);
        ql = std::max(hl + q, ql + c);
        hl = std::max(Hr[j], std::max(std::max(Fr[j], el), std::max(Or[j], ql)));
        Er[j] = el; Qr[j] = ql; Hr[j] = hl;
      }
      blk_count = i - first + 1;
      if (forward && node->out.empty() && W > 1 && best < Hr[W - 1]) {
        best = Hr[W - 1]; best_i = static_cast<uint32_t>(i); best_j = static_cast<uint32_t>(W - 1);
      }
    }
    if (!forward) { ++recomputed_blocks; return; }
    for (uint64_t i = first; i < first + count; ++i) {
      bool later = false;
      for (Edge* ed : rn[i - 1]->out) later |= rank_of[ed->head->id] + 1 >= first + count;
      if (!later) continue;
      std::unique_ptr<int32_t[]> row(new int32_t[3 * W]);
      const uint64_t o = (i - first) * W;
      std::memcpy(row.get(), blk[0].data() + o, W * 4);
      std::memcpy(row.get() + W, blk[1].data() + o, W * 4);
      std::memcpy(row.get() + 2 * W, blk[3].data() + o, W * 4);
      kept.emplace(i, std::move(row));
      ++kept_rows;
      if (static_cast<uint64_t>(kept_rows) * 3 * W * 4 > (40ull << 30)) throw std::runtime_error("row-checkpoint engine: kept rows exceed 40 GB");
    }
  }

  // trace_back's view
  void ensure(uint64_t i) {
    if (i == 0 || (i >= blk_first && i < blk_first + blk_count)) return;
    const uint64_t first = 1 + (i - 1) / B * B;
    compute_block(first, std::min<uint64_t>(B, gr->nodes.size() + 1 - first), false);
  }
  const int32_t* own(int k, uint64_t i) const {
    if (i == 0) return row0[k].data();
    if (i >= blk_first && i < blk_first + blk_count) return blk[k].data() + (i - blk_first) * W;
    return nullptr;
  }
  int32_t H(uint64_t i, uint64_t j) const { const int32_t* r = own(0, i); return r ? r[j] : pred_row(i).H[j]; }
  int32_t F(uint64_t i, uint64_t j) const { const int32_t* r = own(1, i); return r ? r[j] : pred_row(i).F[j]; }
  int32_t O(uint64_t i, uint64_t j) const { const int32_t* r = own(3, i); return r ? r[j] : pred_row(i).O[j]; }
  int32_t E(uint64_t i, uint64_t j) const {
    const int32_t* r = own(2, i);
    if (!r) throw std::logic_error("row-checkpoint engine: E of a row outside the current block");
    return r[j];
  }
  int32_t Q(uint64_t i, uint64_t j) const {
    const int32_t* r = own(4, i);
    if (!r) throw std::logic_error("row-checkpoint engine: Q of a row outside the current block");
    return r[j];
  }

  Alignment align(const char* seq, uint32_t len, const Graph& graph) {
    last_cells = 0; kept_rows = 0; recomputed_blocks = 0;
    if (graph.nodes.empty() || len == 0) return Alignment();
    if (!(g < e && !(g <= q || e >= c))) throw std::invalid_argument("oracle restates the convex (two-piece) gap mode only");
    {
      int64_t worst = static_cast<int64_t>(std::min(std::min(g, q), std::min(e, c))) *
                      (static_cast<int64_t>(len) + static_cast<int64_t>(graph.nodes.size()) + 2);
      if (worst < kNegInf) throw std::invalid_argument("possible score overflow");
    }
    gr = &graph;
    W = static_cast<uint64_t>(len) + 1;
    const uint64_t Hh = graph.nodes.size() + 1;
    const auto& rn = graph.rank_to_node;
    last_cells = static_cast<int64_t>(W * Hh);
    B = block_rows_opt > 0 ? static_cast<uint64_t>(block_rows_opt) : std::max<uint64_t>(64, (2ull << 30) / (20 * W));
    B = std::min(B, Hh);
    for (auto& v : blk) if (v.size() < B * W) { v.clear(); v.shrink_to_fit(); v.resize(B * W); }
    kept.clear();
    blk_first = 0; blk_count = 0;
    if (profile.size() < graph.num_codes * W) profile.resize(graph.num_codes * W);
    if (rank_of.size() < graph.nodes.size()) rank_of.resize(graph.nodes.size());
    for (uint32_t k = 0; k < graph.num_codes; ++k) {
      char ch = static_cast<char>(graph.decoder[k]);
      profile[k * W] = 0;
      for (uint32_t j = 0; j < len; ++j) profile[k * W + j + 1] = (ch == seq[j]) ? m : n;
    }
    for (uint32_t r = 0; r < rn.size(); ++r) rank_of[rn[r]->id] = r;
    // boundary conditions of the global alignment (Engine::init, kNW)
    for (auto& v : row0) v.assign(W, 0);
    for (uint64_t j = 1; j < W; ++j) {
      row0[3][j] = kNegInf; row0[4][j] = q + static_cast<int32_t>(j - 1) * c;
      row0[1][j] = kNegInf; row0[2][j] = g + static_cast<int32_t>(j - 1) * e;
      row0[0][j] = std::max(row0[4][j], row0[2][j]);
    }
    H0.assign(Hh, 0); F0.assign(Hh, 0); O0.assign(Hh, 0);
    for (uint64_t i = 1; i < Hh; ++i) {
      const auto& in = rn[i - 1]->in;
      int32_t po = in.empty() ? q - c : kNegInf, pf = in.empty() ? g - e : kNegInf;
      for (Edge* ed : in) {
        po = std::max(po, O0[rank_of[ed->tail->id] + 1]);
        pf = std::max(pf, F0[rank_of[ed->tail->id] + 1]);
      }
      O0[i] = po + c; F0[i] = pf + e; H0[i] = std::max(O0[i], F0[i]);
    }
    best = kNegInf; best_i = 0; best_j = 0;
    for (uint64_t first = 1; first < Hh; first += B) compute_block(first, std::min<uint64_t>(B, Hh - first), true);
    if (best_i == 0 && best_j == 0) return Alignment();
    last_score = best;
    Alignment aln = trace_back(*this, kNW, e, g, q, c, profile, W, rn, rank_of, best_i, best_j);
    kept.clear();
    return aln;
  }
};

struct Session {
  Engine engine;
  BlockedEngine blocked;
  bool use_blocked = false;
  Graph graph;
  Alignment last;
  std::string err;
  Session(int t, int m, int n, int g, int e, int q, int c) : engine(t, m, n, g, e, q, c), blocked(m, n, g, e, q, c) {}
  Alignment align(const char* seq, uint32_t len) {
    if (use_blocked) {
      if (engine.type != kNW) throw std::invalid_argument("row-checkpoint engine: global alignment only");
      Alignment a = blocked.align(seq, len, graph);
      engine.last_cells = blocked.last_cells;
      engine.last_score = blocked.last_score;
      return a;
    }
    return engine.align(seq, len, graph);
  }
};

}  // namespace

extern "C" {

void* spo_new(int algorithm, int m, int n, int g, int e, int q, int c) {
  try {
    return new Session(algorithm, m, n, g, e, q, c);
  } catch (...) {
    return nullptr;
  }
}

void spo_free(void* h) { delete static_cast<Session*>(h); }

// Align one sequence to the current graph and merge it.  Returns the alignment length
// (>= 0) or -1 on error.
int64_t spo_add(void* h, const char* seq, int64_t len) {
  Session* s = static_cast<Session*>(h);
  try {
    s->last = s->align(seq, static_cast<uint32_t>(len));
    s->graph.add_alignment(s->last, seq, static_cast<uint32_t>(len));
    return static_cast<int64_t>(s->last.size());
  } catch (const std::exception& ex) {
    s->err = ex.what();
    return -1;
  }
}

// Align only (no merge); result readable with spo_last_alignment.
int64_t spo_align_only(void* h, const char* seq, int64_t len) {
  Session* s = static_cast<Session*>(h);
  try {
    s->last = s->align(seq, static_cast<uint32_t>(len));
    return static_cast<int64_t>(s->last.size());
  } catch (const std::exception& ex) {
    s->err = ex.what();
    return -1;
  }
}

int64_t spo_last_alignment(void* h, int32_t* node_ids, int32_t* positions, int64_t cap) {
  Session* s = static_cast<Session*>(h);
  int64_t n = std::min<int64_t>(cap, static_cast<int64_t>(s->last.size()));
  for (int64_t k = 0; k < n; ++k) {
    node_ids[k] = s->last[k].first;
    positions[k] = s->last[k].second;
  }
  return static_cast<int64_t>(s->last.size());
}

// Row-checkpoint engine for the following alignments: block_rows > 0 rows per block, -1 sized for
// ~2 GB per block, 0 back to the flat five-matrix engine.
void spo_set_blocked(void* h, int64_t block_rows) {
  Session* s = static_cast<Session*>(h);
  s->use_blocked = block_rows != 0;
  s->blocked.block_rows_opt = block_rows;
}
int64_t spo_blocked_kept_rows(void* h) { return static_cast<Session*>(h)->blocked.kept_rows; }
int64_t spo_blocked_recomputed(void* h) { return static_cast<Session*>(h)->blocked.recomputed_blocks; }

int64_t spo_last_cells(void* h) { return static_cast<Session*>(h)->engine.last_cells; }
int32_t spo_last_score(void* h) { return static_cast<Session*>(h)->engine.last_score; }
const char* spo_error(void* h) { return static_cast<Session*>(h)->err.c_str(); }

int64_t spo_num_nodes(void* h) { return static_cast<int64_t>(static_cast<Session*>(h)->graph.nodes.size()); }
int64_t spo_num_edges(void* h) { return static_cast<int64_t>(static_cast<Session*>(h)->graph.edges.size()); }
int64_t spo_num_sequences(void* h) { return static_cast<int64_t>(static_cast<Session*>(h)->graph.sequences.size()); }

// Graph dump in rank order: for rank r, node id, letter, in-degree; in-edge tails (node ids)
// and weights are concatenated in stored order.  Arrays sized num_nodes / num_edges.
void spo_graph_dump(void* h, int32_t* rank_node, uint8_t* rank_letter, int32_t* rank_indeg,
                    int32_t* in_tail, int64_t* in_weight, int32_t* rank_naligned,
                    int32_t* rank_outdeg) {
  Session* s = static_cast<Session*>(h);
  int64_t k = 0;
  for (size_t r = 0; r < s->graph.rank_to_node.size(); ++r) {
    Node* n = s->graph.rank_to_node[r];
    rank_node[r] = static_cast<int32_t>(n->id);
    rank_letter[r] = static_cast<uint8_t>(s->graph.decoder[n->code]);
    rank_indeg[r] = static_cast<int32_t>(n->in.size());
    rank_naligned[r] = static_cast<int32_t>(n->aligned.size());
    rank_outdeg[r] = static_cast<int32_t>(n->out.size());
    for (Edge* e : n->in) {
      in_tail[k] = static_cast<int32_t>(e->tail->id);
      in_weight[k] = e->weight;
      ++k;
    }
  }
}

int64_t spo_consensus(void* h, char* out, int64_t cap) {
  Session* s = static_cast<Session*>(h);
  std::string c = s->graph.generate_consensus();
  int64_t n = std::min<int64_t>(cap, static_cast<int64_t>(c.size()));
  std::memcpy(out, c.data(), n);
  return static_cast<int64_t>(c.size());
}

void spo_msa_dims(void* h, int64_t* rows, int64_t* cols) {
  Session* s = static_cast<Session*>(h);
  uint32_t nc = 0;
  s->graph.msa_columns(&nc);
  *rows = static_cast<int64_t>(s->graph.sequences.size());
  *cols = nc;
}

// rows*cols characters, row-major, no terminators
void spo_msa(void* h, char* out) {
  Session* s = static_cast<Session*>(h);
  auto rows = s->graph.msa();
  size_t off = 0;
  for (const auto& r : rows) {
    std::memcpy(out + off, r.data(), r.size());
    off += r.size();
  }
}

}  // extern "C"
