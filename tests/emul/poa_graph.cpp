// Host-side partial-order graph (see poa_graph.h for the behavioural contract).
#include "poa_graph.h"

#include <algorithm>
#include <stdexcept>

#include "../../svscope_b200/csrc/poa_cell.h"

namespace svs {

uint32_t PoaGraph::add_node(uint8_t letter) {
  letter_.push_back(letter);
  in_.emplace_back();
  out_.emplace_back();
  aligned_.emplace_back();
  return static_cast<uint32_t>(letter_.size() - 1);
}

// Re-traversing an existing tail->head edge adds weight; a new edge is appended to the tail's
// out-list and the head's in-list, which fixes the in-edge order used by the DP tie-breaks.
void PoaGraph::add_edge(uint32_t tail, uint32_t head, int64_t w) {
  for (uint32_t e : out_[tail]) {
    if (head_[e] == head) {
      weight_[e] += w;
      return;
    }
  }
  const uint32_t e = static_cast<uint32_t>(tail_.size());
  tail_.push_back(tail);
  head_.push_back(head);
  weight_.push_back(w);
  out_[tail].push_back(e);
  in_[head].push_back(e);
}

// New chain for seq[begin,end); returns id of its first node or -1.
int64_t PoaGraph::add_chain(const uint8_t* seq, uint32_t begin, uint32_t end,
                            std::vector<uint32_t>* path) {
  if (begin == end) return -1;
  int64_t first = -1, prev = -1;
  for (uint32_t i = begin; i < end; ++i) {
    const uint32_t cur = add_node(seq[i]);
    if (first < 0) first = cur;
    if (prev >= 0) add_edge(static_cast<uint32_t>(prev), cur, 2);  // unit weights: 1 + 1
    prev = cur;
    path->push_back(cur);
  }
  return first;
}

void PoaGraph::add_alignment(const int32_t* pair_node, const int32_t* pair_pos, size_t n_pairs,
                             const uint8_t* seq, uint32_t len) {
  if (len == 0) return;
  std::vector<uint32_t> path;
  path.reserve(len);
  if (n_pairs == 0) {
    add_chain(seq, 0, len, &path);
    paths_.push_back(std::move(path));
    topological_sort();
    return;
  }
  int64_t first_pos = -1, last_pos = -1;
  for (size_t k = 0; k < n_pairs; ++k) {
    const int32_t p = pair_pos[k];
    if (p == -1) continue;
    if (p < 0 || p >= static_cast<int32_t>(len))
      throw std::invalid_argument("alignment position out of range");
    if (first_pos < 0) first_pos = p;
    last_pos = p;
  }
  if (first_pos < 0) throw std::invalid_argument("alignment carries no sequence position");

  // Unaligned head and tail of the read become fresh chains; both are created before the
  // aligned part so that they receive the lower node ids.
  std::vector<uint32_t> suffix_path;
  add_chain(seq, 0, static_cast<uint32_t>(first_pos), &path);
  int64_t prev = path.empty() ? -1 : static_cast<int64_t>(path.back());
  const int64_t suffix_first = add_chain(seq, static_cast<uint32_t>(last_pos) + 1, len, &suffix_path);

  for (size_t k = 0; k < n_pairs; ++k) {
    const int32_t p = pair_pos[k];
    if (p == -1) continue;
    const uint8_t ch = seq[p];
    const int32_t at = pair_node[k];
    uint32_t cur;
    if (at == -1) {
      cur = add_node(ch);
    } else if (letter_[at] == ch) {
      cur = static_cast<uint32_t>(at);
    } else {
      int64_t hit = -1;
      for (uint32_t a : aligned_[at]) {
        if (letter_[a] == ch) { hit = a; break; }
      }
      if (hit >= 0) {
        cur = static_cast<uint32_t>(hit);
      } else {
        cur = add_node(ch);
        // note: aligned_[at] must be re-read by index, add_node may have reallocated
        const std::vector<uint32_t> group = aligned_[at];
        for (uint32_t a : group) {
          aligned_[a].push_back(cur);
          aligned_[cur].push_back(a);
        }
        aligned_[at].push_back(cur);
        aligned_[cur].push_back(static_cast<uint32_t>(at));
      }
    }
    if (prev >= 0) add_edge(static_cast<uint32_t>(prev), cur, 2);
    prev = cur;
    path.push_back(cur);
  }
  if (suffix_first >= 0) add_edge(static_cast<uint32_t>(prev), static_cast<uint32_t>(suffix_first), 2);
  path.insert(path.end(), suffix_path.begin(), suffix_path.end());
  paths_.push_back(std::move(path));
  topological_sort();
}

// Depth-first emission over node ids in creation order.  A node is emitted once every in-edge
// tail and (unless it was itself reached as an aligned alternative) every aligned node has
// been emitted; the aligned alternatives follow it immediately.  The explicit stack with
// "push all unfinished, then re-examine" reproduces the visiting order exactly.
void PoaGraph::topological_sort() {
  const uint32_t n = num_nodes();
  rank_to_node_.clear();
  rank_to_node_.reserve(n);
  std::vector<uint8_t> state(n, 0);      // 0 unseen, 1 open, 2 emitted
  std::vector<uint8_t> as_aligned(n, 0); // reached through an aligned link
  std::vector<uint32_t> stack;
  stack.reserve(256);
  for (uint32_t root = 0; root < n; ++root) {
    if (state[root] != 0) continue;
    stack.push_back(root);
    while (!stack.empty()) {
      const uint32_t cur = stack.back();
      bool ready = true;
      if (state[cur] != 2) {
        for (uint32_t e : in_[cur]) {
          const uint32_t t = tail_[e];
          if (state[t] != 2) {
            stack.push_back(t);
            ready = false;
          }
        }
        if (!as_aligned[cur]) {
          for (uint32_t a : aligned_[cur]) {
            if (state[a] != 2) {
              stack.push_back(a);
              as_aligned[a] = 1;
              ready = false;
            }
          }
        }
        if (ready) {
          state[cur] = 2;
          if (!as_aligned[cur]) {
            rank_to_node_.push_back(cur);
            for (uint32_t a : aligned_[cur]) rank_to_node_.push_back(a);
          }
        } else {
          state[cur] = 1;
        }
      }
      if (ready) stack.pop_back();
    }
  }
}

void PoaGraph::export_ranked(const PoaScoring& sc, uint32_t ring_rows, RankedGraph* out) const {
  const uint32_t R = num_nodes();
  std::vector<uint32_t> row_of(R);
  for (uint32_t r = 0; r < R; ++r) row_of[rank_to_node_[r]] = r + 1;
  out->R = R;
  out->letter.assign(R + 1, 0);
  out->pred_off.assign(R + 2, 0);
  out->preds.clear();
  out->flags.assign(R + 1, 0);
  out->xslot.assign(R + 1, -1);
  out->h0.assign(R + 1, 0);
  out->col0code.assign(R + 1, 0);
  out->node_id.assign(R + 1, 0);
  out->max_indeg = 1;
  // column-0 scores follow the initialisation of the scalar engine: per gap piece the best
  // in-edge tail plus one extension, sources open a fresh gap.
  std::vector<int32_t> f0(R + 1, 0), o0(R + 1, 0);
  out->pred_off[0] = 0;
  out->pred_off[1] = 0;  // row 0 has no predecessors
  for (uint32_t i = 1; i <= R; ++i) {
    const uint32_t node = rank_to_node_[i - 1];
    out->letter[i] = letter_[node];
    out->node_id[i] = node;
    if (out_[node].empty()) out->flags[i] |= kFlagSink;
    const auto& in = in_[node];
    if (in.empty()) {
      out->preds.push_back(0);
      f0[i] = sc.g;
      o0[i] = sc.q;
    } else {
      int32_t bf = INT32_MIN, bo = INT32_MIN;
      for (uint32_t e : in) {
        const uint32_t p = row_of[tail_[e]];
        out->preds.push_back(p);
        bf = std::max(bf, f0[p]);
        bo = std::max(bo, o0[p]);
        if (i - p > ring_rows) out->flags[p] |= kFlagExport;
      }
      f0[i] = bf + sc.e;
      o0[i] = bo + sc.c;
      out->max_indeg = std::max<uint32_t>(out->max_indeg, static_cast<uint32_t>(in.size()));
    }
    out->pred_off[i + 1] = static_cast<uint32_t>(out->preds.size());
    out->h0[i] = std::max(f0[i], o0[i]);
  }
  // traceback codes of column 0 (no diagonal, no horizontal move is possible there)
  for (uint32_t i = 1; i <= R; ++i) {
    const uint32_t b = out->pred_off[i], e = out->pred_off[i + 1];
    const int32_t H = out->h0[i];
    uint32_t kH = 0, ext = 0;
    bool found = false;
    for (uint32_t k = b; k < e && !found; ++k) {
      const uint32_t p = out->preds[k];
      const int32_t Hp = out->h0[p], Fp = f0[p], Op = o0[p];  // row 0: all 0
      if (H == Fp + sc.e) { found = true; ext = 1; }
      else if (H == Hp + sc.g) { found = true; }
      else if (H == Op + sc.c) { found = true; ext = 1; }
      else if (H == Hp + sc.q) { found = true; }
      if (found) kH = k - b;
    }
    uint32_t kU = 0, stop = 1;
    const uint32_t node = rank_to_node_[i - 1];
    if (!in_[node].empty()) {
      bool hit = false;
      stop = 0;
      for (uint32_t k = b; k < e && !hit; ++k) {
        const uint32_t p = out->preds[k];
        if (f0[i] == out->h0[p] + sc.g) { hit = true; stop = 1; }
        else if (f0[i] == f0[p] + sc.e) { hit = true; stop = 0; }
        else if (o0[i] == out->h0[p] + sc.q) { hit = true; stop = 1; }
        else if (o0[i] == o0[p] + sc.c) { hit = true; stop = 0; }
        if (hit) kU = k - b;
      }
      if (!hit) { kU = kNoPred; stop = 0; }
    }
    out->col0code[i] = make_code(kMoveVert, ext, 0, stop, kH, kU);
  }
  uint32_t slot = 0;
  out->single_before.assign(R + 2, 0);
  uint32_t singles = 0;
  for (uint32_t i = 1; i <= R; ++i) {
    if (out->flags[i] & kFlagExport) out->xslot[i] = static_cast<int32_t>(slot++);
    out->single_before[i] = singles;
    if (out->pred_off[i + 1] - out->pred_off[i] == 1) {
      ++singles;
      if (out->preds[out->pred_off[i]] + 1 == i) out->flags[i] |= kFlagChain;
    }
  }
  out->single_before[R + 1] = singles;
  out->n_export = slot;
  // path-length intervals used by the pruning bounds (poa_cell.h cell_bound)
  out->depth.assign(4 * (static_cast<size_t>(R) + 1), 0);
  int32_t* dp = out->depth.data();
  for (uint32_t i = 1; i <= R; ++i) {
    int32_t lo = INT32_MAX, hi = 0;
    for (uint32_t k = out->pred_off[i]; k < out->pred_off[i + 1]; ++k) {
      const uint32_t p = out->preds[k];
      lo = std::min(lo, dp[4 * p]);
      hi = std::max(hi, dp[4 * p + 1]);
    }
    dp[4 * i] = lo + 1;
    dp[4 * i + 1] = hi + 1;
  }
  std::vector<uint8_t> seen(R + 1, 0);
  for (uint32_t i = R; i >= 1; --i) {
    for (uint32_t k = out->pred_off[i]; k < out->pred_off[i + 1]; ++k) {
      const uint32_t p = out->preds[k];
      if (p == 0) continue;
      const int32_t a = dp[4 * i + 2] + 1, b = dp[4 * i + 3] + 1;
      if (!seen[p]) { dp[4 * p + 2] = a; dp[4 * p + 3] = b; seen[p] = 1; }
      else { dp[4 * p + 2] = std::min(dp[4 * p + 2], a); dp[4 * p + 3] = std::max(dp[4 * p + 3], b); }
    }
  }
}

std::vector<std::string> PoaGraph::msa() const {
  const uint32_t n = num_nodes();
  std::vector<uint32_t> column(n, 0);
  uint32_t ncols = 0;
  for (uint32_t r = 0; r < rank_to_node_.size(); ++r, ++ncols) {
    const uint32_t node = rank_to_node_[r];
    column[node] = ncols;
    for (uint32_t a : aligned_[node]) {
      column[a] = ncols;
      ++r;
    }
  }
  std::vector<std::string> rows;
  rows.reserve(paths_.size());
  for (const auto& path : paths_) {
    std::string row(ncols, '-');
    for (uint32_t node : path) row[column[node]] = static_cast<char>(letter_[node]);
    rows.push_back(std::move(row));
  }
  return rows;
}

uint32_t PoaGraph::branch_completion(uint32_t rank, std::vector<int64_t>& score,
                                     std::vector<int64_t>& pred) const {
  const uint32_t start = rank_to_node_[rank];
  for (uint32_t oe : out_[start]) {
    for (uint32_t ie : in_[head_[oe]]) {
      if (tail_[ie] != start) score[tail_[ie]] = -1;
    }
  }
  int64_t best = -1;
  for (uint32_t r = rank + 1; r < rank_to_node_.size(); ++r) {
    const uint32_t node = rank_to_node_[r];
    score[node] = -1;
    pred[node] = -1;
    for (uint32_t e : in_[node]) {
      const uint32_t t = tail_[e];
      if (score[t] == -1) continue;
      if (score[node] < weight_[e] ||
          (score[node] == weight_[e] && score[pred[node]] <= score[t])) {
        score[node] = weight_[e];
        pred[node] = t;
      }
    }
    if (pred[node] >= 0) score[node] += score[pred[node]];
    if (best < 0 || score[best] < score[node]) best = node;
  }
  return static_cast<uint32_t>(best);
}

// Heaviest bundle: per node the heaviest in-edge (ties: tail with the larger-or-equal running
// score, i.e. the later edge wins), running score = edge weight + score of the chosen tail;
// the best-scoring node is extended to a sink by branch completion, then traced back.
std::string PoaGraph::consensus() const {
  if (rank_to_node_.empty()) return std::string();
  const uint32_t n = num_nodes();
  std::vector<int64_t> score(n, -1), pred(n, -1);
  int64_t best = -1;
  for (uint32_t node : rank_to_node_) {
    for (uint32_t e : in_[node]) {
      const uint32_t t = tail_[e];
      if (score[node] < weight_[e] ||
          (score[node] == weight_[e] && score[pred[node]] <= score[t])) {
        score[node] = weight_[e];
        pred[node] = t;
      }
    }
    if (pred[node] >= 0) score[node] += score[pred[node]];
    if (best < 0 || score[best] < score[node]) best = node;
  }
  if (!out_[best].empty()) {
    std::vector<uint32_t> rank_of(n, 0);
    for (uint32_t r = 0; r < rank_to_node_.size(); ++r) rank_of[rank_to_node_[r]] = r;
    while (!out_[best].empty()) best = branch_completion(rank_of[best], score, pred);
  }
  std::string s;
  for (int64_t v = best; v >= 0; v = pred[v]) s.push_back(static_cast<char>(letter_[v]));
  std::reverse(s.begin(), s.end());
  return s;
}

}  // namespace svs
