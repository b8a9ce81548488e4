// Host-side partial-order graph of one window (product code).
//
// The GPU computes the sequence-to-graph dynamic programme and its traceback; this class owns
// the graph between two alignments: merge of an alignment path, rank (topological) order,
// export of the rank-ordered arrays the kernels consume, MSA rows and heaviest-bundle
// consensus.  Behavioural contract = spoa's Graph as used by `spoa.poa(sequences, 1)`
// (reference call sites src/DataScanner.py:206,213, src/DecisionMaker.py:160,171; upstream
// behaviour summarised in SURVEY.md Appendix B): node ids in creation order, in-edges in
// order of first traversal, aligned groups on consecutive ranks, DFS visiting order.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace svs {

struct PoaScoring {
  int32_t m = 5, n = -4, g = -8, e = -6, q = -10, c = -4;
};

// Rank-ordered view of the graph handed to the DP / traceback kernels.  Row r (1..R) is
// rank r-1; row 0 is the virtual source row.
struct RankedGraph {
  uint32_t R = 0;
  std::vector<uint8_t> letter;     // [R+1]
  std::vector<uint32_t> pred_off;  // [R+2]
  std::vector<uint32_t> preds;     // row indices, 0 = virtual source
  std::vector<uint8_t> flags;      // [R+1] bit0 sink (no out-edge), bit2 export row to global
  std::vector<int32_t> xslot;      // [R+1] export slot or -1
  std::vector<int32_t> h0;         // [R+1] H[row][0]
  std::vector<uint16_t> col0code;  // [R+1] traceback code of column 0
  std::vector<uint32_t> node_id;   // [R+1]
  std::vector<uint32_t> single_before;  // [R+2] single-predecessor rows among rows 1..i-1
  std::vector<int32_t> depth;      // [R+1][4] dmin, dmax (source->row, inclusive), smin, smax (row->sink, exclusive)
  uint32_t n_export = 0;
  uint32_t max_indeg = 0;
};

class PoaGraph {
 public:
  uint32_t num_nodes() const { return static_cast<uint32_t>(letter_.size()); }
  uint32_t num_sequences() const { return static_cast<uint32_t>(paths_.size()); }
  bool empty() const { return letter_.empty(); }

  // Merge `seq` along alignment pairs (node id | -1, read position | -1), given in forward
  // order.  An empty alignment appends the sequence as a fresh chain.  Empty sequences are
  // ignored (they get no MSA row), as in spoa.
  void add_alignment(const int32_t* pair_node, const int32_t* pair_pos, size_t n_pairs,
                     const uint8_t* seq, uint32_t len);

  // Rank order export; `ring_rows` = number of previous rows the DP kernel keeps on chip:
  // rows with a successor further away than that are flagged for export to global memory.
  void export_ranked(const PoaScoring& sc, uint32_t ring_rows, RankedGraph* out) const;

  std::vector<std::string> msa() const;
  std::string consensus() const;

  const std::vector<uint32_t>& rank_to_node() const { return rank_to_node_; }
  // debug / test accessors
  const std::vector<uint32_t>& in_edges(uint32_t node) const { return in_[node]; }
  uint32_t edge_tail(uint32_t e) const { return tail_[e]; }
  int64_t edge_weight(uint32_t e) const { return weight_[e]; }
  uint8_t node_letter(uint32_t node) const { return letter_[node]; }

 private:
  uint32_t add_node(uint8_t letter);
  void add_edge(uint32_t tail, uint32_t head, int64_t w);
  int64_t add_chain(const uint8_t* seq, uint32_t begin, uint32_t end, std::vector<uint32_t>* path);
  void topological_sort();
  uint32_t branch_completion(uint32_t rank, std::vector<int64_t>& score,
                             std::vector<int64_t>& pred) const;

  std::vector<uint8_t> letter_;
  std::vector<std::vector<uint32_t>> in_, out_;  // edge ids, creation order
  std::vector<std::vector<uint32_t>> aligned_;   // node ids
  std::vector<uint32_t> tail_, head_;
  std::vector<int64_t> weight_;
  std::vector<std::vector<uint32_t>> paths_;     // node path of every added sequence
  std::vector<uint32_t> rank_to_node_;
};

}  // namespace svs
