/* svscope_b200 — C ABI of the B200-native localGraph hot path.
 *
 * The reference (negi2331026/SVScope) is pure Python and reaches its numeric kernels through
 * Python imports, not an FFI.  Each entry point below names the reference interface it
 * stands in for; the Python modules in svscope_b200/ bind them with ctypes and keep the
 * reference's call signatures (INTEGRATION.md shows the binding a maintainer would add).
 *
 * Conventions: plain pointers and sizes, buffers owned by the caller, every function returns
 * 0 on success or a negative error code; svs_last_error() gives the message of the last
 * failure on that context.  Host pointers unless a parameter says "device".  A context is
 * bound to one CUDA device; calls on one context must not overlap in time.
 */
#ifndef SVSCOPE_B200_H_
#define SVSCOPE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct svs_ctx svs_ctx;
typedef struct svs_reads svs_reads;
typedef struct svs_poa_result svs_poa_result;

enum {
  SVS_OK = 0,
  SVS_ERR_CUDA = -1,        /* CUDA runtime failure (message has the CUDA error string) */
  SVS_ERR_ARG = -2,         /* invalid argument */
  SVS_ERR_UNSUPPORTED = -3, /* e.g. scoring outside the supported convex-gap family */
  SVS_ERR_CAPACITY = -4,    /* a single alignment does not fit the device arena */
  SVS_ERR_INTERNAL = -5
};

/* ---- context -------------------------------------------------------------------------- */
int svs_create(int device, svs_ctx** out);
void svs_destroy(svs_ctx* ctx);
const char* svs_last_error(const svs_ctx* ctx);
const char* svs_version(void);
/* Options (all optional): "poa_threads" CTA size of the window kernel (128|256|384|512; default
 * 384: twelve warps on one alignment, one resident window per SM), "poa_cols" read columns per
 * thread (8), "ring_rows" packed rows kept in shared memory per warp (default 8; limited by the
 * 227 KB of shared memory of an SM), "prune" exact score-bound pruning (default 1), "arena_mb"
 * device scratch arena in MiB (0 = 85 % of free memory), "dp_kernel" (2: the warp-pipelined DP). */
int svs_set_option(svs_ctx* ctx, const char* key, int64_t value);
int64_t svs_get_option(const svs_ctx* ctx, const char* key);

/* Measures the issue rate of the integer pipes (G thread-operations/s) for add, max, xor and
 * fused add+max chains; the denominator of the integer-ALU roofline of the alignment kernel. */
int svs_int_alu_probe(svs_ctx* ctx, double* gops, int n);

/* ---- read sets ------------------------------------------------------------------------ */
/* Copies n_seqs sequences (concatenated, 1 byte per base, seq i = seqs[off[i]..off[i+1]))
 * to device memory once, so that later calls start with the reads resident in HBM. */
int svs_reads_upload(svs_ctx* ctx, const uint8_t* seqs, const int64_t* off, int64_t n_seqs,
                     svs_reads** out);
void svs_reads_free(svs_reads* reads);

/* ---- partial-order alignment ----------------------------------------------------------
 * Replaces `spoa.poa(sequences, algorithm, genmsa, m, n, g, e, q, c)` of pyspoa 0.2.1
 * (reference call sites src/DataScanner.py:206,213 for the window MSA and
 * src/DecisionMaker.py:160,171 for the per-cluster consensus), batched over n_groups
 * independent sequence groups.  Group k consists of the sequences
 * members[group_off[k] .. group_off[k+1]) (indices into the read set), aligned in that order.
 * Only algorithm 1 (global) with convex gaps is on the hot path; other modes return
 * SVS_ERR_UNSUPPORTED.  Empty sequences are skipped and get no MSA row, as in spoa. */
int svs_poa_batch(svs_ctx* ctx, const svs_reads* reads, const int64_t* members,
                  const int64_t* group_off, int64_t n_groups, int algorithm, int m, int n, int g,
                  int e, int q, int c, int want_msa, svs_poa_result** out);
/* The same in two halves: submit enqueues the window kernel on its own stream and returns at
 * once (other calls on the context may run meanwhile and overlap with it on the device);
 * wait blocks until the groups are done (and repeats, in a larger memory tier, the groups
 * whose graph or traceback did not fit their scratch slot). */
int svs_poa_submit(svs_ctx* ctx, const svs_reads* reads, const int64_t* members,
                   const int64_t* group_off, int64_t n_groups, int algorithm, int m, int n, int g,
                   int e, int q, int c, int want_msa, svs_poa_result** out);
int svs_poa_wait(svs_poa_result* res);
/* per group: 0 = done; > 0 = the group could not be aligned (1-5, 8: does not fit the largest
 * memory tier; 3: aligned group of more than 8 letters; 7: |V| + L beyond the score format;
 * 10: a node with more than 31 in-edges).  A failed group has empty outputs; the other groups
 * of the call are not affected. */
int svs_poa_result_status(const svs_poa_result* res, int32_t* status);
/* per group: consensus length, MSA rows (non-empty sequences) and MSA columns */
int svs_poa_result_sizes(const svs_poa_result* res, int64_t* cons_len, int64_t* msa_rows,
                         int64_t* msa_cols);
/* consensus strings concatenated in group order; MSA matrices (rows*cols chars, row-major)
 * concatenated in group order.  Either pointer may be NULL. */
int svs_poa_result_copy(const svs_poa_result* res, uint8_t* consensus, uint8_t* msa);
/* stats[0] nominal DP cells sum (|V|+1)(L+1), [1] alignments, [2] ms of the window kernel
 * (events on its stream, summed over launches), [4] wall ms submit..wait, [5] kernel launches,
 * [7] bytes host->device, [8] bytes device->host (records; the MSA/consensus copy is counted by
 * the caller), [9] algorithmic bytes (read + rank-ordered graph + path),
 * [10] rows exported to global memory, [11] graph rows total, [23] pruning retries,
 * [24..29] SM cycles of thread 0 per phase summed over windows: export, bands + DP,
 * traceback, merge, rank order, MSA/consensus, [32] groups with a non-zero status */
int svs_poa_result_stats(const svs_poa_result* res, double* stats, int n_stats);
void svs_poa_result_free(svs_poa_result* res);

/* Debug/test entry: one alignment of `read` against the graph built from the previous
 * sequences of a group, returning the alignment pairs (node id | -1, position | -1). */
int svs_poa_align_pairs(svs_ctx* ctx, const uint8_t* seqs, const int64_t* off, int64_t n_seqs,
                        int32_t* pair_node, int32_t* pair_pos, int64_t cap, int64_t* n_pairs,
                        int64_t* seq_pair_off);

/* ---- MSA post-processing ---------------------------------------------------------------
 * Replaces SeqEncoder + the column statistics of FindNonSameSite (src/DataScanner.py:124-129,
 * 167-179, 214-219), the ZeroParamNum count of EMCluster (src/ReadsCluster.py:226-234) and
 * pariwiseDistance (src/ReadsCluster.py:44-59), batched over windows.  Window w has an
 * encoded read matrix enc[w] of n_rows[w] x n_cols[w] symbols 0..4 (ref row excluded,
 * row-major, concatenated), a column mask `drop` (1 = flank column to ignore) and a cutoff.
 * Outputs per window: keep[col] = 1 for selected feature columns, nf, zero_params, and the
 * n_rows x n_rows identity COUNTS over the selected columns (the caller divides by nf). */
int svs_msa_features(svs_ctx* ctx, int64_t n_windows, const int8_t* enc, const int64_t* enc_off,
                     const int32_t* n_rows, const int32_t* n_cols, const uint8_t* drop,
                     const int64_t* col_off, const double* cutoff, uint8_t* keep, int32_t* nf,
                     int32_t* zero_params, int32_t* ident, const int64_t* ident_off);

/* ---- sequence mixture model -----------------------------------------------------------
 * Replaces ReadsCluster.EM (src/ReadsCluster.py:190-209: pitheta_updating, gamma_updating and
 * the log-likelihood of the last iteration) for n_tasks (window, K) pairs.  Task t works on
 * X (N[t] x nf[t] symbols 0..4, row-major int8 at x_off[t]; tasks of one window may share it)
 * with K[t] components (1..9).  Start state: hard labels 0..K-1 (init_labels at lab_off[t]),
 * or, when lab_off[t] < 0, theta (K x nf x 5 at theta_off[t]) and pi (K at pi_off[t]).
 * Schedule: [M from labels] E, then n_steps[t] (or n_steps_default when n_steps is NULL)
 * times (M, E); the reference uses 20.  Outputs: gamma (N x K), pi (K), per-read
 * log-likelihood (N), theta (only if want_theta) and status[t]: -1 when finished, else the
 * index it (0 = the initial M-step) of the M-step in which some pi*N < 1 or NaN appeared.
 * The reference then re-draws theta from numpy's global RNG (ReadsCluster.py:185-187); that
 * draw stays on the host: the caller supplies theta, uniform pi, lab_off = -1 and
 * n_steps - it remaining steps (svscope_b200/ReadsCluster.py does this). */
int svs_em_batch(svs_ctx* ctx, int64_t n_tasks, const int8_t* X, const int64_t* x_off,
                 const int32_t* N, const int32_t* nf, const int32_t* K, const int32_t* init_labels,
                 const int64_t* lab_off, int32_t n_steps_default, const int32_t* n_steps,
                 int32_t want_theta, double* gamma, const int64_t* gamma_off, double* theta_io,
                 const int64_t* theta_off, double* pi_io, const int64_t* pi_off, double* loglik,
                 const int64_t* lik_off, int32_t* status);

/* ---- read-by-read edit distances -------------------------------------------------------
 * Replaces the Levenshtein.distance matrix of the commented FindSomClust
 * (src/DecisionMaker.py:76-84): for every group the full symmetric matrix of unit-cost edit
 * distances between its member sequences (Myers/Hyyro bit-parallel, batched).
 * dist holds, per group, n x n int32 row-major at dist_off[k]. */
int svs_edit_distance_matrix(svs_ctx* ctx, const svs_reads* reads, const int64_t* members,
                             const int64_t* group_off, int64_t n_groups, int32_t* dist,
                             const int64_t* dist_off, double* stats, int n_stats);
/* pairs form: distance of reads a[i] and b[i] */
int svs_edit_distance_pairs(svs_ctx* ctx, const svs_reads* reads, const int64_t* a,
                            const int64_t* b, int64_t n_pairs, int32_t* dist, double* stats,
                            int n_stats);

/* ---- MisScore alignments (the step after the Raw.bed) -------------------
 * Replaces the per-pair Biopython call of src/PairwiseCompare.py:19-30
 *   pairwise2.align.globalms(Som, Ger, 1, 0, -1, -1)[0]  ->  MisScore = columns - '|' columns
 * (callers src/PairwiseCompare.py:54-64 CalculateMisscore, :77-88 MisScorePipe, src/SVscope.py:282).
 * Pair k aligns reads a[k] (seqA, rows) and b[k] (seqB, columns); integer scores with
 * open == extend <= 0 (anything else: SVS_ERR_UNSUPPORTED), empty sequences: SVS_ERR_ARG
 * (globalms returns no alignment, the reference's [0] raises).  The alignment is the first
 * one of pairwise2's traversal order (gap in seqA before match/mismatch before gap in seqB).
 * out[4k..4k+3] = score, alignment columns, columns with equal symbols, MisScore.
 * lines (optional): match line of each alignment ('|' equal, '.' different, ' ' gap) at
 * lines[line_off[k]], capacity la+lb bytes each.  stats (optional): DP cells, kernel ms,
 * launches, trace bytes written. */
int svs_misscore_pairs(svs_ctx* ctx, const svs_reads* reads, const int64_t* a, const int64_t* b,
                       int64_t n_pairs, int match, int mismatch, int open, int extend,
                       int32_t* out, uint8_t* lines, const int64_t* line_off, double* stats,
                       int n_stats);

#ifdef __cplusplus
}
#endif
#endif /* SVSCOPE_B200_H_ */
