"""ORACLE — TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's localGraph hot path.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module, and only as the checker or as the timed CPU baseline.  The
product package ``svscope_b200`` never imports it.

Pinning status
--------------
* POA (``poa``): **parity unpinned** — restates pyspoa 0.2.1 / rvaser-spoa from recollection
  (oracle/spoa_oracle.cpp); the wheel is absent and the reference has no vectors for it.
* Levenshtein (``levenshtein``): **parity unpinned** — textbook DP; module absent, no live
  reference call site (src/DecisionMaker.py:34,76-84 are comments).
* MisScore (``pairwise_first_alignment``, ``aligment_score``, ``calculate_misscore``): **parity
  unpinned** — restates Bio.pairwise2 (absent; src/PairwiseCompare.py:8-11,19-64) from
  recollection: oracle/pairwise2_oracle.py (literal, strings) and oracle/misscore_oracle.c.
* Feature selection, mixture model, Decision (``msa_feature_selection``, ``em_cluster``,
  ``decision``): **pinned** against the reference's own Python (`src/DataScanner.py`,
  `src/ReadsCluster.py`, `src/DecisionMaker.py`) imported unmodified in the build container by
  ``oracle/gen_golden.py``; the outputs are committed under ``tests/golden/`` and re-checked
  by ``tests/test_oracle_golden.py`` wherever the suite runs; ``oracle/fuzz_vs_reference.py``
  additionally compared 32 549 random windows record by record with the reference's ``Decision``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import List, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile oracle/_build/liboracle.so with the committed Makefile."""
    srcs = [os.path.join(_HERE, f) for f in ("spoa_oracle.cpp", "lev_oracle.c", "misscore_oracle.c", "Makefile")]
    stale = not os.path.exists(_LIB_PATH) or any(os.path.getmtime(f) > os.path.getmtime(_LIB_PATH) for f in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        L.spo_new.restype = ctypes.c_void_p
        L.spo_new.argtypes = [ctypes.c_int] * 7
        L.spo_free.argtypes = [ctypes.c_void_p]
        L.spo_add.restype = ctypes.c_int64
        L.spo_add.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int64]
        L.spo_align_only.restype = ctypes.c_int64
        L.spo_align_only.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int64]
        L.spo_last_alignment.restype = ctypes.c_int64
        L.spo_last_alignment.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
        L.spo_last_cells.restype = ctypes.c_int64
        L.spo_last_cells.argtypes = [ctypes.c_void_p]
        L.spo_last_score.restype = ctypes.c_int32
        L.spo_last_score.argtypes = [ctypes.c_void_p]
        L.spo_error.restype = ctypes.c_char_p
        L.spo_error.argtypes = [ctypes.c_void_p]
        for fn in ("spo_num_nodes", "spo_num_edges", "spo_num_sequences"):
            getattr(L, fn).restype = ctypes.c_int64
            getattr(L, fn).argtypes = [ctypes.c_void_p]
        L.spo_graph_dump.argtypes = [ctypes.c_void_p] * 8
        L.spo_consensus.restype = ctypes.c_int64
        L.spo_consensus.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
        L.spo_msa_dims.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.spo_msa.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.spo_set_blocked.argtypes = [ctypes.c_void_p, ctypes.c_int64]
        for fn in ("spo_blocked_kept_rows", "spo_blocked_recomputed"):
            getattr(L, fn).restype = ctypes.c_int64
            getattr(L, fn).argtypes = [ctypes.c_void_p]
        L.lev_dp.restype = ctypes.c_int64
        L.lev_dp.argtypes = [ctypes.c_char_p, ctypes.c_int64, ctypes.c_char_p, ctypes.c_int64]
        L.lev_myers.restype = ctypes.c_int64
        L.lev_myers.argtypes = [ctypes.c_char_p, ctypes.c_int64, ctypes.c_char_p, ctypes.c_int64]
        L.lev_matrix.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p]
        L.pw2_first.restype = ctypes.c_int64
        L.pw2_first.argtypes = [ctypes.c_char_p, ctypes.c_int64, ctypes.c_char_p, ctypes.c_int64, ctypes.c_int,
                                ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        _lib = L
    return _lib


# --------------------------------------------------------------------------------------
# POA  (restates spoa.poa; reference call sites DataScanner.py:206,213, DecisionMaker.py:160,171)
# --------------------------------------------------------------------------------------
class PoaSession:
    """Step-wise access to the restated spoa graph (alignment pairs, rank order, edges)."""

    def __init__(self, algorithm: int = 1, m=5, n=-4, g=-8, e=-6, q=-10, c=-4, block_rows: int = 0):
        """``block_rows``: 0 = flat five-matrix engine; > 0 = row-checkpoint engine with that many
        rows per block; -1 = row-checkpoint engine with ~2 GB blocks (bounded memory: windows whose
        full matrices do not fit, e.g. BASELINE configs[2] at full size).  Same recurrences, same
        traceback code (spoa_oracle.cpp: trace_back)."""
        self._h = lib().spo_new(algorithm, m, n, g, e, q, c)
        if not self._h:
            raise ValueError("oracle supports the convex gap mode only (g<e, g>q, e<c)")
        self.cells = 0
        if block_rows:
            lib().spo_set_blocked(self._h, int(block_rows))

    @property
    def blocked_stats(self) -> dict:
        return dict(kept_rows=int(lib().spo_blocked_kept_rows(self._h)),
                    recomputed_blocks=int(lib().spo_blocked_recomputed(self._h)))

    def close(self):
        if self._h:
            lib().spo_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _pairs(self, n: int) -> np.ndarray:
        nodes = np.empty(n, np.int32)
        pos = np.empty(n, np.int32)
        lib().spo_last_alignment(self._h, nodes.ctypes.data, pos.ctypes.data, n)
        return np.stack([nodes, pos], axis=1)

    def add(self, seq: str) -> np.ndarray:
        b = seq.encode()
        n = lib().spo_add(self._h, b, len(b))
        if n < 0:
            raise ValueError(lib().spo_error(self._h).decode())
        self.cells += lib().spo_last_cells(self._h)
        return self._pairs(n)

    def align_only(self, seq: str) -> np.ndarray:
        b = seq.encode()
        n = lib().spo_align_only(self._h, b, len(b))
        if n < 0:
            raise ValueError(lib().spo_error(self._h).decode())
        return self._pairs(n)

    @property
    def score(self) -> int:
        return int(lib().spo_last_score(self._h))

    def graph(self) -> dict:
        L = lib()
        nn, ne = L.spo_num_nodes(self._h), L.spo_num_edges(self._h)
        rank_node = np.empty(nn, np.int32)
        letter = np.empty(nn, np.uint8)
        indeg = np.empty(nn, np.int32)
        nal = np.empty(nn, np.int32)
        outdeg = np.empty(nn, np.int32)
        tail = np.empty(ne, np.int32)
        weight = np.empty(ne, np.int64)
        L.spo_graph_dump(self._h, rank_node.ctypes.data, letter.ctypes.data, indeg.ctypes.data,
                         tail.ctypes.data, weight.ctypes.data, nal.ctypes.data, outdeg.ctypes.data)
        return dict(rank_node=rank_node, letter=letter, indeg=indeg, in_tail=tail,
                    in_weight=weight, n_aligned=nal, outdeg=outdeg)

    def consensus(self) -> str:
        cap = lib().spo_num_nodes(self._h) + 1
        buf = ctypes.create_string_buffer(cap)
        n = lib().spo_consensus(self._h, buf, cap)
        return buf.raw[:n].decode()

    def msa(self) -> List[str]:
        r, c = ctypes.c_int64(), ctypes.c_int64()
        lib().spo_msa_dims(self._h, ctypes.byref(r), ctypes.byref(c))
        if r.value == 0 or c.value == 0:
            return ["" for _ in range(r.value)]
        buf = ctypes.create_string_buffer(r.value * c.value)
        lib().spo_msa(self._h, buf)
        raw = buf.raw
        return [raw[i * c.value:(i + 1) * c.value].decode() for i in range(r.value)]


def poa(sequences: Sequence[str], algorithm: int = 0, genmsa: bool = True, m=5, n=-4, g=-8,
        e=-6, q=-10, c=-4, min_coverage=None, block_rows: int = 0) -> Tuple[str, List[str]]:
    """Signature of ``spoa.poa`` (pyspoa 0.2.1).  ``min_coverage`` is not restated."""
    if min_coverage is not None:
        raise NotImplementedError("min_coverage is not used by the reference")
    s = PoaSession(algorithm, m, n, g, e, q, c, block_rows=block_rows)
    try:
        for seq in sequences:
            s.add(seq)
        cons = s.consensus()
        msa = s.msa() if genmsa else []
        poa.last_cells = s.cells
        return cons, msa
    finally:
        s.close()


poa.last_cells = 0


# --------------------------------------------------------------------------------------
# Levenshtein (contract of Levenshtein.distance; commented design DecisionMaker.py:76-84)
# --------------------------------------------------------------------------------------
def levenshtein(a: str, b: str, bitparallel: bool = False) -> int:
    ab, bb = a.encode(), b.encode()
    fn = lib().lev_myers if bitparallel else lib().lev_dp
    return int(fn(ab, len(ab), bb, len(bb)))


def levenshtein_matrix(seqs: Sequence[str], bitparallel: bool = True) -> np.ndarray:
    """dist[i,j] = dist[j,i] = Levenshtein(seq_i, seq_j) (DecisionMaker.py:78-84, commented)."""
    n = len(seqs)
    cat = "".join(seqs).encode()
    off = np.zeros(n + 1, np.int64)
    off[1:] = np.cumsum([len(s) for s in seqs])
    out = np.zeros((n, n), np.int64)
    buf = np.frombuffer(cat, np.uint8) if cat else np.zeros(1, np.uint8)
    lib().lev_matrix(buf.ctypes.data, off.ctypes.data, n, 1 if bitparallel else 0, out.ctypes.data)
    return out


# --------------------------------------------------------------------------------------
# MSA encoding and feature selection (DataScanner.py:124-220)
# --------------------------------------------------------------------------------------
_ALPHA = {"A": 0, "T": 1, "C": 2, "G": 3, "-": 4}
_LUT = np.full(256, 255, np.uint8)
for _k, _v in _ALPHA.items():
    _LUT[ord(_k)] = _v
    _LUT[ord(_k.lower())] = _v
_DEC = np.frombuffer(b"ATCG", np.uint8)


def encode_rows(msa: Sequence[str]) -> np.ndarray:
    """DataScanner.SeqEncoder (:124-129) applied to every MSA row; KeyError on other symbols."""
    if len(msa) == 0:
        return np.zeros((0, 0), np.int64)
    arr = np.frombuffer("".join(msa).encode(), np.uint8).reshape(len(msa), -1)
    enc = _LUT[arr]
    if (enc == 255).any():
        bad = chr(int(arr[enc == 255][0]))
        raise KeyError(bad)
    return enc.astype(np.int64)


def decode_row(row: np.ndarray) -> str:
    """DataScanner.SeqDecoder (:131-137): drop code 4, map 0..3 -> A,T,C,G."""
    row = np.asarray(row)
    keep = row[row != 4].astype(np.int64)
    return _DEC[keep].tobytes().decode()


def call_margin(ref_row: str, flank_5: str, flank_3: str) -> np.ndarray:
    """DataScanner.CallMargin (:146-165).  Forward: collect non-gap columns until the collected
    string equals flank_5 (checked after every column, including gap columns); backward over
    columns len-1 .. 1 likewise for flank_3.  Never-equal => runs to the end."""
    idx: List[int] = []
    acc = ""
    for col, ch in enumerate(ref_row):
        if ch != "-":
            acc += ch
            idx.append(col)
        if acc == flank_5:
            break
    acc = ""
    for col in range(len(ref_row) - 1, 0, -1):
        ch = ref_row[col]
        if ch != "-":
            acc = ch + acc
            idx.append(col)
        if acc == flank_3:
            break
    return np.array(idx)


def column_symbol_counts(mat: np.ndarray) -> np.ndarray:
    """5 x ncols counts of symbols 0..4 (the table FindNonSameSite / EMCluster both build)."""
    return np.stack([(mat == a).sum(axis=0) for a in range(5)]).astype(np.float64) if mat.size else \
        np.zeros((5, mat.shape[1] if mat.ndim == 2 else 0))


def find_non_same_site(mat: np.ndarray, cutoff: float = 3) -> np.ndarray:
    """DataScanner.FindNonSameSite (:167-179): columns whose 2nd-largest symbol count >= cutoff."""
    cnt = column_symbol_counts(mat)
    second = np.sort(cnt, axis=0)[-2]
    return np.where(second >= cutoff)[0]


def msa_feature_selection(sequence_list, flank_5, flank_3, read_ids, hcutoff=3, scutoff=0.05,
                          poa_fn=None):
    """DataScanner.MSAFeatureSelection (:181-220), including the empty-read branch quirk
    (:198-209: ``DELReads`` is built from the NON-empty ids and all sequences are aligned)."""
    poa_fn = poa_fn or poa
    read_ids = np.asarray(read_ids)
    lens = np.array([len(s) for s in sequence_list[1:]])
    empty = np.where(lens == 0)[0]
    if empty.shape[0] > 0:
        keep = np.setdiff1d(np.arange(len(read_ids)), empty)
        kept_ids = list(read_ids[keep])
        _, msa = poa_fn(sequence_list, 1)
        enc_rows = [r for r in encode_rows(msa)]
        width = len(enc_rows[-1])
        read_ids = np.array(kept_ids + kept_ids)
        enc = np.array(enc_rows + [[4] * width] * len(kept_ids))
        msa = list(msa) + [["-"] * width] * len(kept_ids)
    else:
        _, msa = poa_fn(sequence_list, 1)
        enc = encode_rows(msa)
    margin = call_margin(msa[0], flank_5, flank_3)
    inner = np.setdiff1d(np.arange(enc.shape[1]), margin)
    reads_inner = enc[1:, inner]
    cutoff = max([hcutoff, enc.shape[0] * scutoff])
    X = reads_inner[:, find_non_same_site(reads_inner, cutoff=cutoff)]
    return enc, X, read_ids


# --------------------------------------------------------------------------------------
# Sequence mixture model (ReadsCluster.py)
# --------------------------------------------------------------------------------------
EPS = 1e-10
SEED = 2023


def pairwise_identity(X: np.ndarray) -> np.ndarray:
    """ReadsCluster.pariwiseDistance/CallDistance (:44-59): fraction of equal columns, diag 1."""
    N, nf = X.shape
    S = np.eye(N)
    denom = nf if nf != 0 else 1
    for i in range(N):
        for j in range(i):
            S[i, j] = S[j, i] = int((X[i] == X[j]).sum()) / denom
    return S


def _clip(p):
    return np.clip(p, EPS, 1 - EPS)  # ReadsCluster.CheckParam :70-74


def m_step(K: int, gamma: np.ndarray, X: np.ndarray):
    """ReadsCluster.pitheta_updating (:162-188).  Returns (pi, theta, fell_back)."""
    N, nf = X.shape
    pi = gamma.sum(axis=0) / N
    bad = (pi * N < 1).any() or np.isnan(pi).any()
    if not bad:
        tot = np.dot(gamma.T, np.ones((N, nf), dtype=np.int64))
        theta = np.dstack([np.dot(gamma.T, np.where(X == a, 1, 0)) / tot for a in range(5)])
        return pi, theta, False
    pi = np.repeat(1 / K, K)
    theta = np.stack([np.random.dirichlet(np.ones(5), size=nf) for _ in range(K)])  # global RNG
    return pi, theta, True


def e_step(K: int, pi: np.ndarray, theta: np.ndarray, X: np.ndarray):
    """ReadsCluster.gamma_updating (:132-155).  Returns (gamma, logjoint)."""
    N, nf = X.shape
    logt = np.log(_clip(theta))
    L = np.zeros((N, K))
    for a in range(5):
        L += np.dot(np.where(X == a, 1, 0), logt[:, :, a].T)
    L += np.log(pi.reshape((K, 1)).T)
    cols = []
    for i in range(K):
        d = np.clip(L - L[:, i].reshape((N, 1)), -700, 700)
        cols.append(1 / np.exp(d).sum(axis=1))
    return np.vstack(cols).T, L


def per_read_loglik(pi, theta, gamma, X):
    """ReadsCluster.loglik (:104-122): sum_k gamma[n,k] * (sum_f log theta_c + log clip(pi_k))."""
    N, nf = X.shape
    K = pi.shape[0]
    logt = np.log(_clip(theta))
    onehot = np.eye(5)[X]
    out = np.zeros((N,))
    for k in range(K):
        out += ((logt[k] * onehot).sum(axis=2).sum(axis=1) + np.log(_clip(pi[k]))) * gamma[:, k]
    return out


def em_fit(K: int, X: np.ndarray, Z, nstep: int = 20):
    """ReadsCluster.EM (:190-209) with initselection=1 (par_init :76-101).  The per-iteration
    log-likelihood of the reference is dead work except for the last one (:216)."""
    from scipy.cluster.hierarchy import fcluster
    N = X.shape[0]
    labels = fcluster(Z, K, criterion="maxclust")
    g0 = np.zeros((N, K))
    g0[np.arange(N), labels - 1] = 1
    n_fallback = 0
    pi, theta, fb = m_step(K, g0, X)
    n_fallback += fb
    gamma, _ = e_step(K, pi, theta, X)
    for _ in range(nstep):
        pi, theta, fb = m_step(K, gamma, X)
        n_fallback += fb
        gamma, _ = e_step(K, pi, theta, X)
    lik = per_read_loglik(pi, theta, gamma, X)
    return dict(pi=pi, theta=theta, gamma=gamma, lik=lik, n_fallback=int(n_fallback), init_labels=labels)


def bic_score(fit: dict, zero_params: int = 0) -> float:
    """ReadsCluster.BIC (:211-219)."""
    K, nf, A = fit["theta"].shape
    n_theta = len(fit["pi"]) - 1 + K * nf * (A - 1) - zero_params
    N = len(fit["lik"])
    return 2 * fit["lik"].sum() - n_theta * np.log(N)


def zero_param_num(X: np.ndarray) -> int:
    """ReadsCluster.EMCluster :226-234: number of (symbol, column) pairs with zero count."""
    return int((column_symbol_counts(X) == 0).sum())


def em_cluster(X: np.ndarray, max_C: int = 9, reseed: bool = True, return_info: bool = False):
    """ReadsCluster.EMCluster (:221-277) with initselection=1.

    ``reseed=True`` applies the per-window convention of SURVEY.md §8c: the global numpy RNG
    is reset to seed 2023 (ReadsCluster.py:42) at entry, i.e. the state a fresh single-window
    run of the reference sees.  Returns the reference's 7-list."""
    from scipy.cluster.hierarchy import linkage
    if reseed:
        np.random.seed(SEED)
    zp = zero_param_num(X)
    N, nf = X.shape
    sim = pairwise_identity(X)
    Z = linkage(sim, "ward")
    bics, fits = [], []
    total_fb = 0
    for K in range(1, int(np.min([max_C + 1, N]))):
        val, tries, fit = np.nan, 5, None
        while np.isnan(val) and tries != 0:
            fit = em_fit(K, X, Z)
            total_fb += fit["n_fallback"]
            val = bic_score(fit)
            tries -= 1
        bics.append(bic_score(fit, zp))
        fits.append(fit)
    best = int(np.nanargmax(np.array(bics)))
    K = best + 1
    if K == 1 and bics[0] - bics[1] <= nf * np.log(N):
        K, best = 2, 1
    f = fits[best]
    out = [K, X, np.argmax(f["gamma"], axis=1), f["theta"], f["gamma"], f["pi"], np.array(bics)]
    if return_info:
        return out, dict(n_fallback=total_fb, fits=fits, sim=sim, Z=Z, zero_params=zp)
    return out


# --------------------------------------------------------------------------------------
# Decision (DecisionMaker.py:110-191) and the npz driver (SomTDDetector.py:63-73)
# --------------------------------------------------------------------------------------
def _tag(read_id: str) -> str:
    return read_id.split("|")[0].split("_")[-1]


def decision(TDRecord, sequenceList, ReadIDs, flank_5, flank_3, windowFlag="NormalOutput",
             Tlabel="tumor", readcutoff=3, hcutoff=3, scutoff=0.05, poa_fn=None, reseed=True):
    poa_fn = poa_fn or poa
    chrom, start, end = TDRecord.strip().split("\t")[0:3]
    tags, counts = np.unique(np.array([_tag(x) for x in ReadIDs]), return_counts=True)
    record = [chrom, start, end, "-", "-", 0, "-", "-", 0, windowFlag]
    if not (len(sequenceList) > 3 and tags.shape[0] >= 2 and np.min(counts) >= 3):
        return record
    enc, X, ReadIDs = msa_feature_selection(sequenceList, flank_5, flank_3, ReadIDs,
                                            hcutoff=hcutoff, scutoff=scutoff, poa_fn=poa_fn)
    if X.shape[0] == 0 or X.shape[1] < 10:
        return record
    K, _, labels, _, _, _, _ = em_cluster(X, reseed=reseed)
    ids = np.array(ReadIDs)
    som_idx, germ_idx = [], []
    for lab in np.unique(labels):
        members = np.where(labels == lab)[0]
        kinds = np.unique([_tag(x) for x in ids[members]])
        if kinds.shape[0] == 1 and kinds[0] == Tlabel and members.shape[0] >= readcutoff:
            som_idx.append(members)
        elif members.shape[0] >= readcutoff:
            germ_idx.append(members)

    def cluster_consensus(members):
        seqs = [decode_row(r) for r in enc[members + 1]]
        if max(len(s) for s in seqs) > 0:
            return poa_fn(seqs, 1)[0]
        return "-"

    som_seq = [cluster_consensus(mm) for mm in som_idx]
    germ_seq = [cluster_consensus(mm) for mm in germ_idx]
    if len(som_seq) > 0 and len(germ_idx) > 0:
        record = [chrom, start, end,
                  ";".join(som_seq),
                  ";".join(",".join(list(ids[mm])) for mm in som_idx),
                  len(som_seq),
                  ";".join(germ_seq),
                  ";".join(",".join(list(ids[mm])) for mm in germ_idx),
                  len(germ_seq),
                  windowFlag + "|EMOutput"]
    return record


def tdscope_npz(TDRecord, sequenceList, ReadIDs, flank_5, flank_3):
    """SomTDDetector.TDscope_npz (:63-73)."""
    return decision(TDRecord, sequenceList, ReadIDs, flank_5, flank_3)


def format_record(rec) -> str:
    """Raw.bed line as written at SVscope.py:175."""
    return "\t".join(str(x) for x in rec) + "\n"


# --------------------------------------------------------------------------------------
# MisScore  (next row F1: src/PairwiseCompare.py:19-64; Bio.pairwise2 restated, parity unpinned)
def pairwise_first_alignment(a: str, b: str, match: int = 1, mismatch: int = 0, open: int = -1,
                             extend: int = -1, want_line: bool = False):
    """First alignment of ``pairwise2.align.globalms(a, b, match, mismatch, open, extend)``:
    dict(score, length, matches, pops[, line]).  IndexError on empty input, like ``[...][0]``."""
    ab, bb = a.encode(), b.encode()
    out = np.zeros(4, np.int64)
    line = ctypes.create_string_buffer(len(ab) + len(bb) + 1) if want_line else None
    rc = lib().pw2_first(ab, len(ab), bb, len(bb), match, mismatch, open, extend, out.ctypes.data, line)
    if rc == -2:
        raise IndexError("list index out of range")
    if rc != 0:
        raise MemoryError("pw2_first")
    res = dict(score=int(out[0]), length=int(out[1]), matches=int(out[2]), pops=int(out[3]))
    if want_line:
        res["line"] = line.raw[:res["length"]].decode()
    return res


def aligment_score(SomConsensus: str, GerConsensus: str, cutoff: int = 0) -> int:
    """PairwiseCompare.py:19-30 (spelling of the reference kept)."""
    if cutoff == 0:
        r = pairwise_first_alignment(SomConsensus, GerConsensus)
        return r["length"] - r["matches"]
    alig = pairwise_first_alignment(SomConsensus, GerConsensus, want_line=True)["line"]
    td = alig[cutoff:len(alig) - cutoff]
    return len(td) - td.count("|")


def smaller_absolute_value(a, b):
    """PairwiseCompare.py:32-36."""
    return a if abs(a) < abs(b) else b


def calculate_misscore(somSeqList: str, germSeqList: str):
    """PairwiseCompare.py:54-64 on the two ';'-joined consensus columns of a Raw.bed row."""
    mis = 1000000000000000000000
    for som in somSeqList.split(";"):
        for ger in germSeqList.split(";"):
            score = aligment_score(som, ger)
            if len(som) < len(ger):
                score = (-1) * score
            mis = smaller_absolute_value(mis, score)
    return mis
