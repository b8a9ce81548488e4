"""Batched localGraph: every stage of ``Decision`` (src/DecisionMaker.py:110-191) run over a
whole list of windows at once, with the numeric work on the device.

    stage 1  gate on tags / read counts                      DecisionMaker.py:126-134  (host)
    stage 2  window MSA by partial-order alignment           DataScanner.py:206,213    (svs_poa_batch)
    stage 3  encode, flank columns, feature columns, read    DataScanner.py:214-219,   (svs_msa_features;
             identity matrix, ZeroParamNum                    ReadsCluster.py:226-243   CallMargin on host)
    stage 4  Ward tree + K-cluster cuts                      ReadsCluster.py:243,94    (host scipy, as the
                                                                                        reference: bit-identical init)
    stage 5  EM for K = 1..min(9, N-1), BIC, K selection     ReadsCluster.py:190-277   (svs_em_batch)
    stage 6  per-cluster consensus POA                       DecisionMaker.py:156-176  (svs_poa_batch)
    stage 7  10-field records                                DecisionMaker.py:178-190  (host)
    (opt.)   read-by-read edit distances per window          DecisionMaker.py:76-84    (svs_edit_distance_matrix)

Windows are rows ``[sequenceList, ReadIDs, flank_5, flank_3, TDRecord]`` (the npz row format).
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

from ._lib import Context, ReadSet, load, ptr
from .poa_api import poa_groups

ENC_LUT = np.full(256, 255, np.uint8)
for _ch, _v in (("A", 0), ("T", 1), ("C", 2), ("G", 3), ("-", 4)):
    ENC_LUT[ord(_ch)] = _v
    ENC_LUT[ord(_ch.lower())] = _v
DEC_LUT = np.frombuffer(b"ATCG-", np.uint8)
SEED = 2023            # ReadsCluster.py:42
N_STEPS = 20           # ReadsCluster.py:190
MAX_C = 9              # ReadsCluster.py:221


# byte / launch accounting of the auxiliary kernels (filled by the wrappers below)
ACCT: Dict[str, float] = {}


def _acct(key: str, v: float) -> None:
    ACCT[key] = ACCT.get(key, 0.0) + float(v)


def read_tag(read_id: str) -> str:
    return read_id.split("|")[0].split("_")[-1]


def encode_msa(msa) -> np.ndarray:
    """SeqEncoder over all rows (DataScanner.py:124-129, 214); KeyError on foreign symbols.
    ``msa`` is a list of row strings or a (rows, cols) uint8 array of characters."""
    if len(msa) == 0:
        return np.zeros((0, 0), np.int64)
    if isinstance(msa, np.ndarray):
        raw = msa
    else:
        raw = np.frombuffer("".join(msa).encode(), np.uint8).reshape(len(msa), -1)
    enc = ENC_LUT[raw]
    if (enc == 255).any():
        raise KeyError(chr(int(raw[enc == 255][0])))
    return enc


def margin_columns(ref_row, flank_5: str, flank_3: str) -> np.ndarray:
    """CallMargin (DataScanner.py:146-165) without the per-character Python loop.

    Forward walk: the collected non-gap prefix can equal flank_5 only when it has len(flank_5)
    characters, so either the first len(flank_5) non-gap columns are returned (prefix matches)
    or the walk never stops and returns all of them.  The backward walk covers columns
    len-1 .. 1 only (column 0 is never visited, DataScanner.py:159)."""
    row = ref_row if isinstance(ref_row, np.ndarray) else np.frombuffer(ref_row.encode(), np.uint8)
    nongap = np.flatnonzero(row != ord("-"))
    n5, n3 = len(flank_5), len(flank_3)
    if n5 == 0:
        fwd = nongap[:0] if (row.size == 0 or row[0] == ord("-")) else nongap
    elif nongap.size >= n5 and row[nongap[:n5]].tobytes().decode() == flank_5:
        fwd = nongap[:n5]
    else:
        fwd = nongap
    back_all = nongap[nongap >= 1][::-1]
    if n3 == 0:
        bwd = back_all[:0] if (row.size < 2 or row[-1] == ord("-")) else back_all
    elif back_all.size >= n3 and row[back_all[:n3][::-1]].tobytes().decode() == flank_3:
        bwd = back_all[:n3]
    else:
        bwd = back_all
    return np.concatenate([fwd, bwd]).astype(np.int64)


# ----------------------------------------------------------------------------------------------
# device wrappers
# ----------------------------------------------------------------------------------------------
def msa_features(ctx: Context, encs: List[np.ndarray], drops: List[np.ndarray], cutoffs: List[float]):
    """For every window: (keep mask, nf, zero_params, identity counts) via svs_msa_features.
    ``encs[w]`` is the read part of the encoded MSA (rows x cols, symbols 0..4)."""
    nw = len(encs)
    if nw == 0:
        return []
    rows = np.array([e.shape[0] for e in encs], np.int32)
    cols = np.array([e.shape[1] for e in encs], np.int32)
    enc_off = np.zeros(nw + 1, np.int64)
    enc_off[1:] = np.cumsum(rows.astype(np.int64) * cols)
    col_off = np.zeros(nw + 1, np.int64)
    col_off[1:] = np.cumsum(cols.astype(np.int64))
    id_off = np.zeros(nw + 1, np.int64)
    id_off[1:] = np.cumsum(rows.astype(np.int64) ** 2)
    enc_cat = np.zeros(max(int(enc_off[-1]), 1), np.int8)
    drop_cat = np.zeros(max(int(col_off[-1]), 1), np.uint8)
    for w in range(nw):
        e = encs[w]
        enc_cat[enc_off[w]:enc_off[w + 1]] = (e.view(np.int8) if e.dtype == np.uint8 else e.astype(np.int8)).reshape(-1)
        drop_cat[col_off[w]:col_off[w + 1]] = drops[w]
    keep = np.zeros_like(drop_cat)
    nf = np.zeros(nw, np.int32)
    zp = np.zeros(nw, np.int32)
    ident = np.zeros(max(int(id_off[-1]), 1), np.int32)
    cut = np.ascontiguousarray(cutoffs, np.float64)
    ctx.check(load().svs_msa_features(ctx._h, nw, ptr(enc_cat), ptr(enc_off), ptr(rows), ptr(cols), ptr(drop_cat),
                                      ptr(col_off), ptr(cut), ptr(keep), ptr(nf), ptr(zp), ptr(ident), ptr(id_off)))
    _acct("feat_h2d_bytes", enc_cat.nbytes + drop_cat.nbytes)
    _acct("feat_d2h_bytes", keep.nbytes + ident.nbytes + nf.nbytes + zp.nbytes)
    _acct("aux_launches", 2)
    out = []
    for w in range(nw):
        n = int(rows[w])
        out.append((keep[col_off[w]:col_off[w + 1]].astype(bool), int(nf[w]), int(zp[w]),
                    ident[id_off[w]:id_off[w + 1]].reshape(n, n)))
    return out


@dataclass
class EmTaskSpec:
    x_index: int                 # index into the list of X matrices
    K: int
    labels: Optional[np.ndarray] = None     # hard labels 0..K-1, or None -> theta/pi start
    theta: Optional[np.ndarray] = None
    pi: Optional[np.ndarray] = None
    n_steps: int = N_STEPS


def em_batch(ctx: Context, Xs: List[np.ndarray], tasks: List[EmTaskSpec], want_theta: bool = False):
    """Run svs_em_batch.  Returns per task dict(gamma, pi, lik, theta|None, status)."""
    nt = len(tasks)
    if nt == 0:
        return []
    x_off_by = np.zeros(len(Xs) + 1, np.int64)
    x_off_by[1:] = np.cumsum([x.size for x in Xs])
    x_cat = np.zeros(max(int(x_off_by[-1]), 1), np.int8)
    for i, x in enumerate(Xs):
        x_cat[x_off_by[i]:x_off_by[i + 1]] = x.astype(np.int8).ravel()
    N = np.array([Xs[t.x_index].shape[0] for t in tasks], np.int32)
    nf = np.array([Xs[t.x_index].shape[1] for t in tasks], np.int32)
    K = np.array([t.K for t in tasks], np.int32)
    x_off = np.array([x_off_by[t.x_index] for t in tasks], np.int64)
    lab_off = np.full(nt, -1, np.int64)
    labs, pos = [], 0
    for i, t in enumerate(tasks):
        if t.labels is not None:
            lab_off[i] = pos
            labs.append(np.asarray(t.labels, np.int32))
            pos += int(N[i])
    lab_cat = np.concatenate(labs) if labs else np.zeros(1, np.int32)
    g_off = np.zeros(nt + 1, np.int64)
    g_off[1:] = np.cumsum(N.astype(np.int64) * K)
    p_off = np.zeros(nt + 1, np.int64)
    p_off[1:] = np.cumsum(K.astype(np.int64))
    l_off = np.zeros(nt + 1, np.int64)
    l_off[1:] = np.cumsum(N.astype(np.int64))
    any_theta_in = any(t.labels is None for t in tasks)
    need_theta = want_theta or any_theta_in
    t_off = np.zeros(nt + 1, np.int64)
    t_off[1:] = np.cumsum(K.astype(np.int64) * nf * 5)
    theta = np.zeros(max(int(t_off[-1]), 1) if need_theta else 1, np.float64)
    pi = np.zeros(max(int(p_off[-1]), 1), np.float64)
    for i, t in enumerate(tasks):
        if t.labels is None:
            theta[t_off[i]:t_off[i + 1]] = np.asarray(t.theta, np.float64).ravel()
            pi[p_off[i]:p_off[i + 1]] = np.asarray(t.pi, np.float64)
    gamma = np.zeros(max(int(g_off[-1]), 1), np.float64)
    lik = np.zeros(max(int(l_off[-1]), 1), np.float64)
    status = np.zeros(nt, np.int32)
    steps = np.array([t.n_steps for t in tasks], np.int32)
    ctx.check(load().svs_em_batch(ctx._h, nt, ptr(x_cat), ptr(x_off), ptr(N), ptr(nf), ptr(K), ptr(lab_cat),
                                  ptr(lab_off), N_STEPS, ptr(steps), 1 if want_theta else 0, ptr(gamma),
                                  ptr(g_off), ptr(theta) if need_theta else None, ptr(t_off), ptr(pi), ptr(p_off),
                                  ptr(lik), ptr(l_off), ptr(status)))
    _acct("em_h2d_bytes", x_cat.nbytes + lab_cat.nbytes + pi.nbytes + (theta.nbytes if any_theta_in else 0))
    _acct("em_d2h_bytes", gamma.nbytes + pi.nbytes + lik.nbytes + status.nbytes + (theta.nbytes if want_theta else 0))
    _acct("aux_launches", len({(int(k), int(n) > 256) for k, n in zip(K, N)}))
    out = []
    for i in range(nt):
        n, k, f = int(N[i]), int(K[i]), int(nf[i])
        out.append(dict(gamma=gamma[g_off[i]:g_off[i + 1]].reshape(n, k).copy(),
                        pi=pi[p_off[i]:p_off[i + 1]].copy(),
                        lik=lik[l_off[i]:l_off[i + 1]].copy(),
                        theta=theta[t_off[i]:t_off[i + 1]].reshape(k, f, 5).copy() if want_theta else None,
                        status=int(status[i])))
    return out


def edit_distance_matrices(ctx: Context, reads: ReadSet, groups: Sequence[Sequence[int]]):
    """Full symmetric Levenshtein matrix of every group (svs_edit_distance_matrix)."""
    ng = len(groups)
    if ng == 0:
        return [], dict(cells=0.0, ms=0.0, bytes=0.0, pairs=0.0)
    members = np.ascontiguousarray(np.concatenate([np.asarray(g, np.int64) for g in groups]))
    goff = np.zeros(ng + 1, np.int64)
    goff[1:] = np.cumsum([len(g) for g in groups])
    doff = np.zeros(ng + 1, np.int64)
    doff[1:] = np.cumsum([len(g) ** 2 for g in groups])
    dist = np.zeros(max(int(doff[-1]), 1), np.int32)
    stats = np.zeros(4, np.float64)
    mem = members if members.size else np.zeros(1, np.int64)
    ctx.check(load().svs_edit_distance_matrix(ctx._h, reads._h, ptr(mem), ptr(goff), ng, ptr(dist), ptr(doff),
                                              ptr(stats), 4))
    _acct("ed_d2h_bytes", 4.0 * float(stats[3]))
    _acct("aux_launches", 1)
    mats = [dist[doff[g]:doff[g + 1]].reshape(len(groups[g]), len(groups[g])).copy() for g in range(ng)]
    return mats, dict(cells=float(stats[0]), ms=float(stats[1]), bytes=float(stats[2]), pairs=float(stats[3]))


# ----------------------------------------------------------------------------------------------
# mixture model driver (EMCluster, ReadsCluster.py:221-277) over many windows
# ----------------------------------------------------------------------------------------------
def _bic(lik: np.ndarray, K: int, nf: int, zero_params: int = 0) -> float:
    n_theta = K - 1 + K * nf * 4 - zero_params            # ReadsCluster.py:215
    return 2 * lik.sum() - n_theta * np.log(len(lik))     # :218


EM_MAX_READS = 1024   # rows of one window the mixture kernel takes (csrc/em.cu)


def _resume_fit(ctx, X, K, labels, first, want_theta):
    """Finish one (window, K) fit whose M-step hit the re-draw condition (pi*N < 1 or NaN,
    ReadsCluster.py:179-187).  theta is drawn on the host from numpy's global RNG exactly as
    the reference does, then the device continues from (uniform pi, drawn theta)."""
    nf = X.shape[1]
    res, left, draws = first, N_STEPS, 0
    while res["status"] >= 0:
        left -= res["status"]
        theta = np.stack([np.random.dirichlet(np.ones(5), size=nf) for _ in range(K)])
        draws += 1
        spec = EmTaskSpec(0, K, None, theta, np.repeat(1 / K, K), left)
        res = em_batch(ctx, [X], [spec], want_theta=want_theta)[0]
        if left == 0 and res["status"] < 0 and want_theta:
            res["theta"] = theta
    if want_theta and res.get("theta") is None:
        res["theta"] = theta
    return res, draws


def em_cluster_many(ctx: Context, Xs: List[np.ndarray], sims: List[np.ndarray], zero_params: List[int],
                    want_theta: bool = False, reseed: bool = True, max_C: int = 0):
    """EMCluster for every X in ``Xs`` (features already selected, similarity matrices given);
    ``max_C`` = largest K tried (reference default 9 = the kernel's largest K; 0: that default).

    Returns per window dict(K, labels, gamma, pi, theta, bics, n_redraws)."""
    from scipy.cluster.hierarchy import fcluster, linkage
    tasks, owner = [], []
    trees = []
    for w, (X, sim) in enumerate(zip(Xs, sims)):
        Z = linkage(sim, "ward")                                           # ReadsCluster.py:243
        trees.append(Z)
        for K in range(1, int(np.min([(max_C or MAX_C) + 1, X.shape[0]]))):  # :238, :246
            labels = fcluster(Z, K, criterion="maxclust") - 1             # :94-97
            tasks.append(EmTaskSpec(w, K, labels.astype(np.int32)))
            owner.append(w)
    results = em_batch(ctx, Xs, tasks, want_theta=want_theta)
    per_window: Dict[int, list] = {}
    for t, r, w in zip(tasks, results, owner):
        per_window.setdefault(w, []).append((t, r))
    out = []
    for w, X in enumerate(Xs):
        fits = per_window.get(w, [])
        n_redraws = 0
        if any(r["status"] >= 0 or np.isnan(r["lik"]).any() for _, r in fits):
            # windows that consume random numbers are replayed in the reference's order:
            # K ascending, at most 5 attempts per K while the BIC is NaN (:248-252)
            if reseed:
                np.random.seed(SEED)
            redo = []
            for t, r in fits:
                tries = 5
                while True:
                    if r["status"] >= 0:
                        r, d = _resume_fit(ctx, X, t.K, t.labels, r, want_theta)
                        n_redraws += d
                    tries -= 1
                    if not np.isnan(_bic(r["lik"], t.K, X.shape[1])) or tries == 0:
                        break
                    r = em_batch(ctx, [X], [EmTaskSpec(0, t.K, t.labels)], want_theta=want_theta)[0]
                redo.append((t, r))
            fits = redo
        N, nf = X.shape
        bics = np.array([_bic(r["lik"], t.K, nf, zero_params[w]) for t, r in fits])
        best = int(np.nanargmax(bics))
        K = best + 1
        if K == 1 and bics[0] - bics[1] <= nf * np.log(N):                # :268-272
            K, best = 2, 1
        r = fits[best][1]
        out.append(dict(K=K, labels=np.argmax(r["gamma"], axis=1), gamma=r["gamma"], pi=r["pi"],
                        theta=r["theta"], bics=bics, n_redraws=n_redraws))
    return out


# ----------------------------------------------------------------------------------------------
# the batched Decision
# ----------------------------------------------------------------------------------------------
@dataclass
class BatchOutput:
    records: List[list]
    timings: Dict[str, float] = field(default_factory=dict)
    stats: Dict[str, float] = field(default_factory=dict)
    edit_distances: Optional[List[Optional[np.ndarray]]] = None
    aux: Optional[List[dict]] = None


def upload_windows(ctx: Context, windows) -> ReadSet:
    """All sequences of all windows (reference row first) plus one trailing empty sequence."""
    seqs = []
    for w in windows:
        seqs.extend(w[0])
    seqs.append("")
    return ReadSet(ctx, seqs)


def localgraph_batch(windows, ctx: Optional[Context] = None, reads: Optional[ReadSet] = None,
                     windowFlags: Optional[Sequence[str]] = None, Tlabel="tumor", readcutoff=3, hcutoff=3,
                     scutoff=0.05, edit_distance: bool = False, keep_aux: bool = False,
                     reseed: bool = True, chunks: int = 3) -> BatchOutput:
    """``Decision`` for every window of the list; returns the 10-field records in input order.

    ``reads`` may be a ReadSet made by ``upload_windows`` beforehand (inputs already resident
    in HBM); otherwise the sequences are uploaded here.

    The windows are dealt (by cost) into ``chunks`` interleaved sub-batches whose device stages
    are submitted asynchronously: the window-MSA kernel of sub-batch k+1 and the consensus
    kernel of sub-batch k-1 run (filling each other's tails) while the host selects features
    and drives the mixture model of sub-batch k; the edit-distance matrices run on a side
    thread behind the first MSA kernels."""
    import threading
    from .poa_api import PoaJob
    ctx = ctx or Context.default()
    tm: Dict[str, float] = {k: 0.0 for k in ("gate", "poa_msa", "features", "mixture", "poa_consensus", "records")}
    ACCT.clear()
    t0 = time.perf_counter()
    nw = len(windows)
    flags = list(windowFlags) if windowFlags is not None else ["NormalOutput"] * nw
    if reads is None:
        reads = upload_windows(ctx, windows)
    base = np.zeros(nw + 1, np.int64)
    base[1:] = np.cumsum([len(w[0]) for w in windows])
    empty_idx = int(base[-1])
    records: List[Optional[list]] = [None] * nw
    aux: List[dict] = [dict() for _ in range(nw)]
    # ---- stage 1: gate ------------------------------------------------------------------
    live = []
    for i, w in enumerate(windows):
        seqs, ids, f5, f3, rec = w
        chrom, start, end = rec.strip().split("\t")[0:3]
        records[i] = [chrom, start, end, "-", "-", 0, "-", "-", 0, flags[i]]
        tags, counts = np.unique(np.array([read_tag(x) for x in ids]), return_counts=True)
        if len(seqs) > 3 and tags.shape[0] >= 2 and np.min(counts) >= 3:
            live.append(i)
    tm["gate"] = time.perf_counter() - t0
    # the per-cluster consensus runs on the upper-cased sequences (DecisionMaker.py:157-171 decodes
    # the encoded MSA rows, and the encoder upper-cases); the window MSA is case-sensitive
    cons_reads, own_cons_reads = reads, False
    if any(s != s.upper() for i in live for s in windows[i][0]):
        cons_reads = ReadSet(ctx, [s.upper() for w in windows for s in w[0]] + [""])
        own_cons_reads = True
    # ---- sub-batches: largest windows first, dealt round-robin ----------------------------
    cost = {i: sum(len(s) for s in windows[i][0]) * max(len(s) for s in windows[i][0]) for i in live}
    order = sorted(live, key=lambda i: -cost[i])
    nchunk = max(1, min(int(chunks), len(order) // 64 if len(order) >= 128 else 1))
    parts = [order[k::nchunk] for k in range(nchunk)]
    st_poa: Dict[str, float] = {}

    def add_stats(st):
        for k, v in st.items():
            if k != "status":
                st_poa[k] = st_poa.get(k, 0.0) + float(v)

    # ---- stage 2: window MSA, all sub-batches submitted at once -------------------------------
    t1 = time.perf_counter()
    msa_jobs = [PoaJob(ctx, reads, [list(range(int(base[i]), int(base[i + 1]))) for i in part], want_msa=True)
                for part in parts]
    # ---- optional: read-by-read edit distances, behind the MSA kernels on a side thread -----------
    dists = None
    st_ed = dict(cells=0.0, ms=0.0, bytes=0.0, pairs=0.0)
    ed_thread, ed_box = None, {}
    if edit_distance:
        def _ed():
            t_ed = time.perf_counter()
            try:
                ed_groups = [list(range(int(base[i]) + 1, int(base[i + 1]))) for i in live]
                ed_box["mats"], ed_box["st"] = edit_distance_matrices(ctx, reads, ed_groups)
            except BaseException as exc:   # re-raised on the main thread
                ed_box["err"] = exc
            ed_box["t"] = time.perf_counter() - t_ed
        ed_thread = threading.Thread(target=_ed, daemon=True)
        ed_thread.start()
    tm["poa_msa"] += time.perf_counter() - t1
    n_failed = 0
    n_em = 0
    n_redraw = 0
    pending_cons = []
    for part, job in zip(parts, msa_jobs):
        t1 = time.perf_counter()
        _, msas, st_msa = job.result(as_array=True, strict=False)
        add_stats(st_msa)
        tm["poa_msa"] += time.perf_counter() - t1
        # ---- stage 3: encode, margins, features ---------------------------------------------------
        t1 = time.perf_counter()
        if st_msa["status"].any():
            # a window the device cannot align (graph beyond the largest memory tier, more than 31
            # in-edges at one node, ...) fails alone: flagged record, the batch goes on
            for k in np.flatnonzero(st_msa["status"]):
                records[part[k]][9] = flags[part[k]] + "|GraphLimit%d" % int(st_msa["status"][k])
                n_failed += 1
            ok = [k for k in range(len(part)) if st_msa["status"][k] == 0]
            part = [part[k] for k in ok]
            msas = [msas[k] for k in ok]
        encs, drops, cutoffs, id_lists, row_src = [], [], [], [], []
        for i, msa in zip(part, msas):
            seqs, ids, f5, f3, _ = windows[i]
            ids = np.asarray(ids)
            lens = np.array([len(s) for s in seqs[1:]])
            enc = encode_msa(msa)
            nonempty = np.flatnonzero(lens != 0)
            if nonempty.size != lens.size:
                # DataScanner.py:198-209: gap rows are appended once per NON-empty read and the
                # id list becomes the non-empty ids twice (the reference's own quirk)
                kept = list(ids[nonempty])
                enc = np.concatenate([enc, np.full((len(kept), enc.shape[1]), 4, enc.dtype)], axis=0)
                ids = np.array(kept + kept)
                src = [int(base[i]) + 1 + int(r) for r in nonempty] + [empty_idx] * len(kept)
            else:
                src = [int(base[i]) + 1 + r for r in range(len(lens))]
            drop = np.zeros(enc.shape[1], np.uint8)
            drop[margin_columns(msa[0], f5, f3)] = 1
            encs.append(enc)
            drops.append(drop)
            cutoffs.append(float(max([hcutoff, enc.shape[0] * scutoff])))
            id_lists.append(ids)
            row_src.append(src)
        feats = msa_features(ctx, [e[1:] for e in encs], drops, cutoffs)
        tm["features"] += time.perf_counter() - t1
        # ---- stage 4/5: mixture model ----------------------------------------------------------
        t1 = time.perf_counter()
        em_idx, Xs, sims, zps = [], [], [], []
        for k, (enc, (keep, nf, zp, ident)) in enumerate(zip(encs, feats)):
            n = enc.shape[0] - 1
            if n > EM_MAX_READS and nf >= 10:
                # the mixture kernel holds one window's responsibilities in shared memory (svs_em_batch: 1..1024
                # rows): such a window fails alone with a flagged record, the batch goes on
                records[part[k]][9] = flags[part[k]] + "|MixtureLimit%d" % n
                n_failed += 1
                continue
            if n != 0 and nf >= 10:                                            # DecisionMaker.py:137
                X = np.ascontiguousarray(enc[1:][:, keep])
                sim = ident.astype(np.float64) / nf
                np.fill_diagonal(sim, 1.0)
                em_idx.append(k)
                Xs.append(X)
                sims.append(sim)
                zps.append(zp)
        fits = em_cluster_many(ctx, Xs, sims, zps, want_theta=False, reseed=reseed)
        n_em += len(em_idx)
        n_redraw += sum(1 for f in fits if f["n_redraws"] > 0)
        tm["mixture"] += time.perf_counter() - t1
        # ---- stage 6: cluster consensus (submitted; collected after the next sub-batch) ------------
        t1 = time.perf_counter()
        cons_groups, cons_owner = [], []
        plan = {}
        for k, fit in zip(em_idx, fits):
            ids = id_lists[k]
            labels = fit["labels"]
            som, germ = [], []
            for lab in np.unique(labels):
                members = np.where(labels == lab)[0]
                kinds = np.unique([read_tag(x) for x in ids[members]])
                if kinds.shape[0] == 1 and kinds[0] == Tlabel and members.shape[0] >= readcutoff:
                    som.append(members)
                elif members.shape[0] >= readcutoff:
                    germ.append(members)
            plan[k] = (som, germ)
            if len(som) > 0:
                for kind, lst in (("som", som), ("germ", germ)):
                    for c, members in enumerate(lst):
                        src = [row_src[k][m] for m in members]
                        nonzero = any(reads.off[s + 1] - reads.off[s] > 0 for s in src)
                        cons_groups.append(src if nonzero else [])
                        cons_owner.append((k, kind, c, nonzero))
        cjob = PoaJob(ctx, cons_reads, cons_groups, want_msa=False)
        tm["poa_consensus"] += time.perf_counter() - t1
        pending_cons.append((part, cjob, cons_owner, plan, em_idx, fits, id_lists, Xs))
    # ---- stage 7: records ----------------------------------------------------------------------
    for part, cjob, cons_owner, plan, em_idx, fits, id_lists, Xs in pending_cons:
        t1 = time.perf_counter()
        cons, _, st_cons = cjob.result(strict=False)
        add_stats(st_cons)
        bad_cons = {cons_owner[k][0] for k in np.flatnonzero(st_cons["status"])}
        n_failed += len(bad_cons)
        tm["poa_consensus"] += time.perf_counter() - t1
        t1 = time.perf_counter()
        seq_out: Dict[tuple, str] = {}
        for (k, kind, c, nonzero), sq in zip(cons_owner, cons):
            seq_out[(k, kind, c)] = sq if nonzero else "-"
        for n_k, (k, fit) in enumerate(zip(em_idx, fits)):
            i = part[k]
            som, germ = plan[k]
            ids = id_lists[k]
            if keep_aux:
                aux[i] = dict(K=fit["K"], labels=fit["labels"], gamma=fit["gamma"], pi=fit["pi"], bics=fit["bics"],
                              n_redraws=fit["n_redraws"], nf=Xs[n_k].shape[1])
            if k in bad_cons:
                records[i][9] = flags[i] + "|GraphLimit"
            elif len(som) > 0 and len(germ) > 0:                                  # DecisionMaker.py:178
                rec = records[i]
                records[i] = [rec[0], rec[1], rec[2],
                              ";".join(seq_out[(k, "som", c)] for c in range(len(som))),
                              ";".join(",".join(list(ids[m])) for m in som),
                              len(som),
                              ";".join(seq_out[(k, "germ", c)] for c in range(len(germ))),
                              ";".join(",".join(list(ids[m])) for m in germ),
                              len(germ),
                              flags[i] + "|EMOutput"]
        tm["records"] += time.perf_counter() - t1
    if ed_thread is not None:
        t1 = time.perf_counter()
        ed_thread.join()
        if "err" in ed_box:
            raise ed_box["err"]
        st_ed = ed_box["st"]
        dists = [None] * nw
        for i, mt in zip(live, ed_box["mats"]):
            dists[i] = mt
        tm["edit_distance"] = ed_box["t"]
        tm["edit_distance_wait"] = time.perf_counter() - t1
    if own_cons_reads:
        cons_reads.close()
    tm["total"] = time.perf_counter() - t0
    stats = {"poa_" + k: v for k, v in st_poa.items()}
    stats["poa_failed_windows"] = n_failed
    stats.update({"ed_" + k: v for k, v in st_ed.items()})
    stats.update(ACCT)
    stats["windows"] = nw
    stats["windows_live"] = len(live)
    stats["windows_em"] = n_em
    stats["em_redraw_windows"] = n_redraw
    stats["sub_batches"] = nchunk
    return BatchOutput(records=records, timings=tm, stats=stats, edit_distances=dists,
                       aux=aux if keep_aux else None)
