"""Batched partial-order alignment on the device (wraps svs_poa_batch)."""
from __future__ import annotations

import ctypes
from typing import List, Sequence

import numpy as np

from ._lib import Context, ReadSet, SvsError, c_vp, load, ptr

STAT_NAMES = ["cells", "alignments", "dp_ms", "tb_ms", "wall_ms", "dp_launches", "tb_launches",
              "h2d_bytes", "d2h_bytes", "algo_bytes", "exported_rows", "rows", "eval_cells", "r13", "r14", "r15", "r16", "r17",
              "r18", "r19", "r20", "r21", "r22", "prune_retries", "cyc_export", "cyc_dp", "cyc_traceback", "cyc_merge",
              "cyc_rank", "cyc_finish", "r30", "r31", "failed_groups", "wcyc_loop", "wcyc_wait_left", "wcyc_wait_right", "wcyc_wait_end"]
N_STATS = len(STAT_NAMES)
STATUS_TEXT = {1: "graph nodes exceed the largest memory tier", 2: "graph edges exceed the largest memory tier",
               3: "aligned group of more than 8 distinct letters", 4: "rank-order stack exceeds the memory tier",
               5: "traceback codes exceed the largest memory tier", 6: "traceback failed",
               7: "|V| + L beyond the packed score format", 8: "output arena exhausted", 9: "not run",
               10: "graph node with more than 31 in-edges"}
DEFAULT_SCORES = dict(m=5, n=-4, g=-8, e=-6, q=-10, c=-4)


class PoaJob:
    """A submitted batch of groups (svs_poa_submit): the window kernel runs on its own stream
    while the caller issues other work; ``result()`` waits and fetches the outputs."""

    def __init__(self, ctx: Context, reads: ReadSet, groups, algorithm=1, want_msa=True, scores=None):
        sc = dict(DEFAULT_SCORES)
        if scores:
            sc.update(scores)
        self.ctx, self.reads, self.ng, self.want_msa = ctx, reads, len(groups), want_msa
        members = np.ascontiguousarray(np.concatenate([np.asarray(g, np.int64) for g in groups])
                                       if len(groups) else np.zeros(0, np.int64))
        goff = np.zeros(len(groups) + 1, np.int64)
        if len(groups):
            goff[1:] = np.cumsum([len(g) for g in groups])
        self._res = c_vp()
        mem = members if members.size else np.zeros(1, np.int64)
        ctx.check(load().svs_poa_submit(ctx._h, reads._h, ptr(mem), ptr(goff), len(groups), algorithm,
                                        sc["m"], sc["n"], sc["g"], sc["e"], sc["q"], sc["c"],
                                        1 if want_msa else 0, ctypes.byref(self._res)))

    def result(self, as_array=False, strict=True):
        try:
            self.ctx.check(load().svs_poa_wait(self._res))
            return _fetch(self._res, self.ng, self.want_msa, as_array, strict)
        finally:
            self.close()

    def close(self):
        if self._res is not None and self._res.value:
            load().svs_poa_result_free(self._res)
        self._res = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def poa_groups(ctx: Context, reads: ReadSet, groups: Sequence[Sequence[int]], algorithm: int = 1,
               want_msa: bool = True, scores=None, as_array: bool = False, strict: bool = True):
    """Align every group (list of read indices, in alignment order) into its own graph.

    Returns (consensus list, msa list (list of row strings per group), stats dict).  With
    ``as_array`` every MSA is a (rows, cols) uint8 array of characters instead of strings
    (views into one buffer: no per-row decoding).  A group the device could not align
    (``stats["status"][k] != 0``) raises unless ``strict`` is false, in which case its outputs
    are empty and the other groups are unaffected."""
    return PoaJob(ctx, reads, groups, algorithm, want_msa, scores).result(as_array=as_array, strict=strict)


def _fetch(res, ng, want_msa, as_array, strict):
    if True:
        clen = np.zeros(max(ng, 1), np.int64)
        rows = np.zeros(max(ng, 1), np.int64)
        cols = np.zeros(max(ng, 1), np.int64)
        load().svs_poa_result_sizes(res, ptr(clen), ptr(rows), ptr(cols))
        cbuf = np.zeros(max(int(clen[:ng].sum()), 1), np.uint8)
        mbuf = np.zeros(max(int((rows[:ng] * cols[:ng]).sum()), 1), np.uint8)
        load().svs_poa_result_copy(res, ptr(cbuf), ptr(mbuf) if want_msa else None)
        stats = np.zeros(N_STATS, np.float64)
        load().svs_poa_result_stats(res, ptr(stats), N_STATS)
        status = np.zeros(max(ng, 1), np.int32)
        load().svs_poa_result_status(res, ptr(status))
        status = status[:ng]
    if strict and status.any():
        k = int(np.flatnonzero(status)[0])
        raise SvsError("svscope_b200: group %d could not be aligned: %s" % (k, STATUS_TEXT.get(int(status[k]), status[k])))
    cons, msas = [], []
    co = mo = 0
    craw, mraw = cbuf.tobytes(), mbuf.tobytes()
    for k in range(ng):
        cons.append(craw[co:co + int(clen[k])].decode())
        co += int(clen[k])
        r, c = int(rows[k]), int(cols[k])
        if want_msa and as_array:
            msas.append(mbuf[mo:mo + r * c].reshape(r, c))
            mo += r * c
        elif want_msa:
            msas.append([mraw[mo + i * c: mo + (i + 1) * c].decode() for i in range(r)])
            mo += r * c
        else:
            msas.append([])
    out = {k: float(stats[i]) for i, k in enumerate(STAT_NAMES)}
    out["status"] = status
    out["copy_bytes"] = float(cbuf.nbytes + (mbuf.nbytes if want_msa else 0))
    return cons, msas, out


def align_pairs(ctx: Context, seqs: Sequence[str]) -> List[np.ndarray]:
    """Debug/test: the alignment pairs of every sequence of one group (svs_poa_align_pairs)."""
    enc = [s.encode() for s in seqs]
    off = np.zeros(len(enc) + 1, np.int64)
    off[1:] = np.cumsum([len(b) for b in enc])
    joined = b"".join(enc)
    buf = np.frombuffer(joined, np.uint8) if joined else np.zeros(1, np.uint8)
    cap = int(4 * off[-1] + 64) * max(1, len(seqs))
    nodes = np.zeros(cap, np.int32)
    pos = np.zeros(cap, np.int32)
    n = ctypes.c_int64()
    poff = np.zeros(len(seqs) + 1, np.int64)
    ctx.check(load().svs_poa_align_pairs(ctx._h, ptr(buf), ptr(off), len(seqs), ptr(nodes), ptr(pos), cap,
                                         ctypes.byref(n), ptr(poff)))
    if n.value > cap:
        raise RuntimeError("pair buffer too small")
    return [np.stack([nodes[poff[k]:poff[k + 1]], pos[poff[k]:poff[k + 1]]], axis=1) for k in range(len(seqs))]
