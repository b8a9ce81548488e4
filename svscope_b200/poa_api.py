"""Batched partial-order alignment on the device (wraps svs_poa_batch)."""
from __future__ import annotations

import ctypes
from typing import List, Sequence

import numpy as np

from ._lib import Context, ReadSet, c_vp, load, ptr

STAT_NAMES = ["cells", "alignments", "dp_ms", "tb_ms", "wall_ms", "dp_launches", "tb_launches",
              "h2d_bytes", "d2h_bytes", "algo_bytes", "exported_rows", "rows", "host_wait_ms", "host_merge_ms",
              "host_plan_ms", "host_pack_ms", "refill_ms", "starved_polls", "launch_ms", "final_ms", "inflight_ms",
              "h2d_ms", "d2h_ms", "prune_retries"]
DEFAULT_SCORES = dict(m=5, n=-4, g=-8, e=-6, q=-10, c=-4)


def poa_groups(ctx: Context, reads: ReadSet, groups: Sequence[Sequence[int]], algorithm: int = 1,
               want_msa: bool = True, scores=None, as_array: bool = False):
    """Align every group (list of read indices, in alignment order) into its own graph.

    Returns (consensus list, msa list (list of row strings per group), stats dict).  With
    ``as_array`` every MSA is a (rows, cols) uint8 array of characters instead of strings
    (views into one buffer: no per-row decoding)."""
    sc = dict(DEFAULT_SCORES)
    if scores:
        sc.update(scores)
    members = np.ascontiguousarray(np.concatenate([np.asarray(g, np.int64) for g in groups])
                                   if len(groups) else np.zeros(0, np.int64))
    goff = np.zeros(len(groups) + 1, np.int64)
    if len(groups):
        goff[1:] = np.cumsum([len(g) for g in groups])
    res = c_vp()
    mem = members if members.size else np.zeros(1, np.int64)
    ctx.check(load().svs_poa_batch(ctx._h, reads._h, ptr(mem), ptr(goff), len(groups), algorithm,
                                   sc["m"], sc["n"], sc["g"], sc["e"], sc["q"], sc["c"],
                                   1 if want_msa else 0, ctypes.byref(res)))
    try:
        ng = len(groups)
        clen = np.zeros(max(ng, 1), np.int64)
        rows = np.zeros(max(ng, 1), np.int64)
        cols = np.zeros(max(ng, 1), np.int64)
        load().svs_poa_result_sizes(res, ptr(clen), ptr(rows), ptr(cols))
        cbuf = np.zeros(max(int(clen[:ng].sum()), 1), np.uint8)
        mbuf = np.zeros(max(int((rows[:ng] * cols[:ng]).sum()), 1), np.uint8)
        load().svs_poa_result_copy(res, ptr(cbuf), ptr(mbuf) if want_msa else None)
        stats = np.zeros(24, np.float64)
        load().svs_poa_result_stats(res, ptr(stats), 24)
    finally:
        load().svs_poa_result_free(res)
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
    return cons, msas, {k: float(stats[i]) for i, k in enumerate(STAT_NAMES)}


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
