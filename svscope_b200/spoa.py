"""Drop-in for ``from spoa import poa`` (pyspoa 0.2.1) on the localGraph path.

Reference call sites: src/DataScanner.py:206,213 (window MSA) and src/DecisionMaker.py:160,171
(per-cluster consensus), always ``poa(sequences, 1)``.  The dynamic programme and traceback
run on the GPU (svs_poa_batch); there is no CPU fallback."""
from __future__ import annotations

from typing import List, Sequence, Tuple

from ._lib import Context, ReadSet
from .poa_api import poa_groups


def poa(sequences: Sequence[str], algorithm: int = 0, genmsa: bool = True, m: int = 5, n: int = -4,
        g: int = -8, e: int = -6, q: int = -10, c: int = -4, min_coverage=None) -> Tuple[str, List[str]]:
    """Same signature and return value as ``spoa.poa``: (consensus, msa rows).

    Only ``algorithm=1`` (global alignment) with convex gap scores is implemented — the mode
    the reference uses on this path; anything else raises."""
    if min_coverage is not None:
        raise NotImplementedError("min_coverage is not used on the localGraph path")
    ctx = Context.default()
    reads = ReadSet(ctx, list(sequences))
    try:
        cons, msas, stats = poa_groups(ctx, reads, [list(range(len(sequences)))], algorithm=algorithm,
                                       want_msa=genmsa, scores=dict(m=m, n=n, g=g, e=e, q=q, c=c))
    finally:
        reads.close()
    poa.last_stats = stats
    return cons[0], (msas[0] if genmsa else [])


poa.last_stats = {}
