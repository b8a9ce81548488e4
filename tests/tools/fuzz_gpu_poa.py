"""GPU-side fuzzing (needs a B200): the adversarial sequence groups of fuzz_emul.py through
svs_poa_batch (MSA + consensus, default pruning and kernel configuration) and through
svs_poa_align_pairs (alignment pairs) against the oracle.

    python tests/tools/fuzz_gpu_poa.py --seconds 120 --seed 1 --batch 256
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle as O                      # noqa: E402
from tests.tools.fuzz_emul import make_group        # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=60)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--batch", type=int, default=256)
    args = ap.parse_args()
    from svscope_b200 import _lib
    from svscope_b200.poa_api import align_pairs, poa_groups
    ctx = _lib.Context.default(0)
    rng = np.random.default_rng(args.seed)
    t0, n = time.time(), 0
    while time.time() - t0 < args.seconds:
        groups = [make_group(rng) for _ in range(args.batch)]
        flat = [s for g in groups for s in g]
        reads = _lib.ReadSet(ctx, flat)
        idx, pos = [], 0
        for g in groups:
            idx.append(list(range(pos, pos + len(g))))
            pos += len(g)
        cons, msas, _ = poa_groups(ctx, reads, idx, want_msa=True)
        reads.close()
        for k, g in enumerate(groups):
            oc, om = O.poa(g, 1)
            if oc != cons[k] or om != msas[k]:
                print("MISMATCH msa/consensus", dict(seed=args.seed, batch=n, group=k), g, flush=True)
                sys.exit(1)
        g = groups[int(rng.integers(len(groups)))]
        sess = O.PoaSession(1)
        want = [np.asarray(sess.add(s)) for s in g]
        got = align_pairs(ctx, g)
        for a, b in zip(want, got):
            if not np.array_equal(a.reshape(-1, 2), b.reshape(-1, 2)):
                print("MISMATCH pairs", dict(seed=args.seed, batch=n), g, flush=True)
                sys.exit(1)
        sess.close()
        n += 1
    print("ok: %d batches of %d groups in %.0f s (seed %d)" % (n, args.batch, time.time() - t0, args.seed))


if __name__ == "__main__":
    main()
