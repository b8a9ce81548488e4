"""Fuzz the numpy restatement (oracle.decision / em_cluster / msa_feature_selection) against the
REFERENCE's own DecisionMaker.Decision imported unmodified (build container only; stubs as in
oracle/gen_golden.py, ``spoa.poa`` = oracle.poa): random small windows with varying depth,
carriers, SV type/length, error rate, empty reads, three tags, tiny clusters (Dirichlet
fallback).  Not part of the test suite.

    python oracle/fuzz_vs_reference.py --seconds 600 --seed 1
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from gen_golden import import_reference  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=60)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    import logging
    logging.disable(logging.CRITICAL)
    RC, DS, DM = import_reference()
    from oracle import oracle as O
    from svscope_b200 import synth
    rng = np.random.default_rng(args.seed)
    t0, n, em = time.time(), 0, 0
    while time.time() - t0 < args.seconds:
        nt, nn = int(rng.integers(2, 12)), int(rng.integers(2, 12))
        w = synth.make_sv_window(int(rng.integers(1 << 30)), int(rng.integers(120, 420)),
                                 "DEL" if rng.random() < 0.5 else "INS", int(rng.integers(10, 110)), nt, nn,
                                 int(rng.integers(0, nt + 1)), float(rng.choice([0.0, 0.02, 0.05, 0.12])))
        seqs, ids = list(w[0]), np.array(w[1])
        r = rng.random()
        if r < 0.15:
            for k in rng.choice(np.arange(1, len(seqs)), size=int(rng.integers(1, 3)), replace=False):
                seqs[int(k)] = ""
        elif r < 0.25:      # a third tag
            ids = np.array([x.replace("_normal|", "_other|") if rng.random() < 0.4 else x for x in ids])
        kw = {}
        if rng.random() < 0.2:
            kw = dict(readcutoff=int(rng.integers(2, 5)), hcutoff=int(rng.integers(2, 5)), scutoff=float(rng.choice([0.05, 0.2])))
        np.random.seed(2023)
        try:
            ref = DM.Decision(w[4], list(seqs), ids.copy(), w[2], w[3], **kw)
            ref_err = None
        except Exception as exc:   # noqa: BLE001
            ref, ref_err = None, type(exc).__name__
        np.random.seed(2023)
        try:
            got = O.decision(w[4], list(seqs), ids.copy(), w[2], w[3], reseed=False, **kw)
            got_err = None
        except Exception as exc:   # noqa: BLE001
            got, got_err = None, type(exc).__name__
        if ref_err != got_err or (ref is not None and [str(x) for x in ref] != [str(x) for x in got]):
            print("MISMATCH", dict(seed=args.seed, n=n, kw=kw, ref_err=ref_err, got_err=got_err), flush=True)
            print(" ref", ref, flush=True)
            print(" got", got, flush=True)
            np.save("/tmp/fuzz_ref_case_%d_%d.npy" % (args.seed, n), np.array([seqs, list(ids), w[2], w[3], w[4]], dtype=object), allow_pickle=True)
            sys.exit(1)
        n += 1
        em += ref is not None and str(ref[-1]).endswith("EMOutput")
    print("ok: %d windows (%d EMOutput) in %.0f s (seed %d)" % (n, em, time.time() - t0, args.seed))


if __name__ == "__main__":
    main()
