"""Fuzz oracle.em_cluster against the REFERENCE's ``ReadsCluster.EMCluster`` (imported
unmodified, build container only) on random feature matrices: planted clusters, pure noise,
tiny N (Dirichlet fallback from the global RNG), constant columns.  K, assignments, BIC list,
theta, gamma, pi must agree (1e-9 relative; NaN positions equal).

    python oracle/fuzz_em_vs_reference.py --seconds 600 --seed 1
"""
import argparse
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from gen_golden import import_reference  # noqa: E402


def close(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return a.shape == b.shape and np.array_equal(np.isnan(a), np.isnan(b)) and \
        np.allclose(np.nan_to_num(a), np.nan_to_num(b), rtol=1e-9, atol=1e-12)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=60)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    warnings.filterwarnings("ignore")
    RC, _, _ = import_reference()
    from oracle import oracle as O
    rng = np.random.default_rng(args.seed)
    t0, n, fb = time.time(), 0, 0
    while time.time() - t0 < args.seconds:
        N, nf = int(rng.integers(3, 26)), int(rng.integers(10, 90))
        kind = rng.random()
        base = rng.integers(0, 4, nf)
        X = np.tile(base, (N, 1))
        if kind < 0.6:      # planted groups
            for g in range(int(rng.integers(1, 4))):
                rows = rng.choice(N, size=int(rng.integers(1, max(2, N // 2))), replace=False)
                cols = rng.choice(nf, size=int(rng.integers(1, nf)), replace=False)
                X[np.ix_(rows, cols)] = rng.integers(0, 5)
        noise = rng.random((N, nf)) < float(rng.choice([0.0, 0.03, 0.15, 0.6]))
        X[noise] = rng.integers(0, 5, int(noise.sum()))
        X = X.astype(np.int64)
        np.random.seed(2023)
        ref = RC.EMCluster(X.copy(), initselection=1)
        np.random.seed(2023)
        got, info = O.em_cluster(X.copy(), reseed=False, return_info=True)
        ok = ref[0] == got[0] and np.array_equal(ref[2], got[2]) and close(ref[6], got[6]) and close(ref[3], got[3]) \
            and close(ref[4], got[4]) and close(ref[5], got[5])
        if not ok:
            print("MISMATCH", dict(seed=args.seed, n=n, N=N, nf=nf, Kref=ref[0], Kgot=got[0]), flush=True)
            np.save("/tmp/fuzz_em_case_%d_%d.npy" % (args.seed, n), X)
            sys.exit(1)
        fb += info["n_fallback"] > 0
        n += 1
    print("ok: %d matrices (%d with Dirichlet fallback) in %.0f s (seed %d)" % (n, fb, time.time() - t0, args.seed))


if __name__ == "__main__":
    main()
