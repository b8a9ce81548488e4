"""GPU-side fuzzing (needs a B200): random small windows - varying depth, carriers, SV type and
length, error rate, empty reads, a third tag, non-default cutoffs - through the product's batch
path (svscope_b200.batch.localgraph_batch) against oracle.decision, record by record.

    python tests/tools/fuzz_gpu_windows.py --seconds 120 --seed 1 --batch 64
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle as O          # noqa: E402
from svscope_b200 import synth          # noqa: E402


def random_window(rng):
    nt, nn = int(rng.integers(2, 12)), int(rng.integers(2, 12))
    w = synth.make_sv_window(int(rng.integers(1 << 30)), int(rng.integers(120, 420)),
                             "DEL" if rng.random() < 0.5 else "INS", int(rng.integers(10, 110)), nt, nn,
                             int(rng.integers(0, nt + 1)), float(rng.choice([0.0, 0.02, 0.05, 0.12])))
    seqs, ids = list(w[0]), np.array(w[1])
    r = rng.random()
    if r < 0.15:
        for k in rng.choice(np.arange(1, len(seqs)), size=int(rng.integers(1, 3)), replace=False):
            seqs[int(k)] = ""
    elif r < 0.25:
        ids = np.array([x.replace("_normal|", "_other|") if rng.random() < 0.4 else x for x in ids])
    return [seqs, ids, w[2], w[3], w[4]]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=60)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--batch", type=int, default=64)
    args = ap.parse_args()
    from svscope_b200.batch import localgraph_batch
    rng = np.random.default_rng(args.seed)
    t0, n = time.time(), 0
    while time.time() - t0 < args.seconds:
        wins = [random_window(rng) for _ in range(args.batch)]
        kw = {}
        if rng.random() < 0.3:
            kw = dict(readcutoff=int(rng.integers(2, 5)), hcutoff=int(rng.integers(2, 5)), scutoff=float(rng.choice([0.05, 0.2])))
        got = localgraph_batch(wins, **kw).records        # reseeds 2023 per window
        for k, w in enumerate(wins):
            want = O.decision(w[4], w[0], w[1], w[2], w[3], **kw)   # reseed=True: same convention
            if [str(x) for x in want] != [str(x) for x in got[k]]:
                print("MISMATCH", dict(seed=args.seed, batch=n, window=k, kw=kw), flush=True)
                print(" want", want, flush=True)
                print(" got ", got[k], flush=True)
                np.save("gpurun_out/fuzz_gpu_case.npy", np.array(w, dtype=object), allow_pickle=True)
                sys.exit(1)
        n += 1
    print("ok: %d batches of %d windows in %.0f s (seed %d)" % (n, args.batch, time.time() - t0, args.seed))


if __name__ == "__main__":
    main()
