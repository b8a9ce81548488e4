"""Offline fuzzing of the product's DP kernel SOURCE on the CPU (tests/emul/cuda_shim.h: poa_dp2.cuh run by one
OS thread per warp and one context per lane) against the five-matrix oracle: the adversarial groups of
fuzz_emul.py scaled up so that alignments span several strips, tandem-repeat expansions, all CTA shapes, ring depths 1-8, pruning on/off.

    python tests/tools/fuzz_dp2_cpu.py --seconds 600 --seed 1
    SVS_EMU_WATCHDOG=120 python tests/tools/fuzz_dp2_cpu.py --seconds 600 --seed 1 --last-group /tmp/last.json

With SVS_EMU_WATCHDOG=<seconds> an emulated CTA that has not finished by then (a cyclic wait between warps) prints the
progress words of its warps and exits with code 4 (tests/emul/dp2_threads.cpp); --last-group keeps the group that was
running.  This is how the hand-over deadlock fixed in poa_dp2.cuh (end-of-batch fence) was found.
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle as O          # noqa: E402
from tests.emul.emul import EmuSession  # noqa: E402
from tests.tools.fuzz_emul import make_group, mutate, rand_seq  # noqa: E402


def make_long_group(rng):
    """A few reads of 300-2600 bases around one template: large indels, a tandem repeat, 0-15 % noise."""
    L = int(rng.integers(300, 2600))
    base = rand_seq(rng, L)
    if rng.random() < 0.4:
        motif = rand_seq(rng, int(rng.integers(1, 12)))
        p = int(rng.integers(0, L - 100))
        rep = motif * (int(rng.integers(20, 200)) // len(motif) + 1)
        base = base[:p] + rep + base[p + len(rep):]
    seqs = []
    for _ in range(int(rng.integers(2, 6))):
        s = base
        for _ in range(int(rng.integers(0, 3))):
            p = int(rng.integers(0, len(s) + 1))
            ln = int(rng.integers(1, 400))
            s = s[:p] + s[p + ln:] if rng.random() < 0.5 else s[:p] + rand_seq(rng, ln) + s[p:]
        e = float(rng.choice([0.0, 0.03, 0.05, 0.15]))
        s = mutate(rng, s, e * 0.4, e * 0.3, e * 0.3)
        seqs.append(s or rand_seq(rng, 2))
    r = rng.random()
    if r < 0.1:
        seqs[int(rng.integers(1, len(seqs)))] = rand_seq(rng, int(rng.integers(50, 1500)))   # unrelated
    elif r < 0.2:
        seqs[int(rng.integers(1, len(seqs)))] = ""
    return seqs


def make_repeat_group(rng):
    """Tandem-repeat expansion: motif of 9-60 bases, reads with different copy numbers (edges that enter a long
    branch late, branches the pruning removes), 3-10 % error - the shape that produced the cyclic wait of seed 11."""
    motif = rand_seq(rng, int(rng.integers(9, 60)))
    f5, f3 = rand_seq(rng, int(rng.integers(100, 700))), rand_seq(rng, int(rng.integers(100, 500)))
    copies = int(rng.integers(2, 12))
    seqs = []
    for _ in range(int(rng.integers(4, 9))):
        e = float(rng.choice([0.03, 0.05, 0.1]))
        s = f5 + motif * (copies + int(rng.choice([0, 0, 1, 2, 3, 5, 8]))) + f3
        seqs.append(mutate(rng, s, e * 0.4, e * 0.3, e * 0.3))
    return seqs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=60)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--threads", type=int, default=0, help="CTA size (0: drawn from 128 / 256 / 384 / 512)")
    ap.add_argument("--ring", type=int, default=0, help="ring depth (0: drawn from 1..8)")
    ap.add_argument("--last-group", default="", help="file that always holds the group being run (json)")
    args = ap.parse_args()
    rng = np.random.default_rng(args.seed)
    t0 = time.time()
    n = n_long = retries = 0
    while time.time() - t0 < args.seconds:
        kind = rng.random()
        long_group = kind < 0.5
        seqs = make_long_group(rng) if kind < 0.4 else make_repeat_group(rng) if kind < 0.5 else make_group(rng)
        threads = int(rng.choice([128, 128, 256, 384, 512]))
        ring = int(rng.integers(1, 9))
        threads, ring = args.threads or threads, args.ring or ring
        kw = dict(ring_rows=ring, warp_threads=threads, warp_prune=int(rng.random() < 0.7))
        if args.last_group:
            import json
            with open(args.last_group, "w") as f:
                json.dump(dict(seed=args.seed, group=n, kw=kw, seqs=seqs), f)
        o, e = O.PoaSession(1), EmuSession(**kw)
        for k, s in enumerate(seqs):
            a, b = o.add(s), e.add(s)
            if not (a.shape == b.shape and np.array_equal(a, b)):
                print("MISMATCH alignment", dict(seed=args.seed, group=n, read=k, kw=kw), seqs, flush=True)
                sys.exit(1)
        if o.consensus() != e.consensus():
            print("MISMATCH consensus", dict(seed=args.seed, group=n, kw=kw), seqs, flush=True)
            sys.exit(1)
        retries += e.warp_retries()
        o.close()
        e.close()
        n += 1
        n_long += long_group
    print("ok: %d groups (%d with reads of 300-2600 bases), %d pruning retries, in %.0f s (seed %d)"
          % (n, n_long, retries, time.time() - t0, args.seed))


if __name__ == "__main__":
    main()
