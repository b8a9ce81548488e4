"""Offline fuzzing of the MisScore kernel logic (CPU emulation, shared headers) against the
literal pairwise2 oracle: random scoring (open == extend), alphabets incl. '-', strip plans.

    python tests/tools/fuzz_misscore.py --seconds 600 --seed 1
"""
import argparse
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle as O      # noqa: E402
from tests.emul import emul         # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=60)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    rng = random.Random(args.seed)
    t0, n = time.time(), 0
    while time.time() - t0 < args.seconds:
        alpha = rng.choice(["ACGT", "AC", "A", "ACGT-", "-A"])
        la = rng.randint(1, 400)
        a = "".join(rng.choice(alpha) for _ in range(la))
        kind = rng.random()
        if kind < 0.5:
            b = []
            e = rng.choice([0.01, 0.1, 0.4])
            for ch in a:
                r = rng.random()
                if r < e / 3:
                    b.append(rng.choice(alpha))
                elif r < 2 * e / 3:
                    b.append(ch)
                    b.append(rng.choice(alpha))
                elif r < e:
                    continue
                else:
                    b.append(ch)
            b = "".join(b) or alpha[0]
        elif kind < 0.7:
            k = rng.randint(0, la)
            b = (a[:k] + a[k + rng.randint(1, 200):]) or alpha[0]
        elif kind < 0.85:
            k = rng.randint(0, la)
            b = a[:k] + "".join(rng.choice(alpha) for _ in range(rng.randint(1, 200))) + a[k:]
        else:
            b = "".join(rng.choice(alpha) for _ in range(rng.randint(1, 400)))
        m = rng.randint(0, 6)
        mm = rng.randint(-9, m)
        g = rng.randint(0, 7)
        threads, cols = rng.choice([(256, 16), (1, 16), (2, 16), (3, 4), (7, 4), (16, 4), (5, 16)])
        c = O.pairwise_first_alignment(a, b, m, mm, -g, -g, want_line=True)
        e = emul.misscore_emul(a, b, m, mm, g, threads, cols, want_line=True)
        if e["status"] != 0 or c["pops"] != 1 or (c["score"], c["length"], c["matches"], c["line"]) != (
                e["score"], e["length"], e["matches"], e["line"]):
            print("MISMATCH", dict(seed=args.seed, n=n, a=a, b=b, par=(m, mm, g), threads=threads, cols=cols), c, e, flush=True)
            sys.exit(1)
        n += 1
    print("ok: %d pairs in %.0f s (seed %d)" % (n, time.time() - t0, args.seed))


if __name__ == "__main__":
    main()
