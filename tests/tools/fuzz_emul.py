"""Offline fuzzing of the product's host graph + shared cell arithmetic (CPU emulation of the
kernels, tests/emul) against the five-matrix oracle: adversarial sequence groups (homopolymers,
tandem repeats, one-symbol reads, empty reads, duplicates, large indels, unrelated reads), all
pruning modes.  Not part of the test suite; run for as long as you like:

    python tests/tools/fuzz_emul.py --seconds 600 --seed 1
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle as O          # noqa: E402
from tests.emul.emul import EmuSession  # noqa: E402

ALPHA = np.frombuffer(b"ACGT", np.uint8)


def rand_seq(rng, n, k=4):
    return ALPHA[rng.integers(0, k, n)].tobytes().decode()


def mutate(rng, s, sub, ins, dele):
    out = []
    for ch in s:
        r = rng.random()
        if r < sub:
            out.append("ACGT"[rng.integers(4)])
        elif r < sub + ins:
            out.append(ch)
            out.append("ACGT"[rng.integers(4)])
        elif r < sub + ins + dele:
            continue
        else:
            out.append(ch)
    return "".join(out)


def make_group(rng):
    kind = int(rng.integers(0, 8))
    L = int(rng.integers(1, 260))
    n = int(rng.integers(2, 10))
    if kind == 0:      # homopolymer-rich
        base = "".join(ch * int(rng.integers(1, 9)) for ch in rand_seq(rng, max(1, L // 4)))
    elif kind == 1:    # tandem repeat with copy-number changes
        motif = rand_seq(rng, int(rng.integers(1, 7)))
        base = motif * max(1, L // len(motif))
    elif kind == 2:    # two-letter alphabet
        base = rand_seq(rng, L, 2)
    else:
        base = rand_seq(rng, L)
    seqs = []
    for _ in range(n):
        s = base
        if kind == 1 and rng.random() < 0.6:
            motif_len = max(1, len(base) // max(1, L // 3))
            k = int(rng.integers(0, 5)) * motif_len
            s = base[:len(base) // 2] + base[len(base) // 2 + k:] if rng.random() < 0.5 else base[:len(base) // 2] + base[:k] + base[len(base) // 2:]
        for _ in range(int(rng.integers(0, 3))):   # large indels
            p = int(rng.integers(0, len(s) + 1))
            ln = int(rng.integers(1, 60))
            s = s[:p] + s[p + ln:] if rng.random() < 0.5 else s[:p] + rand_seq(rng, ln) + s[p:]
        e = float(rng.choice([0.0, 0.02, 0.1, 0.3]))
        s = mutate(rng, s, e * 0.4, e * 0.3, e * 0.3)
        seqs.append(s or rand_seq(rng, 2))
    r = rng.random()
    if r < 0.15:
        seqs[int(rng.integers(1, n))] = ""
    elif r < 0.3:
        seqs[int(rng.integers(1, n))] = rand_seq(rng, int(rng.integers(1, 120)))   # unrelated
    elif r < 0.4:
        seqs[int(rng.integers(1, n))] = seqs[0]                                    # duplicate
    elif r < 0.45:
        seqs[int(rng.integers(1, n))] = "A" * int(rng.integers(1, 80))
    return seqs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=60)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    rng = np.random.default_rng(args.seed)
    t0 = time.time()
    n = 0
    while time.time() - t0 < args.seconds:
        seqs = make_group(rng)
        mode = int(rng.integers(0, 4))
        kw = [dict(), dict(prune=int(rng.choice([1, 4, 32]))), dict(dyn=float(rng.choice([1.0, 3.5, 4.4, 5.0]))),
              dict(dyn=float(rng.choice([3.5, 4.4, 5.0])), dyn_ext=int(rng.choice([0, 1, 3])))][mode]
        o, e = O.PoaSession(1), EmuSession(ring_rows=int(rng.integers(1, 13)), **kw)
        for k, s in enumerate(seqs):
            a, b = o.add(s), e.add(s)
            if not np.array_equal(a, b):
                print("MISMATCH alignment", dict(seed=args.seed, group=n, read=k, kw=kw), seqs, flush=True)
                sys.exit(1)
        if o.msa() != e.msa() or o.consensus() != e.consensus():
            print("MISMATCH msa/consensus", dict(seed=args.seed, group=n, kw=kw), seqs, flush=True)
            sys.exit(1)
        o.close()
        e.close()
        n += 1
    print("ok: %d groups in %.0f s (seed %d)" % (n, time.time() - t0, args.seed))


if __name__ == "__main__":
    main()
