"""MisScore kernel probe on one GPU: consensus-sized pairs (configs[1] scale) through
svs_misscore_pairs, kernel time from the library's CUDA events, end-to-end time with host
strings in and counts out, and the oracle on one core for a few pairs.

    python tests/tools/misscore_probe.py [--pairs 592] > gpurun_out/misscore_probe.json
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def make_pairs(n, seed=1):
    rng = np.random.default_rng(seed)
    pairs = []
    for _ in range(n):
        L = int(rng.integers(5000, 15001))
        ger = rng.integers(0, 4, L).astype(np.uint8)
        sv = int(np.exp(rng.uniform(np.log(50), np.log(2000))))
        k0 = int(rng.integers(100, L - sv - 100))
        som = np.concatenate([ger[:k0], ger[k0 + sv:]]) if rng.random() < 0.5 else \
            np.concatenate([ger[:k0], rng.integers(0, 4, sv).astype(np.uint8), ger[k0:]])
        flip = rng.random(som.size) < 0.005
        som = np.where(flip, (som + 1) % 4, som)
        lut = np.frombuffer(b"ACGT", np.uint8)
        pairs.append((lut[som].tobytes().decode(), lut[ger].tobytes().decode()))
    return pairs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=592)
    ap.add_argument("--cpu-pairs", type=int, default=3)
    args = ap.parse_args()
    from svscope_b200 import _lib
    from svscope_b200.PairwiseCompare import misscore_pairs
    ctx = _lib.Context.default(0)
    pairs = make_pairs(args.pairs)
    misscore_pairs(pairs[:8], ctx=ctx)  # warm-up (arena, module load)
    res = {"pairs": args.pairs, "mean_len": float(np.mean([len(a) + len(b) for a, b in pairs]) / 2)}
    for rep in range(3):
        st = {}
        t0 = time.perf_counter()
        out = misscore_pairs(pairs, ctx=ctx, stats=st)
        wall = time.perf_counter() - t0
        res["run%d" % rep] = {"kernel_ms": st["kernel_ms"], "wall_s": wall, "launches": st["launches"],
                              "gcups_kernel": st["cells"] / st["kernel_ms"] / 1e6,
                              "gcups_e2e": st["cells"] / wall / 1e9,
                              "trace_GBps": st["trace_bytes"] / st["kernel_ms"] / 1e6}
    from oracle import oracle as O
    t0 = time.perf_counter()
    cells = 0
    ok = True
    for k in range(args.cpu_pairs):
        a, b = pairs[k]
        r = O.pairwise_first_alignment(a, b)
        ok = ok and (r["score"], r["length"], r["matches"]) == tuple(int(x) for x in out[k, :3])
        cells += len(a) * len(b)
    res["cpu_gcups_one_core"] = cells / (time.perf_counter() - t0) / 1e9
    res["cpu_equal"] = ok
    print(json.dumps(res))


if __name__ == "__main__":
    main()
