"""Offline fuzzing of batch.localgraph_batch's HOST logic with oracle/numpy stand-ins for the
device wrappers (tests/test_host_logic.py::_device_stand_ins) against oracle.decision.

    python tests/tools/fuzz_batch_host.py --seconds 600 --seed 1
"""
import argparse
import os
import sys
import time
import warnings

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle as O                 # noqa: E402
from svscope_b200 import batch, synth          # noqa: E402
from tests.test_host_logic import _device_stand_ins  # noqa: E402


class _Patch:
    def setattr(self, obj, name, value):
        setattr(obj, name, value)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=60)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    warnings.filterwarnings("ignore")
    O.build()
    _device_stand_ins(O, _Patch())
    rng = np.random.default_rng(args.seed)
    t0, n, n_em = time.time(), 0, 0
    while time.time() - t0 < args.seconds:
        wins = []
        for _ in range(int(rng.integers(1, 9))):
            nt, nn = int(rng.integers(1, 12)), int(rng.integers(1, 12))
            w = synth.make_sv_window(int(rng.integers(1 << 30)), int(rng.integers(120, 380)),
                                     "DEL" if rng.random() < 0.5 else "INS", int(rng.integers(10, 110)), nt, nn,
                                     int(rng.integers(0, nt + 1)), float(rng.choice([0.0, 0.02, 0.05, 0.12])))
            seqs, ids = list(w[0]), np.array(w[1])
            r = rng.random()
            if r < 0.2:
                for k in rng.choice(np.arange(1, len(seqs)), size=int(rng.integers(1, min(4, len(seqs)))), replace=False):
                    seqs[int(k)] = ""
            elif r < 0.3:
                ids = np.array([x.replace("_normal|", "_other|") if rng.random() < 0.4 else x for x in ids])
            wins.append([seqs, ids, w[2], w[3], w[4]])
        kw = dict(readcutoff=int(rng.integers(2, 5)), hcutoff=int(rng.integers(2, 5)),
                  scutoff=float(rng.choice([0.05, 0.2]))) if rng.random() < 0.3 else {}
        flags = [str(rng.choice(["NormalOutput", "UnspanedSV"])) for _ in wins]
        got = batch.localgraph_batch(wins, ctx=object(), windowFlags=flags, **kw).records
        for w, g, fl in zip(wins, got, flags):
            want = O.decision(w[4], w[0], w[1], w[2], w[3], windowFlag=fl, **kw)
            if [str(x) for x in want] != [str(x) for x in g]:
                print("MISMATCH", dict(seed=args.seed, batch=n, kw=kw), "\n want", want, "\n got ", g, flush=True)
                sys.exit(1)
            n_em += str(g[-1]).endswith("EMOutput")
        n += 1
    print("ok: %d batches (%d EMOutput windows) in %.0f s (seed %d)" % (n, n_em, time.time() - t0, args.seed))


if __name__ == "__main__":
    main()
