"""Golden records for windows of the bench's own configs[1] batch (svscope_b200.synth.make_c2_window),
made by the CPU oracle (oracle.decision + the window MSA / consensus of oracle.poa).  Full-size
windows: minutes of CPU each, so only digests of the large outputs are stored.

    python oracle/gen_golden_c2.py [n_windows]      ->  tests/golden/c2_windows.json
"""
import hashlib
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(index):
    from oracle import oracle as O
    from svscope_b200 import synth
    t0 = time.time()
    w = synth.make_c2_window(index)
    rec = O.decision(w[4], w[0], w[1], w[2], w[3])
    cons, msa = O.poa(w[0], 1)
    return dict(index=index, ref_len=len(w[0][0]), n_seqs=len(w[0]),
                record=[str(x) for x in rec],
                record_sha256=hashlib.sha256("\t".join(str(x) for x in rec).encode()).hexdigest(),
                msa_cols=len(msa[0]), msa_sha256=hashlib.sha256("\n".join(msa).encode()).hexdigest(),
                consensus_sha256=hashlib.sha256(cons.encode()).hexdigest(), consensus_len=len(cons),
                seconds=round(time.time() - t0, 1))


def main():
    from svscope_b200 import synth
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    # indices of the bench batch (rank 0: 0..999), spread over the cheaper two thirds of the cost range
    cand = list(range(0, 240))
    cost = {i: sum(len(s) for s in synth.make_c2_window(i)[0]) for i in cand}
    order = sorted(cand, key=lambda i: cost[i])
    picks = [order[int((k + 0.5) * 0.66 * len(order) / n)] for k in range(n)]
    with mp.get_context("fork").Pool(min(n, os.cpu_count() or 1)) as pool:
        res = pool.map(one, picks, chunksize=1)
    out = dict(generator="oracle/gen_golden_c2.py", numpy=np.__version__, windows=res,
               note="windows of svscope_b200.synth.make_c2_window(index); records are shortened to digests where large")
    for r in res:   # keep the file small: the record's sequence fields can be tens of kb
        r["record"] = [x if len(x) < 200 else "sha256:" + hashlib.sha256(x.encode()).hexdigest() for x in r["record"]]
    with open(os.path.join(ROOT, "tests", "golden", "c2_windows.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(res), "windows;", [r["seconds"] for r in res], "s each")


if __name__ == "__main__":
    main()
