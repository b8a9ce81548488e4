"""Generate tests/golden/* by running the REFERENCE's own Python (imported unmodified from
/root/reference/src) in the build container.  The reference cannot travel to the GPU box, so
the vectors are committed; this script is the recipe that made them.

    python oracle/gen_golden.py            # rewrites tests/golden/

What is pinned by the reference itself:
  * ReadsCluster.EMCluster (+ pariwiseDistance)               -> em_*.npz
  * DataScanner.MSAFeatureSelection / CallMargin / FindNonSameSite / SeqEncoder / SeqDecoder
    and DecisionMaker.Decision, with `spoa.poa` supplied by oracle.poa (pyspoa is absent:
    the POA half stays "parity unpinned")                     -> window_*.npz
What is only frozen (self-golden, to detect drift of the restatement): poa_cases.json,
levenshtein known answers in lev_cases.json.

Stubs: matplotlib / matplotlib.pyplot (imported at ReadsCluster.py:39, unused on the path),
pysam (DataScanner.py:39, unused by Decision), spoa (-> oracle.poa).
"""
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/src"
OUT = os.path.join(ROOT, "tests", "golden")


def import_reference():
    from oracle import oracle as O
    for name in ("matplotlib", "matplotlib.pyplot", "pysam"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    spoa = types.ModuleType("spoa")
    spoa.poa = O.poa
    sys.modules["spoa"] = spoa
    sys.path.insert(0, REF)
    import ReadsCluster  # noqa
    import DataScanner  # noqa
    import DecisionMaker  # noqa
    return ReadsCluster, DataScanner, DecisionMaker


def planted_matrix(seed, N, nf, err, nsom):
    """Survey Appendix C probe: base row 0..3, i.i.d. noise resampled to 0..4, last nsom rows
    carry symbol 4 over the middle half of the columns."""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 4, nf)
    X = np.tile(base, (N, 1))
    noise = rng.random((N, nf)) < err
    X[noise] = rng.integers(0, 5, int(noise.sum()))
    if nsom:
        X[N - nsom:, nf // 4: nf // 4 + nf // 2] = 4
    return X.astype(np.int64)


def main():
    os.makedirs(OUT, exist_ok=True)
    RC, DS, DM = import_reference()
    from oracle import oracle as O
    from svscope_b200 import synth
    import scipy

    versions = dict(numpy=np.__version__, scipy=scipy.__version__)

    # ---- mixture model ------------------------------------------------------------
    em_cases = [
        dict(name="em_a", seed=11, N=24, nf=60, err=0.08, nsom=8),
        dict(name="em_b", seed=12, N=40, nf=120, err=0.10, nsom=12),
        dict(name="em_c", seed=13, N=16, nf=40, err=0.05, nsom=0),
        dict(name="em_d", seed=14, N=12, nf=40, err=0.05, nsom=4),   # RNG fallback fires
        dict(name="em_e", seed=15, N=7, nf=12, err=0.15, nsom=3),    # K range limited by N
        dict(name="em_f", seed=16, N=60, nf=200, err=0.08, nsom=15),
    ]
    for cs in em_cases:
        X = planted_matrix(cs["seed"], cs["N"], cs["nf"], cs["err"], cs["nsom"])
        calls = {"n": 0}
        orig = np.random.dirichlet

        def counting(*a, **k):
            calls["n"] += 1
            return orig(*a, **k)

        np.random.dirichlet = counting
        try:
            np.random.seed(2023)  # per-window convention (ReadsCluster.py:42)
            K, _, Rclust, theta, gamma, pie, bics = RC.EMCluster(X.copy(), initselection=1)
        finally:
            np.random.dirichlet = orig
        sim = RC.pariwiseDistance(X)
        np.savez_compressed(os.path.join(OUT, cs["name"] + ".npz"), X=X, K=K, Rclust=Rclust,
                            theta=theta, gamma=gamma, pie=pie, bics=bics, sim=sim,
                            dirichlet_calls=calls["n"], versions=json.dumps(versions))
        print(cs["name"], "K", K, "fallback draws", calls["n"])

    # ---- feature selection + Decision on small synthetic windows -----------------------
    windows = {
        "window_del": synth.make_small_window(21, body_len=400, sv_len=120, n_tumor=10, n_normal=10, n_carriers=5),
        "window_ins": synth.make_small_window(22, body_len=300, sv_len=90, n_tumor=9, n_normal=8, n_carriers=4, sv_type="INS"),
        "window_nosv": synth.make_small_window(23, body_len=300, sv_len=1, n_tumor=6, n_normal=6, n_carriers=0),
        "window_shallow": synth.make_small_window(24, body_len=200, sv_len=50, n_tumor=2, n_normal=5, n_carriers=2),
        "window_lowerr": synth.make_small_window(25, body_len=200, sv_len=3, n_tumor=5, n_normal=5, n_carriers=0, err=0.002),
    }
    # a window with fully deleted (empty) reads: exercises DataScanner.py:198-209
    w = synth.make_small_window(26, body_len=250, sv_len=80, n_tumor=8, n_normal=8, n_carriers=4)
    w[0][3] = ""
    w[0][12] = ""
    windows["window_emptyreads"] = w
    for name, w in windows.items():
        seqs, ids, f5, f3, rec = w
        out = dict(seqs=np.array(seqs, dtype=object), ids=np.array(ids), f5=f5, f3=f3, rec=rec,
                   versions=json.dumps(versions))
        np.random.seed(2023)
        record = DM.Decision(rec, list(seqs), np.array(ids), f5, f3)
        out["record"] = np.array([str(x) for x in record], dtype=object)
        tags, cnt = np.unique([x.split("|")[0].split("_")[-1] for x in ids], return_counts=True)
        if len(seqs) > 3 and len(tags) >= 2 and cnt.min() >= 3:
            enc, X, ids2 = DS.MSAFeatureSelection(list(seqs), f5, f3, np.array(ids))
            cons, msa = O.poa(list(seqs), 1)
            out.update(enc=np.asarray(enc), X=np.asarray(X), ids2=np.array(ids2),
                       margin=DS.CallMargin(msa, f5, f3),
                       msa=np.array(msa, dtype=object), consensus=cons)
            if X.shape[0] and X.shape[1] >= 10:
                np.random.seed(2023)
                K, _, Rclust, theta, gamma, pie, bics = RC.EMCluster(X.copy(), initselection=1)
                out.update(K=K, Rclust=Rclust, gamma=gamma, pie=pie, bics=bics)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        print(name, record[5], record[8], record[9], "nf", out.get("X", np.zeros((0, 0))).shape)

    # ---- frozen POA cases (unpinned restatement; detects drift) ---------------------------
    rng = np.random.default_rng(5)
    poa_in = [
        ["ACGT"],
        ["ACGT", "ACGT"],
        ["ACGTACGT", "ACGACGT", "ACGTTACGT"],
        ["AAAA", "TTTT"],
        ["GATTACA", "GATACA", "GATTTACA", "CATTACA", "GATTACAT"],
        ["ACGTTGCA", "", "ACGTGCA"],
        ["ACGT" * 10, "ACGT" * 12, "ACGT" * 9, "ACGT" * 10],
    ]
    for _ in range(6):
        base = synth._rand_seq(rng, int(rng.integers(30, 120)))
        poa_in.append([synth._to_str(synth.noisy_copy(rng, base, 0.12)) for _ in range(int(rng.integers(3, 9)))])
    cases = []
    for seqs in poa_in:
        s = O.PoaSession(1)
        alns = [s.add(x).tolist() for x in seqs]
        g = s.graph()
        cases.append(dict(seqs=seqs, consensus=s.consensus(), msa=s.msa(), alignments=alns,
                          rank_node=g["rank_node"].tolist(), in_tail=g["in_tail"].tolist(),
                          in_weight=g["in_weight"].tolist(), indeg=g["indeg"].tolist()))
        s.close()
    with open(os.path.join(OUT, "poa_cases.json"), "w") as fh:
        json.dump(dict(note="self-golden of the unpinned spoa restatement", cases=cases), fh)

    lev = [("kitten", "sitting", 3), ("", "abc", 3), ("abc", "", 3), ("flaw", "lawn", 2),
           ("GATTACA", "GATTACA", 0), ("intention", "execution", 5), ("ACGT", "TGCA", 4)]
    with open(os.path.join(OUT, "lev_cases.json"), "w") as fh:
        json.dump(dict(note="textbook known answers", cases=lev), fh)
    print("golden written to", OUT)


if __name__ == "__main__":
    main()
