"""Generate tests/golden/misscore_*.json (next row F1, src/PairwiseCompare.py).

    python oracle/gen_golden_misscore.py

misscore_pipe.json — the REFERENCE's own ``PairwiseCompare.MisScorePipe`` / ``CalculateMisscore``
/ ``CallAlleleFreq`` (imported unmodified from /root/reference/src) run on a synthetic Raw.bed,
with ``Bio.pairwise2`` supplied by this repo's restatement (Biopython is absent here): the
record-level logic is pinned by the reference, the aligner underneath it stays **parity
unpinned**.  Stubs: ``statsmodels`` (imported at PairwiseCompare.py:7, unused), ``Bio.Seq.Seq``
(-> str), ``Bio.pairwise2.align.globalms`` / ``format_alignment`` (-> oracle.misscore C oracle).

misscore_pairs.json — alignment-level vectors: the literal Python restatement
(oracle/pairwise2_oracle.py) on small pairs, and the two alignments of the examples of the
Biopython documentation as recalled (``globalxx("ACCGT", "ACG")`` -> ``A-CG-`` then ``AC-G-``);
they are self-golden, not a pin.
"""
import json
import os
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/src"
OUT = os.path.join(ROOT, "tests", "golden")


def import_reference_pairwisecompare():
    from oracle import oracle as O
    for name in ("statsmodels", "statsmodels.stats", "statsmodels.stats.multitest", "Bio", "Bio.Seq",
                 "Bio.pairwise2"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["statsmodels"].stats = sys.modules["statsmodels.stats"]
    sys.modules["statsmodels.stats"].multitest = sys.modules["statsmodels.stats.multitest"]
    sys.modules["Bio.Seq"].Seq = str
    pw = sys.modules["Bio.pairwise2"]

    class _Align:
        @staticmethod
        def globalms(a, b, match, mismatch, open, extend):
            if not a or not b:
                return []
            r = O.pairwise_first_alignment(a, b, match, mismatch, open, extend, want_line=True)
            return [(r["line"],)]

    pw.align = _Align
    pw.format_alignment = lambda line: "seqA\n" + line + "\nseqB\n  Score=0\n"
    sys.modules["Bio"].pairwise2 = pw
    sys.modules["Bio"].Seq = sys.modules["Bio.Seq"]
    sys.path.insert(0, REF)
    import PairwiseCompare
    return PairwiseCompare


def mutate(rng, s, rate):
    out = []
    for ch in s:
        r = rng.random()
        if r < rate * 0.4:
            out.append("ACGT"[rng.integers(4)])
        elif r < rate * 0.7:
            out.append(ch)
            out.append("ACGT"[rng.integers(4)])
        elif r < rate:
            pass
        else:
            out.append(ch)
    return "".join(out)


def synthetic_raw_bed(seed=7, n=14):
    rng = np.random.default_rng(seed)
    rows = []
    for i in range(n):
        L = int(rng.integers(150, 500))
        base = "".join("ACGT"[k] for k in rng.integers(0, 4, L))
        k0 = int(rng.integers(20, L - 60))
        sv = int(rng.integers(10, 50))
        ins = "".join("ACGT"[k] for k in rng.integers(0, 4, sv))
        som_variants = [base[:k0] + base[k0 + sv:], base[:k0] + ins + base[k0:]]
        nsom = int(rng.integers(1, 3))
        ngerm = int(rng.integers(1, 4))
        som = [mutate(rng, som_variants[int(rng.integers(2))], 0.01) for _ in range(nsom)]
        germ = [mutate(rng, base, 0.01) for _ in range(ngerm)]
        som_ids = ";".join(",".join("S_tumor|r%d_%d_%d" % (i, c, j) for j in range(int(rng.integers(3, 7))))
                           for c in range(nsom))
        germ_ids = ";".join(",".join("S_%s|g%d_%d_%d" % ("normal" if j % 2 else "tumor", i, c, j)
                                     for j in range(int(rng.integers(3, 9)))) for c in range(ngerm))
        flag = "NormalOutput|EMOutput" if i % 4 != 3 else "NormalOutput"
        if flag == "NormalOutput":
            rows.append(["chr%d" % (1 + i % 3), 1000 * i, 1000 * i + L, "", "", 0, ";".join(germ), germ_ids, ngerm, flag])
        else:
            rows.append(["chr%d" % (1 + i % 3), 1000 * i, 1000 * i + L, ";".join(som), som_ids, nsom,
                         ";".join(germ), germ_ids, ngerm, flag])
    return "".join("\t".join(str(x) for x in r) + "\n" for r in rows)


def main():
    os.makedirs(OUT, exist_ok=True)
    PC = import_reference_pairwisecompare()
    import pandas  # noqa
    import scipy
    text = synthetic_raw_bed()
    with tempfile.NamedTemporaryFile("w", suffix=".bed", delete=False) as fh:
        fh.write(text)
        path = fh.name
    res = PC.MisScorePipe(path)
    os.unlink(path)
    rows = [[str(v) if c in ("AF", "window", "chrom", "somSupportReadID", "germSupportReadID") else int(v)
             for c, v in zip(res.columns, r)] for r in res.itertuples(index=False)]
    # record-level helpers of the reference on the same rows
    import pandas as pd
    import io
    df = pd.read_csv(io.StringIO(text), sep="\t", header=None)
    df.columns = ['chrom', 'start', 'end', 'somSeqList', 'somSupportReadID', 'someventCount', 'germSeqList',
                  'germSupportReadID', 'germeventCount', 'flag']
    em = df.loc[df['flag'] == 'NormalOutput|EMOutput']
    mismatch_abs = [PC.Mismatch_abs(r) for _, r in em.iterrows()]
    json.dump({"generator": "oracle/gen_golden_misscore.py",
               "reference": "PairwiseCompare.MisScorePipe / Mismatch_abs (imported unmodified), aligner = oracle restatement",
               "versions": {"numpy": np.__version__, "pandas": pandas.__version__, "scipy": scipy.__version__},
               "raw_bed": text, "columns": list(res.columns), "rows": rows, "mismatch_abs": mismatch_abs},
              open(os.path.join(OUT, "misscore_pipe.json"), "w"), indent=1)

    from oracle import pairwise2_oracle as P
    rng = np.random.default_rng(11)
    cases = []
    for it in range(60):
        alpha = ["ACGT", "AC", "ACGT-"][it % 3]
        a = "".join(alpha[k] for k in rng.integers(0, len(alpha), int(rng.integers(1, 70))))
        b = mutate(rng, a, [0.05, 0.2, 0.5][it % 3]) or "A"
        if it % 5 == 0:
            b = "".join(alpha[k] for k in rng.integers(0, len(alpha), int(rng.integers(1, 70))))
        par = [(1, 0, -1, -1), (1, 0, -1, -1), (2, -1, -2, -2), (1, -3, -1, -1), (1, 0, 0, 0)][it % 5]
        r = P.globalms(a, b, *par, max_alignments=1)[0]
        line = P.match_line(r[0], r[1])
        cases.append({"a": a, "b": b, "params": list(par), "score": int(r[2]), "columns": len(line),
                      "matches": line.count("|"), "line": line, "alignedA": r[0], "alignedB": r[1]})
    doc = []
    for a, b, par, want in (("ACCGT", "ACG", (1, 0, 0, 0), ["A-CG-", "AC-G-"]),
                            ("GAACT", "GAT", (1, 0, 0, 0), None)):
        r = P.globalms(a, b, *par)
        doc.append({"a": a, "b": b, "params": list(par), "alignedB": [x[1] for x in r], "score": r[0][2],
                    "recalled_from_docs": want})
    json.dump({"generator": "oracle/gen_golden_misscore.py", "cases": cases, "doc_examples": doc},
              open(os.path.join(OUT, "misscore_pairs.json"), "w"), indent=1)
    print("rows", len(rows), "cases", len(cases))
    for r in rows[:3]:
        print(r[0:4], r[6:])


if __name__ == "__main__":
    main()
