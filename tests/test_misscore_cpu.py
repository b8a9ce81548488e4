"""MisScore (next row F1, src/PairwiseCompare.py) without a GPU: the oracle against its golden
vectors, the kernel's cell/traceback logic (shared headers, CPU emulation) against the oracle,
and the host-side record logic against the reference's own PairwiseCompare outputs."""
import io
import json
import os
import random

import numpy as np
import pytest

from oracle import oracle as O
from oracle import pairwise2_oracle as P
from tests.emul import emul

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _mutate(rng, s, rate):
    out = []
    for ch in s:
        r = rng.random()
        if r < rate * 0.4:
            out.append(rng.choice("ACGT"))
        elif r < rate * 0.7:
            out.append(ch)
            out.append(rng.choice("ACGT"))
        elif r < rate:
            pass
        else:
            out.append(ch)
    return "".join(out)


def test_c_oracle_matches_literal_restatement_golden():
    gold = json.load(open(os.path.join(GOLD, "misscore_pairs.json")))
    for c in gold["cases"]:
        r = O.pairwise_first_alignment(c["a"], c["b"], *c["params"], want_line=True)
        assert (r["score"], r["length"], r["matches"], r["line"]) == (c["score"], c["columns"], c["matches"], c["line"])
        # the literal restatement itself still gives the frozen alignment
        lit = P.globalms(c["a"], c["b"], *c["params"], max_alignments=1)[0]
        assert (lit[0], lit[1]) == (c["alignedA"], c["alignedB"])


def test_documentation_examples_order():
    """Two alignments, in this order, in Biopython's documentation of globalxx (recalled)."""
    gold = json.load(open(os.path.join(GOLD, "misscore_pairs.json")))
    for d in gold["doc_examples"]:
        got = [x[1] for x in P.globalms(d["a"], d["b"], *d["params"])]
        assert got == d["alignedB"]
        if d["recalled_from_docs"]:
            assert got == d["recalled_from_docs"]
    assert len(P.globalms("GAACT", "GAT", 1, 0, 0, 0)) == 2


def test_first_alignment_is_first_of_full_list_and_affine_dead_ends():
    rng = random.Random(3)
    pops = 0
    for it in range(150):
        a = "".join(rng.choice("AC") for _ in range(rng.randint(1, 12)))
        b = "".join(rng.choice("AC") for _ in range(rng.randint(1, 12)))
        par = rng.choice([(1, 0, -1, -1), (2, -1, -3, -1), (1, -1, -2, -1), (5, -4, -8, -6)])
        full = P.globalms(a, b, *par)
        first = P.globalms(a, b, *par, max_alignments=1)[0]
        assert full[0] == first
        r = O.pairwise_first_alignment(a, b, *par, want_line=True)
        assert r["line"] == P.match_line(first[0], first[1])
        pops = max(pops, r["pops"])
    assert pops > 1  # affine penalties do hit pairwise2's dead-end rule


def test_linear_gap_first_alignment_never_backtracks():
    """What the kernel's greedy traceback relies on (misscore_tb.h): with open == extend the
    literal stack traversal pops exactly once."""
    rng = random.Random(5)
    for it in range(400):
        alpha = rng.choice(["ACGT", "AC", "A"])
        a = "".join(rng.choice(alpha) for _ in range(rng.randint(1, 60)))
        b = "".join(rng.choice(alpha) for _ in range(rng.randint(1, 60)))
        m, mm, g = rng.choice([(1, 0, 1), (1, 0, 0), (1, -3, 1), (2, -5, 1), (0, -1, 1), (3, -7, 2), (1, -2, 0)])
        assert O.pairwise_first_alignment(a, b, m, mm, -g, -g)["pops"] == 1


@pytest.mark.parametrize("threads,cols", [(256, 16), (2, 4), (3, 4), (1, 16), (5, 4), (2, 16)])
def test_kernel_emulation_matches_oracle(threads, cols):
    rng = random.Random(threads * 100 + cols)
    for it in range(250):
        alpha = rng.choice(["ACGT", "AC", "A", "ACGT-"])
        a = "".join(rng.choice(alpha) for _ in range(rng.randint(1, 150)))
        b = _mutate(rng, a, rng.choice([0.05, 0.2, 0.5])) or "A"
        if it % 5 == 0:
            b = "".join(rng.choice(alpha) for _ in range(rng.randint(1, 150)))
        if it % 11 == 0:
            k = rng.randint(0, len(a))
            b = a[:k] + a[k + rng.randint(1, 50):] or "C"
        m, mm, g = rng.choice([(1, 0, 1), (1, 0, 1), (1, 0, 0), (2, -1, 2), (1, -3, 1), (5, -4, 3), (0, 0, 0)])
        c = O.pairwise_first_alignment(a, b, m, mm, -g, -g, want_line=True)
        e = emul.misscore_emul(a, b, m, mm, g, threads, cols, want_line=True)
        assert e["status"] == 0
        assert (c["score"], c["length"], c["matches"], c["line"]) == (e["score"], e["length"], e["matches"], e["line"]), (a, b)


def test_kernel_emulation_multi_strip_kilobases():
    rng = random.Random(9)
    a = "".join(rng.choice("ACGT") for _ in range(2500))
    b = _mutate(rng, a[:900] + a[1300:], 0.03)
    c = O.pairwise_first_alignment(a, b, want_line=True)
    for threads, cols in ((256, 16), (16, 16), (37, 4)):
        e = emul.misscore_emul(a, b, 1, 0, 1, threads, cols, want_line=True)
        assert (c["score"], c["length"], c["matches"], c["line"]) == (e["score"], e["length"], e["matches"], e["line"])
    # identities of the (1, 0, -1, -1) scoring: gaps = 2*columns - la - lb, score = matches - gaps
    gaps = 2 * c["length"] - len(a) - len(b)
    assert c["score"] == c["matches"] - gaps
    assert O.aligment_score(a, b) == c["length"] - c["matches"]


def test_oracle_misscore_helpers():
    assert O.smaller_absolute_value(-3, 3) == 3 and O.smaller_absolute_value(-2, 3) == -2
    with pytest.raises(IndexError):
        O.aligment_score("", "ACGT")
    assert O.aligment_score("ACGT", "ACGT") == 0
    assert O.aligment_score("ACGT", "AGT") == 1
    assert O.aligment_score("ACGTACGT", "ACGTACGT", cutoff=2) == 0
    assert O.calculate_misscore("ACGTACGT", "ACGTTTACGT;ACGACGT") == 1  # |1| < |-2|, later tie wins otherwise


def test_host_record_logic_matches_reference_golden(monkeypatch):
    """svscope_b200.PairwiseCompare's record reduction / AF / pipe against the reference's own
    PairwiseCompare (tests/golden/misscore_pipe.json); the GPU call is replaced by the oracle
    here, the GPU suite runs the same golden through the kernel."""
    import pandas as pd
    from svscope_b200 import PairwiseCompare as PC
    gold = json.load(open(os.path.join(GOLD, "misscore_pipe.json")))

    def fake_pairs(pairs, *a, stats=None, **kw):
        out = np.zeros((len(pairs), 4), np.int32)
        for k, (x, y) in enumerate(pairs):
            r = O.pairwise_first_alignment(x, y)
            out[k] = (r["score"], r["length"], r["matches"], r["length"] - r["matches"])
        return out

    monkeypatch.setattr(PC, "misscore_pairs", fake_pairs)
    path = os.path.join(os.environ.get("TMPDIR", "/tmp"), "svs_misscore_golden.bed")
    with open(path, "w") as fh:
        fh.write(gold["raw_bed"])
    res = PC.MisScorePipe(path)
    os.unlink(path)
    assert list(res.columns) == gold["columns"]
    rows = [[str(v) if c in ("AF", "window", "chrom", "somSupportReadID", "germSupportReadID") else int(v)
             for c, v in zip(res.columns, r)] for r in res.itertuples(index=False)]
    assert rows == gold["rows"]
    df = pd.read_csv(io.StringIO(gold["raw_bed"]), sep="\t", header=None)
    df.columns = PC.COLUMNS
    em = df.loc[df["flag"] == "NormalOutput|EMOutput"]
    assert [PC.Mismatch_abs(r) for _, r in em.iterrows()] == gold["mismatch_abs"]
    first = em.iloc[0]
    assert PC.CalculateMisscore(first) == gold["rows"][0][6]
    assert PC.smaller_absolute_value(-3, 3) == 3
    with pytest.raises(IndexError):
        PC.AligmentScore("", "ACGT")
