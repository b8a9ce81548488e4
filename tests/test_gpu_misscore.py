"""GPU parity of svs_misscore_pairs / svscope_b200.PairwiseCompare (next row F1) against the
oracle and the golden vectors made with the reference's own PairwiseCompare."""
import json
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def ctx():
    from svscope_b200 import _lib
    return _lib.Context.default(0)


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


def _check(ctx, oracle, pairs, par=(1, 0, -1, -1)):
    from svscope_b200.PairwiseCompare import misscore_pairs
    out, lines = misscore_pairs(pairs, *par, want_lines=True, ctx=ctx)
    plain = misscore_pairs(pairs, *par, ctx=ctx)
    assert np.array_equal(out, plain)
    for k, (a, b) in enumerate(pairs):
        r = oracle.pairwise_first_alignment(a, b, *par, want_line=True)
        assert tuple(out[k]) == (r["score"], r["length"], r["matches"], r["length"] - r["matches"]), (k, len(a), len(b))
        assert lines[k] == r["line"], (k, len(a), len(b))


def test_golden_pairs(ctx, oracle):
    from svscope_b200.PairwiseCompare import misscore_pairs
    gold = json.load(open(os.path.join(GOLD, "misscore_pairs.json")))
    by_par = {}
    for c in gold["cases"]:
        by_par.setdefault(tuple(c["params"]), []).append(c)
    for par, cases in by_par.items():
        out, lines = misscore_pairs([(c["a"], c["b"]) for c in cases], *par, want_lines=True, ctx=ctx)
        for k, c in enumerate(cases):
            assert tuple(out[k]) == (c["score"], c["columns"], c["matches"], c["columns"] - c["matches"])
            assert lines[k] == c["line"]


@pytest.mark.parametrize("par", [(1, 0, -1, -1), (1, 0, 0, 0), (2, -1, -2, -2), (1, -3, -1, -1), (5, -4, -3, -3)])
def test_random_small_pairs(ctx, oracle, par):
    rng = random.Random(hash(par) & 0xffff)
    pairs = []
    for it in range(300):
        alpha = rng.choice(["ACGT", "AC", "A", "ACGT-"])
        a = "".join(rng.choice(alpha) for _ in range(rng.randint(1, 200)))
        b = _mutate(rng, a, rng.choice([0.05, 0.2, 0.5])) or "A"
        if it % 5 == 0:
            b = "".join(rng.choice(alpha) for _ in range(rng.randint(1, 200)))
        pairs.append((a, b))
    pairs += [("A", "A"), ("A", "C"), ("A", "ACGTACGTACGTACGTACGT"), ("ACGTACGTACGTACGTAC", "G"), ("-", "ACGT"), ("-", "-")]
    _check(ctx, oracle, pairs, par)


def test_strip_edges_and_multi_strip(ctx, oracle):
    """Widths around the 16-column thread tile, the 4096-column strip, and several strips."""
    rng = random.Random(17)
    pairs = []
    base = "".join(rng.choice("ACGT") for _ in range(9000))
    for lb in (15, 16, 17, 31, 32, 33, 4095, 4096, 4097, 4112, 8191, 8193):
        b = base[:lb]
        a = _mutate(rng, b[: max(1, lb // 3)] + b[lb // 3 + min(40, lb // 4):], 0.02) or "A"
        pairs.append((a[:700] if lb > 5000 else a, b))
    # tall and narrow / short and wide
    pairs.append((base[:6000], _mutate(rng, base[1000:1300], 0.05)))
    pairs.append((_mutate(rng, base[200:500], 0.05), base[:8800]))
    _check(ctx, oracle, pairs)


def test_consensus_sized_pair_and_identities(ctx, oracle):
    """10-12 kb consensus pair with a 1.5 kb deletion (configs[1] scale): equal to the oracle,
    and the (1, 0, -1, -1) identities gaps = 2*columns - la - lb, score = matches - gaps."""
    from svscope_b200.PairwiseCompare import misscore_pairs, AligmentScore
    rng = random.Random(23)
    ger = "".join(rng.choice("ACGT") for _ in range(11000))
    som = _mutate(rng, ger[:4000] + ger[5500:], 0.01)
    stats = {}
    out = misscore_pairs([(som, ger), (ger, som)], ctx=ctx, stats=stats)
    for k, (a, b) in enumerate([(som, ger), (ger, som)]):
        score, cols, matches, mis = (int(x) for x in out[k])
        assert score == matches - (2 * cols - len(a) - len(b))
        assert mis == cols - matches
        r = oracle.pairwise_first_alignment(a, b)
        assert (score, cols, matches) == (r["score"], r["length"], r["matches"])
    assert stats["cells"] == 2.0 * len(som) * len(ger) and stats["kernel_ms"] > 0
    assert AligmentScore(som, ger) == int(out[0, 3])
    assert AligmentScore(som, ger, cutoff=50) == oracle.aligment_score(som, ger, cutoff=50)


def test_pipe_golden_from_reference(ctx, tmp_path):
    """Raw.bed -> Somatic records: equal to what the reference's MisScorePipe returned
    (oracle/gen_golden_misscore.py)."""
    from svscope_b200 import PairwiseCompare as PC
    gold = json.load(open(os.path.join(GOLD, "misscore_pipe.json")))
    path = tmp_path / "x.Raw.bed"
    path.write_text(gold["raw_bed"])
    res = PC.MisScorePipe(str(path))
    rows = [[str(v) if c in ("AF", "window", "chrom", "somSupportReadID", "germSupportReadID") else int(v)
             for c, v in zip(res.columns, r)] for r in res.itertuples(index=False)]
    assert list(res.columns) == gold["columns"]
    assert rows == gold["rows"]


def test_errors_are_loud(ctx):
    from svscope_b200 import _lib
    from svscope_b200.PairwiseCompare import misscore_pairs, AligmentScore, CalculateMisscore
    with pytest.raises(_lib.SvsError):
        misscore_pairs([("ACGT", "ACG")], 1, 0, -2, -1, ctx=ctx)      # affine: unsupported, not approximated
    with pytest.raises(_lib.SvsError):
        misscore_pairs([("ACGT", "")], ctx=ctx)
    with pytest.raises(IndexError):
        AligmentScore("ACGT", "")
    with pytest.raises(IndexError):
        CalculateMisscore({"somSeqList": "ACGT;", "germSeqList": "ACGT"})
    assert misscore_pairs([], ctx=ctx).shape == (0, 4)


def test_small_arena_runs_in_rounds_and_reports_capacity(oracle):
    """A context with a 48 MiB arena: 120 pairs of ~2 kb (1-2 MiB of trace each) need several
    launches and still equal the oracle; a pair whose trace does not fit is an error, not a
    silent truncation."""
    from svscope_b200 import _lib
    from svscope_b200.PairwiseCompare import misscore_pairs
    small = _lib.Context(0, arena_mb=48)
    try:
        rng = random.Random(31)
        pairs = []
        for _ in range(120):
            a = "".join(rng.choice("ACGT") for _ in range(rng.randint(1500, 2200)))
            pairs.append((a, _mutate(rng, a, 0.04)))
        stats = {}
        out = misscore_pairs(pairs, ctx=small, stats=stats)
        assert stats["launches"] >= 3
        for k in (0, 17, 63, 119):
            r = oracle.pairwise_first_alignment(*pairs[k])
            assert tuple(out[k][:3]) == (r["score"], r["length"], r["matches"])
        big = "".join(rng.choice("ACGT") for _ in range(12000))
        with pytest.raises(_lib.SvsError, match="MiB of trace"):
            misscore_pairs([(big, big)], ctx=small)
    finally:
        small.close()
