"""The CPU oracle against the vectors produced by the reference's own Python
(oracle/gen_golden.py; reference files src/ReadsCluster.py, src/DataScanner.py,
src/DecisionMaker.py).  CPU only."""
import json
import os

import numpy as np
import pytest


def _load(path):
    return np.load(path, allow_pickle=True)


@pytest.mark.parametrize("name", ["em_a", "em_b", "em_c", "em_d", "em_e", "em_f"])
def test_em_cluster_matches_reference(oracle, golden_dir, name):
    g = _load(os.path.join(golden_dir, name + ".npz"))
    X = g["X"]
    out, info = oracle.em_cluster(X.copy(), reseed=True, return_info=True)
    K, _, labels, theta, gamma, pie, bics = out
    assert K == int(g["K"])
    assert np.array_equal(labels, g["Rclust"])           # assignments: exact
    np.testing.assert_allclose(bics, g["bics"], rtol=1e-9, atol=0)
    np.testing.assert_allclose(gamma, g["gamma"], rtol=1e-9, atol=1e-300)
    np.testing.assert_allclose(pie, g["pie"], rtol=1e-12)
    np.testing.assert_allclose(theta, g["theta"], rtol=1e-9, atol=1e-300)
    np.testing.assert_array_equal(info["sim"], g["sim"])
    # the Dirichlet fallback consumes the global RNG exactly like the reference
    assert (info["n_fallback"] > 0) == (int(g["dirichlet_calls"]) > 0)


WINDOWS = ["window_del", "window_ins", "window_nosv", "window_shallow", "window_lowerr",
           "window_emptyreads"]


@pytest.mark.parametrize("name", WINDOWS)
def test_decision_matches_reference(oracle, golden_dir, name):
    g = _load(os.path.join(golden_dir, name + ".npz"))
    rec = oracle.decision(str(g["rec"]), list(g["seqs"]), np.array(g["ids"]), str(g["f5"]), str(g["f3"]))
    assert [str(x) for x in rec] == list(g["record"])


@pytest.mark.parametrize("name", [w for w in WINDOWS if w != "window_shallow"])
def test_feature_selection_matches_reference(oracle, golden_dir, name):
    g = _load(os.path.join(golden_dir, name + ".npz"))
    enc, X, ids = oracle.msa_feature_selection(list(g["seqs"]), str(g["f5"]), str(g["f3"]), np.array(g["ids"]))
    assert np.array_equal(enc, g["enc"])
    assert np.array_equal(X, g["X"])
    assert list(ids) == list(g["ids2"])
    assert np.array_equal(oracle.call_margin(list(g["msa"])[0], str(g["f5"]), str(g["f3"])), g["margin"])
    # SeqDecoder round trip (DataScanner.py:131-137): rows of the MSA de-gap to the inputs
    seqs = [s for s in g["seqs"] if len(s) > 0]
    for row, s in zip(enc[:len(seqs)], seqs):
        assert oracle.decode_row(row) == s


def test_poa_frozen_cases(oracle, golden_dir):
    cases = json.load(open(os.path.join(golden_dir, "poa_cases.json")))["cases"]
    for cs in cases:
        s = oracle.PoaSession(1)
        alns = [s.add(x).tolist() for x in cs["seqs"]]
        assert alns == cs["alignments"]
        assert s.consensus() == cs["consensus"]
        msa = s.msa()
        assert msa == cs["msa"]
        g = s.graph()
        assert g["rank_node"].tolist() == cs["rank_node"]
        assert g["in_tail"].tolist() == cs["in_tail"]
        assert g["in_weight"].tolist() == cs["in_weight"]
        # size-independent properties: every row de-gaps to its input; equal widths
        nonempty = [x for x in cs["seqs"] if x]
        assert [r.replace("-", "") for r in msa] == nonempty
        assert len({len(r) for r in msa}) <= 1
        s.close()


def test_poa_trivial_known_answers(oracle):
    assert oracle.poa(["ACGT"], 1) == ("ACGT", ["ACGT"])
    assert oracle.poa(["ACGT", "ACGT", "ACGT"], 1) == ("ACGT", ["ACGT"] * 3)
    cons, msa = oracle.poa(["ACGTACGT", "ACGACGT"], 1)
    assert msa == ["ACGTACGT", "ACG-ACGT"]
    cons, msa = oracle.poa(["ACGTACGT", "ACGTTACGT", "ACGTTACGT"], 1)
    assert cons == "ACGTTACGT"


def test_levenshtein_known_answers(oracle, golden_dir):
    cases = json.load(open(os.path.join(golden_dir, "lev_cases.json")))["cases"]
    for a, b, d in cases:
        assert oracle.levenshtein(a, b) == d
        assert oracle.levenshtein(a, b, bitparallel=True) == d


def test_levenshtein_bitparallel_equals_dp(oracle):
    rng = np.random.default_rng(3)
    for _ in range(150):
        la, lb = int(rng.integers(0, 300)), int(rng.integers(0, 300))
        a = "".join(rng.choice(list("ACGT"), la))
        b = "".join(rng.choice(list("ACGT"), lb))
        assert oracle.levenshtein(a, b, True) == oracle.levenshtein(a, b, False)
    seqs = ["".join(rng.choice(list("ACGT"), int(rng.integers(1, 150)))) for _ in range(6)]
    m = oracle.levenshtein_matrix(seqs)
    assert np.array_equal(m, m.T) and (np.diag(m) == 0).all()
    assert m[1, 4] == oracle.levenshtein(seqs[1], seqs[4])


def test_levenshtein_agrees_with_an_independent_engine(oracle):
    """The reference's Levenshtein module is absent (parity stays unpinned), but the image has
    the third-party ``regex`` package, whose fuzzy BESTMATCH minimises substitutions +
    insertions + deletions: an implementation independent of ours gives the same distances on
    short, similar strings (its search is exponential on dissimilar ones)."""
    import random
    regex = pytest.importorskip("regex")
    rng = random.Random(1)
    n = 0
    for _ in range(300):
        a = "".join(rng.choice("ACGT") for _ in range(rng.randint(1, 12)))
        b = list(a)
        for _ in range(rng.randint(0, 3)):
            r, p = rng.random(), rng.randrange(len(b) + 1)
            if r < 0.4 and b:
                b[min(p, len(b) - 1)] = rng.choice("ACGT")
            elif r < 0.7:
                b.insert(p, rng.choice("ACGT"))
            elif b:
                b.pop(min(p, len(b) - 1))
        b = "".join(b) or "A"
        m = regex.fullmatch(r"(?b)(?:%s){e<=6}" % a, b)
        if m is None:
            continue
        n += 1
        assert sum(m.fuzzy_counts) == oracle.levenshtein(a, b) == oracle.levenshtein(a, b, bitparallel=True)
    assert n > 250


def _flat_equals_blocked(oracle, seqs, block_rows):
    a, b = oracle.PoaSession(1), oracle.PoaSession(1, block_rows=block_rows)
    try:
        for x in seqs:
            pa, pb = a.add(x), b.add(x)
            assert np.array_equal(pa, pb), (block_rows, len(x))
            if len(pa):
                assert a.score == b.score
        assert a.msa() == b.msa() and a.consensus() == b.consensus()
        return b.blocked_stats
    finally:
        a.close()
        b.close()


def test_row_checkpoint_engine_equals_flat_engine(oracle, golden_dir):
    """The bounded-memory engine (blocks of rows, kept rows, blocks recomputed by the traceback) gives the flat
    five-matrix engine's alignments read after read: frozen cases, windows with deletions / insertions /
    tandem repeats, empty reads, blocks of 1..64 rows (so that blocks, kept rows and recomputation are all hit)."""
    from svscope_b200 import synth
    cases = json.load(open(os.path.join(golden_dir, "poa_cases.json")))["cases"]
    recomputed = kept = 0
    for cs in cases:
        for B in (1, 3, 16):
            st = _flat_equals_blocked(oracle, cs["seqs"], B)
            recomputed += st["recomputed_blocks"]
            kept += st["kept_rows"]
    rng = np.random.default_rng(11)
    for seed in range(12):
        w = synth.make_small_window(100 + seed, body_len=int(rng.integers(30, 260)), sv_len=int(rng.integers(5, 90)),
                                    n_tumor=6, n_normal=6, n_carriers=3, sv_type="INS" if seed % 2 else "DEL")
        seqs = list(w[0])
        if seed % 4 == 0:
            seqs.insert(3, "")
        for B in (1, 2, 5, 64, -1):
            st = _flat_equals_blocked(oracle, seqs, B)
            recomputed += st["recomputed_blocks"]
            kept += st["kept_rows"]
    unit = "ACGGT"
    rep = ["TTGACC" + unit * k + "GGATCA" for k in (6, 9, 6, 12, 9, 7)]
    for B in (1, 4, 9):
        _flat_equals_blocked(oracle, rep, B)
    assert recomputed > 0 and kept > 0


def test_full_size_configs2_golden_is_self_consistent(golden_dir):
    """tests/golden/c3_full.json (oracle/gen_golden_c3.py: the row-checkpoint engine on configs[2] at FULL size)
    names the window the GPU test aligns; the digest fields are present."""
    path = os.path.join(golden_dir, "c3_full.json")
    if not os.path.exists(path):
        pytest.skip("c3_full.json not generated")
    g = json.load(open(path))
    from svscope_b200 import synth
    import hashlib
    w = synth.make_c3(seed=g["seed"])
    assert len(w[0]) == g["n_seqs"]
    assert hashlib.sha256("\n".join(w[0]).encode()).hexdigest() == g["input_sha256"]
    assert len(g["msa_sha256"]) == 64 and len(g["consensus_sha256"]) == 64 and g["msa_cols"] > 20_000
