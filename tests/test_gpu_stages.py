"""Feature selection, mixture model, edit distance and the whole Decision on the GPU against
the oracle and the golden vectors made from the reference (through the C ABI)."""
import os

import numpy as np
import pytest

from svscope_b200 import synth

pytestmark = pytest.mark.gpu

WINDOWS = ["window_del", "window_ins", "window_nosv", "window_shallow", "window_lowerr", "window_emptyreads"]
LOGLIK_RTOL = 1e-6   # tolerance stated by BASELINE.json north_star for mixture-model log-likelihoods


@pytest.fixture(scope="module")
def ctx():
    # the process-wide default context (also used by spoa.poa / Decision / EMCluster): one device
    # arena for the whole test session instead of one per context
    from svscope_b200 import _lib
    return _lib.Context.default(0)


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=True)


def test_feature_kernel_equals_oracle(ctx, oracle):
    from svscope_b200 import batch
    rng = np.random.default_rng(4)
    encs, drops, cuts = [], [], []
    for _ in range(12):
        n, w = int(rng.integers(1, 40)), int(rng.integers(1, 700))
        enc = rng.choice(5, size=(n, w), p=[.3, .3, .2, .15, .05]).astype(np.int64)
        enc[:, rng.random(w) < 0.5] = rng.integers(0, 5)
        drop = (rng.random(w) < 0.1).astype(np.uint8)
        encs.append(enc)
        drops.append(drop)
        cuts.append(float(rng.choice([1, 2, 3, 3.05, 4.5])))
    got = batch.msa_features(ctx, encs, drops, cuts)
    for enc, drop, cut, (keep, nf, zp, ident) in zip(encs, drops, cuts, got):
        cols = oracle.find_non_same_site(enc, cutoff=cut)
        expect = np.zeros(enc.shape[1], bool)
        expect[cols] = True
        expect &= drop == 0
        assert np.array_equal(keep, expect)
        assert nf == int(expect.sum())
        X = enc[:, expect]
        assert zp == oracle.zero_param_num(X)
        if nf:
            sim = ident.astype(np.float64) / nf
            np.fill_diagonal(sim, 1.0)
            assert np.array_equal(sim, oracle.pairwise_identity(X))   # identical integer counts


@pytest.mark.parametrize("name", ["em_a", "em_b", "em_c", "em_d", "em_e", "em_f"])
def test_emcluster_matches_reference_vectors(ctx, golden_dir, name):
    """EMCluster drop-in vs the reference's own output: K, assignments exact; BIC (2*loglik -
    penalty) and gamma within the stated tolerance.  em_c exercises the Dirichlet re-draw."""
    from svscope_b200 import ReadsCluster
    g = _load(golden_dir, name)
    np.random.seed(2023)
    K, _, labels, theta, gamma, pie, bics = ReadsCluster.EMCluster(g["X"].copy(), initselection=1)
    assert K == int(g["K"])
    assert np.array_equal(labels, g["Rclust"])
    np.testing.assert_allclose(bics, g["bics"], rtol=LOGLIK_RTOL)
    np.testing.assert_allclose(gamma, g["gamma"], rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(pie, g["pie"], rtol=1e-9)
    np.testing.assert_allclose(theta, g["theta"], rtol=1e-9, atol=1e-300)
    assert np.array_equal(ReadsCluster.pariwiseDistance(g["X"]), g["sim"])


def test_em_loglik_against_oracle_random(ctx, oracle):
    from svscope_b200 import batch
    from scipy.cluster.hierarchy import fcluster, linkage
    rng = np.random.default_rng(8)
    for N, nf, nsom in [(30, 333, 9), (61, 700, 20), (10, 25, 3)]:
        base = rng.integers(0, 4, nf)
        X = np.tile(base, (N, 1))
        noise = rng.random((N, nf)) < 0.08
        X[noise] = rng.integers(0, 5, int(noise.sum()))
        X[N - nsom:, nf // 4: nf // 2] = 4
        Z = linkage(oracle.pairwise_identity(X), "ward")
        for K in (1, 2, 3):
            np.random.seed(1)
            ref = oracle.em_fit(K, X, Z)
            if ref["n_fallback"]:
                continue
            labels = (fcluster(Z, K, criterion="maxclust") - 1).astype(np.int32)
            got = batch.em_batch(ctx, [X], [batch.EmTaskSpec(0, K, labels)], want_theta=True)[0]
            assert got["status"] == -1
            np.testing.assert_allclose(got["lik"], ref["lik"], rtol=LOGLIK_RTOL)
            np.testing.assert_allclose(got["gamma"], ref["gamma"], rtol=1e-6, atol=1e-12)
            np.testing.assert_allclose(got["theta"], ref["theta"], rtol=1e-9, atol=1e-300)
            assert np.array_equal(np.argmax(got["gamma"], 1), np.argmax(ref["gamma"], 1))


def test_edit_distance_equals_textbook_dp(ctx, oracle):
    from svscope_b200 import batch
    from svscope_b200._lib import ReadSet
    rng = np.random.default_rng(5)
    seqs = []
    for L in [0, 1, 31, 32, 33, 64, 100, 1023, 1024, 1025, 2500, 5000]:
        seqs.append(synth._to_str(synth._rand_seq(rng, L)))
    base = synth._rand_seq(rng, 3000)
    seqs += [synth._to_str(synth.noisy_copy(rng, base, 0.1)) for _ in range(4)]
    reads = ReadSet(ctx, seqs)
    mats, st = batch.edit_distance_matrices(ctx, reads, [list(range(len(seqs)))])
    m = mats[0]
    expect = oracle.levenshtein_matrix(seqs, bitparallel=False)
    assert np.array_equal(m, expect)
    assert st["cells"] > 0 and st["ms"] > 0
    reads.close()


def test_edit_distance_multi_strip(ctx, oracle):
    """Patterns longer than one 32768-row strip; checked against the oracle's bit-parallel
    routine (itself checked against the DP) and by the triangle/identity properties."""
    from svscope_b200 import batch
    from svscope_b200._lib import ReadSet
    rng = np.random.default_rng(6)
    base = synth._rand_seq(rng, 40_000)
    seqs = [synth._to_str(base), synth._to_str(synth.noisy_copy(rng, base, 0.1)),
            synth._to_str(synth.noisy_copy(rng, base, 0.1))]
    reads = ReadSet(ctx, seqs)
    m = batch.edit_distance_matrices(ctx, reads, [[0, 1, 2]])[0][0]
    assert np.array_equal(m, oracle.levenshtein_matrix(seqs, bitparallel=True))
    assert m[0, 0] == 0 and m[1, 2] <= m[0, 1] + m[0, 2]
    reads.close()


@pytest.mark.parametrize("name", WINDOWS)
def test_decision_equals_reference_record(ctx, golden_dir, name):
    """10-field record identical to the one the reference's Decision produced (golden)."""
    from svscope_b200.DecisionMaker import Decision
    g = _load(golden_dir, name)
    np.random.seed(2023)
    rec = Decision(str(g["rec"]), list(g["seqs"]), np.array(g["ids"]), str(g["f5"]), str(g["f3"]))
    assert [str(x) for x in rec] == list(g["record"])


@pytest.mark.parametrize("name", [w for w in WINDOWS if w != "window_shallow"])
def test_feature_selection_equals_reference(ctx, golden_dir, name):
    from svscope_b200.DataScanner import CallMargin, FindNonSameSite, MSAFeatureSelection, SeqDecoder
    g = _load(golden_dir, name)
    enc, X, ids = MSAFeatureSelection(list(g["seqs"]), str(g["f5"]), str(g["f3"]), np.array(g["ids"]))
    assert np.array_equal(enc, g["enc"])
    assert np.array_equal(X, g["X"])
    assert list(ids) == list(g["ids2"])
    assert np.array_equal(CallMargin(list(g["msa"]), str(g["f5"]), str(g["f3"])), g["margin"])
    nonempty = [s for s in g["seqs"] if len(s)]
    assert [SeqDecoder(r) for r in enc[:len(nonempty)]] == nonempty
    if X.shape[1]:
        assert len(FindNonSameSite(X, cutoff=1)) <= X.shape[1]


def test_batch_equals_oracle_and_per_window(ctx, oracle):
    """A mixed batch through localgraph_batch: records equal the oracle's Decision for every
    window, and equal the one-window-at-a-time path; edit distances equal the DP."""
    from svscope_b200.batch import localgraph_batch
    from svscope_b200.SomTDDetector import TDscope_npz
    wins = [synth.make_small_window(40 + k, body_len=int(b), sv_len=int(b) // 4, n_tumor=8, n_normal=8,
                                    n_carriers=int(c), sv_type=t)
            for k, (b, c, t) in enumerate([(300, 4, "DEL"), (500, 5, "INS"), (250, 0, "DEL"), (400, 8, "INS"),
                                           (350, 3, "DEL")])]
    wins.append(synth.make_small_window(77, body_len=200, sv_len=40, n_tumor=2, n_normal=4, n_carriers=2))
    wins.append(synth.make_c3(seed=5, total_len=500, n_tumor=8, n_normal=8, n_carriers=4, err=0.08))
    out = localgraph_batch(wins, ctx=ctx, edit_distance=True, keep_aux=True)
    n_em = 0
    for w, rec, d in zip(wins, out.records, out.edit_distances):
        expect = oracle.decision(w[4], w[0], w[1], w[2], w[3])
        assert [str(x) for x in rec] == [str(x) for x in expect]
        n_em += rec[-1].endswith("EMOutput")
        if d is not None:
            assert np.array_equal(d, oracle.levenshtein_matrix(w[0][1:], bitparallel=False))
    assert n_em >= 2
    single = TDscope_npz(wins[0][4], wins[0][0], wins[0][1], wins[0][2], wins[0][3])
    assert single == out.records[0]
    assert out.stats["poa_cells"] > 0 and out.stats["windows"] == len(wins)


def test_raw_bed_roundtrip(ctx, oracle, tmp_path):
    """localGraph_npz: npz batches in, sorted 10-column Raw.bed out, --Continue appends only
    the missing windows; bytes equal the oracle's records after the same sort."""
    import argparse
    from svscope_b200 import SVscope
    wins = [synth.make_small_window(60 + k, body_len=260, sv_len=70, n_tumor=7, n_normal=7, n_carriers=4)
            for k in range(4)]
    d = str(tmp_path)
    synth.save_npz(os.path.join(d, "T.vs.N.TandemRepeat.batch0.npz"), wins[:3])
    args = argparse.Namespace(savedir=d, TSampleID="T", NSampleID="N", Continue=False, thread="1", offset=50, mapQ=5)
    path = SVscope.localGraph_npz(args)
    lines = open(path).read().splitlines()
    expect = sorted((oracle.format_record(oracle.decision(w[4], w[0], w[1], w[2], w[3])).rstrip("\n") for w in wins[:3]),
                    key=lambda s: (s.split("\t")[0], int(s.split("\t")[1])))
    assert lines == expect
    assert all(len(x.split("\t")) == 10 for x in lines)
    synth.save_npz(os.path.join(d, "T.vs.N.TandemRepeat.batch1.npz"), wins[3:])
    args.Continue = True
    SVscope.localGraph_npz(args)
    assert len(open(path).read().splitlines()) == 4


@pytest.mark.parametrize("name", ["large_c1", "large_c3_scaled"])
def test_full_size_windows_equal_oracle_golden(ctx, golden_dir, name):
    """BASELINE.json configs[0] at full size (30+30 reads ~10 kb, 2 kb somatic DEL) and a scaled
    configs[2] (tandem-repeat INS, 10 % error): record, MSA and consensus equal the oracle's
    (tests/golden/large_*.json, made by oracle/gen_golden_large.py; minutes of CPU there)."""
    import hashlib
    import json
    from svscope_b200.DecisionMaker import Decision
    from svscope_b200.spoa import poa
    g = json.load(open(os.path.join(golden_dir, name + ".json")))
    w = synth.make_c1(seed=1) if name == "large_c1" else \
        synth.make_c3(seed=3, total_len=6000, n_tumor=20, n_normal=20, n_carriers=10, err=0.10)
    cons, msa = poa(w[0], 1)
    assert len(msa[0]) == g["msa_cols"]
    assert hashlib.sha256("\n".join(msa).encode()).hexdigest() == g["msa_sha"]
    assert hashlib.sha256(cons.encode()).hexdigest() == g["consensus_sha"]
    assert [r.replace("-", "") for r in msa] == w[0]
    np.random.seed(2023)
    rec = Decision(w[4], w[0], w[1], w[2], w[3])
    assert [str(x) for x in rec] == g["record"]


def test_edit_distance_pairs_and_alu_probe(ctx, oracle):
    """The pair form of the edit-distance entry point, and the integer-ALU probe bench.py uses."""
    import ctypes
    from svscope_b200._lib import ReadSet, load, ptr
    rng = np.random.default_rng(12)
    base = synth._rand_seq(rng, 1500)
    seqs = [synth._to_str(synth.noisy_copy(rng, base, 0.08)) for _ in range(5)] + ["", "ACGT"]
    reads = ReadSet(ctx, seqs)
    a = np.array([0, 1, 2, 5, 6, 3], np.int64)
    b = np.array([1, 2, 4, 0, 5, 3], np.int64)
    dist = np.zeros(len(a), np.int32)
    stats = np.zeros(5, np.float64)
    ctx.check(load().svs_edit_distance_pairs(ctx._h, reads._h, ptr(a), ptr(b), len(a), ptr(dist), ptr(stats), 5))
    expect = [oracle.levenshtein(seqs[i], seqs[j]) for i, j in zip(a, b)]
    assert dist.tolist() == expect
    assert dist[5] == 0 and dist[3] == len(seqs[0]) and dist[4] == 4
    rates = ctx.int_alu_probe()
    assert rates["max"] > 1000 and rates["addmax"] > 1000      # G thread-operations per second
    reads.close()
