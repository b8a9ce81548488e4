"""CUDA partial-order alignment against the CPU oracle (through the C ABI)."""
import json
import os

import numpy as np
import pytest

from svscope_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    # the process-wide default context (also used by spoa.poa / Decision / EMCluster): one device
    # arena for the whole test session instead of one per context
    from svscope_b200 import _lib
    return _lib.Context.default(0)


def _random_group(rng, it, lmax=400):
    L = int(rng.integers(1, lmax))
    n = int(rng.integers(2, 10))
    base = synth._rand_seq(rng, L)
    if it % 3 == 0:
        mot = synth._rand_seq(rng, int(rng.integers(1, 6)))
        base = np.tile(mot, max(1, L // len(mot)))
    seqs = []
    for _ in range(n):
        s = base.copy()
        for _ in range(int(rng.integers(0, 3))):
            p = int(rng.integers(0, len(s) + 1))
            ln = int(rng.integers(1, 40))
            s = np.concatenate([s[:p], s[p + ln:]]) if rng.random() < 0.5 else \
                np.concatenate([s[:p], synth._rand_seq(rng, ln), s[p:]])
        if len(s) == 0:
            s = synth._rand_seq(rng, 3)
        seqs.append(synth._to_str(synth.noisy_copy(rng, s, float(rng.choice([0, 0.05, 0.15, 0.3])))))
    if it % 7 == 0:
        seqs[int(rng.integers(1, n))] = ""
    return seqs


@pytest.mark.parametrize("threads,ring,cols,dp", [(384, 8, 8, 2), (128, 8, 8, 2), (512, 8, 8, 2), (384, 2, 8, 2), (256, 10, 8, 2), (128, 1, 8, 2),
                                                  (512, 3, 8, 2), (128, 24, 8, 2), (128, 2, 8, 2), (256, 5, 8, 2),
                                                  (256, 1, 8, 2), (512, 1, 8, 2)])
def test_alignment_pairs_bit_exact(ctx, oracle, threads, ring, cols, dp):
    """Every alignment (node id, read position) list equals the oracle's, for several CTA
    sizes and ring depths (ring 1 forces almost every non-adjacent predecessor through the
    exported rows in global memory).  All shapes run the window kernel (graph resident on the
    device); the default is 384 threads x 8 columns with a ring of 8 rows (one resident window per SM)."""
    from svscope_b200.poa_api import align_pairs
    ctx.set_option("poa_threads", threads)
    ctx.set_option("ring_rows", ring)
    ctx.set_option("poa_cols", cols)
    ctx.set_option("dp_kernel", dp)
    rng = np.random.default_rng(100 + threads + ring + cols)
    try:
        for it in range(40):
            seqs = _random_group(rng, it)
            o = oracle.PoaSession(1)
            ref = [o.add(s) for s in seqs]
            got = align_pairs(ctx, seqs)
            for a, b in zip(ref, got):
                assert a.shape == b.shape and np.array_equal(a, b)
            o.close()
    finally:
        ctx.set_option("poa_threads", 384)
        ctx.set_option("ring_rows", 8)
        ctx.set_option("poa_cols", 8)
        ctx.set_option("dp_kernel", 2)


def test_frozen_cases_and_known_answers(ctx, golden_dir):
    from svscope_b200.spoa import poa
    cases = json.load(open(os.path.join(golden_dir, "poa_cases.json")))["cases"]
    for cs in cases:
        cons, msa = poa(cs["seqs"], 1)
        assert cons == cs["consensus"]
        assert msa == cs["msa"]
    assert poa(["ACGT"], 1) == ("ACGT", ["ACGT"])
    assert poa(["ACGTACGT", "ACGACGT"], 1)[1] == ["ACGTACGT", "ACG-ACGT"]
    assert poa(["ACGTACGT", "ACGTTACGT", "ACGTTACGT"], 1)[0] == "ACGTTACGT"
    assert poa([], 1) == ("", [])
    assert poa(["", ""], 1) == ("", [])


def test_unsupported_modes_fail_loudly(ctx):
    from svscope_b200._lib import SvsError
    from svscope_b200.spoa import poa
    with pytest.raises(SvsError):
        poa(["ACGT", "ACGT"], 0)                 # local alignment is off the hot path
    with pytest.raises(SvsError):
        poa(["ACGT", "ACGT"], 1, g=-4, e=-4)     # linear gaps


def test_windows_msa_and_consensus_equal_oracle(ctx, oracle):
    """Batched groups, multi-pass reads (longer than one 2048-column strip), tandem repeats."""
    from svscope_b200._lib import ReadSet
    from svscope_b200.poa_api import poa_groups
    wins = [synth.make_small_window(s, body_len=b, sv_len=b // 4, n_tumor=5, n_normal=5, n_carriers=3,
                                    sv_type=t) for s, b, t in [(1, 300, "DEL"), (2, 900, "INS"), (3, 2600, "DEL")]]
    wins.append(synth.make_c3(seed=3, total_len=1200, n_tumor=5, n_normal=5, n_carriers=3))
    seqs, groups = [], []
    for w in wins:
        groups.append(list(range(len(seqs), len(seqs) + len(w[0]))))
        seqs += w[0]
    reads = ReadSet(ctx, seqs)
    cons, msas, st = poa_groups(ctx, reads, groups)
    for w, c, m in zip(wins, cons, msas):
        oc, om = oracle.poa(w[0], 1)
        assert c == oc
        assert m == om
        assert [r.replace("-", "") for r in m] == w[0]          # size-independent round trip
    assert st["alignments"] == sum(len(w[0]) - 1 for w in wins)
    # idempotence: the same call again gives the same answer
    cons2, msas2, _ = poa_groups(ctx, reads, groups)
    assert cons2 == cons and msas2 == msas
    reads.close()


def test_high_indegree_and_long_gaps(ctx, oracle):
    """Many distinct alternatives at one site (large in-degree) and gaps longer than any
    ring depth."""
    from svscope_b200.poa_api import align_pairs
    rng = np.random.default_rng(9)
    left, right = synth._to_str(synth._rand_seq(rng, 40)), synth._to_str(synth._rand_seq(rng, 40))
    seqs = [left + right]
    for k in range(1, 14):
        seqs.append(left + synth._to_str(synth._rand_seq(rng, k * 3)) + right)
    seqs.append(left[:20] + right[20:])
    o = oracle.PoaSession(1)
    ref = [o.add(s) for s in seqs]
    got = align_pairs(ctx, seqs)
    for a, b in zip(ref, got):
        assert np.array_equal(a, b)
    assert max(o.graph()["indeg"]) >= 5


def test_exact_pruning_equals_unpruned_and_oracle(ctx, oracle):
    """Score-bound pruning (scout pass + provable band) must not change a single alignment:
    pruned == unpruned on long windows (DEL, INS, tandem repeat, an unrelated read that makes
    the scout pass fail), and == the oracle where the oracle is affordable."""
    from svscope_b200._lib import ReadSet
    from svscope_b200.poa_api import poa_groups
    rng = np.random.default_rng(17)
    wins = [synth.make_small_window(70, body_len=1400, sv_len=300, n_tumor=6, n_normal=6, n_carriers=3),
            synth.make_small_window(71, body_len=2600, sv_len=700, n_tumor=6, n_normal=6, n_carriers=3, sv_type="INS"),
            synth.make_c3(seed=7, total_len=3000, n_tumor=6, n_normal=6, n_carriers=3),
            synth.make_small_window(72, body_len=5200, sv_len=1500, n_tumor=5, n_normal=5, n_carriers=3, sv_type="DEL"),
            synth.make_small_window(73, body_len=6000, sv_len=900, n_tumor=5, n_normal=5, n_carriers=2, sv_type="INS", err=0.12)]
    wins[0][0][4] = synth._to_str(synth._rand_seq(rng, 1500))      # unrelated read: no path in the narrow band
    wins[1][0][3] = wins[1][0][3][:1100]                           # truncated read
    seqs, groups = [], []
    for w in wins:
        groups.append(list(range(len(seqs), len(seqs) + len(w[0]))))
        seqs += w[0]
    reads = ReadSet(ctx, seqs)
    try:
        ctx.set_option("prune", 1)
        cons1, msa1, st1 = poa_groups(ctx, reads, groups)
        ctx.set_option("prune", 0)
        cons0, msa0, st0 = poa_groups(ctx, reads, groups)
    finally:
        ctx.set_option("prune", 1)
    assert cons1 == cons0
    assert msa1 == msa0
    for w, c, m in zip(wins[:3], cons1, msa1):
        oc, om = oracle.poa(w[0], 1)
        assert c == oc and m == om
    reads.close()


def test_bench_windows_equal_committed_oracle_goldens(ctx, golden_dir):
    """Windows of the bench's own configs[1] batch (synth.make_c2_window, full size: 60 reads of
    5-15 kb) against records / MSA / consensus the CPU oracle produced (oracle/gen_golden_c2.py,
    digests committed in tests/golden/c2_windows.json)."""
    import hashlib
    from svscope_b200.batch import localgraph_batch
    from svscope_b200.spoa import poa
    gold = json.load(open(os.path.join(golden_dir, "c2_windows.json")))["windows"]
    assert len(gold) >= 8
    wins = [synth.make_c2_window(g["index"]) for g in gold]
    recs = localgraph_batch(wins, ctx=ctx).records
    for g, w, rec in zip(gold, wins, recs):
        line = "\t".join(str(x) for x in rec)
        assert hashlib.sha256(line.encode()).hexdigest() == g["record_sha256"], g["index"]
    for g, w in list(zip(gold, wins))[:3]:     # MSA and consensus of the window graph itself
        cons, msa = poa(w[0], 1)
        assert len(msa[0]) == g["msa_cols"]
        assert hashlib.sha256("\n".join(msa).encode()).hexdigest() == g["msa_sha256"]
        assert hashlib.sha256(cons.encode()).hexdigest() == g["consensus_sha256"]


def test_window_with_more_than_31_in_edges_fails_alone(ctx, oracle):
    """A node with more than 31 in-edges is beyond the 5-bit in-edge index of the traceback codes:
    that group is reported (status 10) and the other groups of the call are unaffected."""
    from svscope_b200._lib import ReadSet
    from svscope_b200.poa_api import poa_groups
    rng = np.random.default_rng(31)
    left, right = synth._to_str(synth._rand_seq(rng, 30)), synth._to_str(synth._rand_seq(rng, 30))
    body = synth._to_str(synth._rand_seq(rng, 120))
    # deletions of 1..45 bases that all end at the same position: that node collects an in-edge per length
    bad = [body + right] + [body[:len(body) - k] + right for k in range(1, 46)]
    good = [synth._to_str(synth._rand_seq(rng, 50))] * 2 + [left + right, left + "ACGT" + right]
    reads = ReadSet(ctx, good + bad + good)
    groups = [list(range(0, 4)), list(range(4, 4 + len(bad))), list(range(4 + len(bad), 8 + len(bad)))]
    cons, msas, st = poa_groups(ctx, reads, groups, strict=False)
    o = oracle.PoaSession(1)
    for s in bad:
        o.add(s)
    assert max(o.graph()["indeg"]) > 31
    assert list(st["status"]) == [0, 10, 0]
    oc, om = oracle.poa(good, 1)
    assert cons[0] == oc and msas[0] == om and cons[2] == oc and msas[2] == om
    assert cons[1] == "" and msas[1] == []
    with pytest.raises(Exception):
        poa_groups(ctx, reads, groups)        # strict: the caller of a single group gets an error
    reads.close()


def test_full_size_deep_tandem_repeat_window_properties(ctx, c3_full_gpu):
    """configs[2] at FULL size (120 reads of 20 kb, 10 % error, tandem-repeat expansion): the graph
    outgrows the tier-0 scratch slots, so this exercises the memory tiers.  Size-independent properties:
    every MSA row spells its read, the rows are equally long, the consensus is a path of plausible length,
    and a second CTA shape gives the identical result.  The comparison with the CPU oracle (row-checkpoint
    engine, committed digests) is tests/test_zz_gpu_full_c3_golden.py; the scaled-down window against the
    flat oracle is test_full_size_windows_equal_oracle_golden."""
    from svscope_b200._lib import ReadSet
    from svscope_b200.poa_api import poa_groups
    seqs, cons, msas, st = (c3_full_gpu[k] for k in ("seqs", "cons", "msas", "st"))
    assert len(seqs) == 121 and min(len(s) for s in seqs[1:]) > 15_000
    assert st["failed_groups"] == 0
    msa = msas[0]
    assert len(msa) == len(seqs) and len({len(r) for r in msa}) == 1
    assert [r.replace("-", "") for r in msa] == seqs
    assert 0.8 * len(seqs[0]) < len(cons[0]) < 1.6 * len(seqs[0])
    reads = ReadSet(ctx, seqs)
    try:
        ctx.set_option("poa_threads", 512)
        cons2, msas2, _ = poa_groups(ctx, reads, [list(range(len(seqs)))])
        assert cons2 == cons and msas2 == msas
    finally:
        ctx.set_option("poa_threads", 384)
        reads.close()
