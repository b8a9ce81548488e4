"""Host-side logic of the product on CPU: rank-order graph + cell arithmetic (through the
sequential emulation of the kernels), flank columns, sharding, world_size-2 gather (gloo)."""
import os

import numpy as np
import pytest

from svscope_b200 import batch, shard, synth


def _random_group(rng, it):
    L = int(rng.integers(1, 220))
    n = int(rng.integers(2, 9))
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


def test_emulated_kernel_path_equals_oracle(oracle):
    """Product host graph + packed-cell arithmetic + code walker == five-matrix oracle."""
    from tests.emul.emul import EmuSession
    rng = np.random.default_rng(11)
    for it in range(120):
        seqs = _random_group(rng, it)
        o, e = oracle.PoaSession(1), EmuSession(ring_rows=int(rng.integers(1, 6)))
        for s in seqs:
            assert np.array_equal(o.add(s), e.add(s))
        assert np.array_equal(o.graph()["rank_node"], e.rank_to_node())
        assert o.consensus() == e.consensus()
        assert o.msa() == e.msa()
        o.close()
        e.close()


def test_pruned_emulation_equals_oracle(oracle):
    """Exact pruning theory on CPU: cells whose score upper bound is below a feasible alignment
    score are set to minus infinity (static bands from path-length intervals, and a dynamic
    variant with a guessed bound and retry); every alignment must stay identical."""
    from tests.emul.emul import EmuSession
    rng = np.random.default_rng(23)
    for it in range(60):
        seqs = _random_group(rng, it)
        if it % 5 == 0:
            seqs[-1] = synth._to_str(synth._rand_seq(rng, int(rng.integers(1, 90))))   # unrelated read
        kw = dict(prune=int(rng.choice([1, 8, 64]))) if it % 2 == 0 else dict(dyn=float(rng.choice([2.0, 4.2, 5.0])))
        o, e = oracle.PoaSession(1), EmuSession(ring_rows=3, **kw)
        for s in seqs:
            assert np.array_equal(o.add(s), e.add(s))
        assert o.msa() == e.msa() and o.consensus() == e.consensus()
        o.close()
        e.close()
    w = synth.make_small_window(3, body_len=1200, sv_len=300, n_tumor=6, n_normal=6, n_carriers=3)
    e = EmuSession(ring_rows=3, prune=32)
    o = oracle.PoaSession(1)
    for s in w[0]:
        assert np.array_equal(o.add(s), e.add(s))
    assert e.kept_fraction() < 0.6          # the bound really prunes
    o.close()
    e.close()


def test_lookahead_pruning_emulation_equals_oracle(oracle):
    """Round-2 candidate, exactness on CPU: prefix-exact pruning ("H + suffix bound >= LB") where a
    row is computed over its predecessors' relevant columns plus a FIXED lookahead of chunks; a
    row whose last computed chunk is still relevant repeats the alignment with more lookahead.
    Alignments stay identical for every lookahead, including 0 (always overflowing)."""
    from tests.emul.emul import EmuSession
    rng = np.random.default_rng(41)
    for it in range(60):
        seqs = _random_group(rng, it)
        if it % 6 == 0:
            seqs[-1] = synth._to_str(synth._rand_seq(rng, int(rng.integers(1, 90))))
        o = oracle.PoaSession(1)
        e = EmuSession(ring_rows=3, dyn=float(rng.choice([2.0, 4.2, 4.6, 5.0])), dyn_ext=int(rng.choice([0, 1, 2, 5])))
        for s in seqs:
            assert np.array_equal(o.add(s), e.add(s))
        assert o.msa() == e.msa() and o.consensus() == e.consensus()
        o.close()
        e.close()
    # found by tests/tools/fuzz_emul.py: the co-optimal path runs through column 0 of a row whose other
    # cells are all irrelevant (leading graph nodes skipped) - column 0 counts for the next row's range
    for seqs, kw in ((['GA', 'AC', 'A', 'AC', 'A'], dict(dyn=1.0)),
                     (['CG', 'AA', 'CG', 'T', 'T', 'G'], dict(dyn=5.0)),
                     (['GA', 'C', 'A', 'CGGGAACGCTTATGAAAGA', 'GA', 'C', 'GA', 'C', 'C'], dict(dyn=5.0, dyn_ext=1))):
        o, e = oracle.PoaSession(1), EmuSession(ring_rows=3, **kw)
        for s in seqs:
            assert np.array_equal(o.add(s), e.add(s))
        o.close()
        e.close()
    w = synth.make_small_window(7, body_len=1500, sv_len=40, n_tumor=5, n_normal=5, n_carriers=3)
    e, o = EmuSession(ring_rows=12, dyn=4.4, dyn_ext=1), oracle.PoaSession(1)
    for s in w[0]:
        assert np.array_equal(o.add(s), e.add(s))
    from tests.emul.emul import lib
    assert e.kept_fraction() < lib().emu_static_fraction(e.h)    # fewer cells than the static bands
    o.close()
    e.close()


def test_margin_columns_equals_reference_loop(oracle):
    rng = np.random.default_rng(0)
    for _ in range(2000):
        n = int(rng.integers(1, 30))
        row = "".join(rng.choice(list("ACGT-"), n))
        nog = row.replace("-", "")
        k5, k3 = int(rng.integers(0, 6)), int(rng.integers(0, 6))
        f5 = nog[:k5] if rng.random() < 0.7 else "".join(rng.choice(list("ACGT"), k5))
        f3 = nog[len(nog) - k3:] if (rng.random() < 0.7 and k3 > 0) else "".join(rng.choice(list("ACGT"), k3))
        assert oracle.call_margin(row, f5, f3).tolist() == batch.margin_columns(row, f5, f3).tolist()


def test_encode_rejects_foreign_symbols():
    assert batch.encode_msa(["ATCG-", "atcg-"]).tolist() == [[0, 1, 2, 3, 4]] * 2
    with pytest.raises(KeyError):
        batch.encode_msa(["ACGN"])


def test_lpt_shards_cover_and_balance():
    costs = [float(c) for c in np.random.default_rng(1).integers(1, 100, 57)]
    for n in (1, 2, 4, 8):
        sh = shard.lpt_shards(costs, n)
        assert sorted(i for s in sh for i in s) == list(range(57))
        loads = [sum(costs[i] for i in s) for s in sh]
        assert max(loads) - min(loads) <= max(costs)
    assert shard.lpt_shards(costs, 4) == shard.lpt_shards(costs, 4)


def _gather_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    costs = [5.0, 1.0, 3.0, 2.0, 9.0, 4.0, 7.0]
    mine = shard.my_shard(costs, rank, world)
    recs = [["chr1", str(i), str(i + 1), rank] for i in mine]
    full = shard.gather_records(mine, recs, len(costs), group=shard.host_group())
    if rank == 0:
        q.put(full)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    full = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in full] == [str(i) for i in range(7)]
    assert {r[3] for r in full} == {0, 1}


def test_synth_formats(tmp_path):
    w = synth.make_small_window(3)
    assert len(w) == 5 and w[0][0].startswith(w[2]) and w[0][0].endswith(w[3])
    assert set("".join(w[0])) <= set("ACGT")
    p = str(tmp_path / "b.npz")
    synth.save_npz(p, [w, synth.make_small_window(4)])
    back = synth.load_npz(p)
    assert back[0][4] == w[4] and list(back[0][1]) == list(w[1]) and back[0][0] == w[0]
    a, b = synth.make_c2_window(5), synth.make_c2_window(5)
    assert a[0] == b[0] and a[4] == b[4]


def test_raw_bed_streaming_writer_resumes(tmp_path, monkeypatch):
    """localGraph_npz appends every batch to a part file; after an interruption --Continue
    computes only the missing windows and the merged Raw.bed is complete and sorted
    (reference --Continue: src/SVscope.py:195-200,213).  The GPU batch call is replaced by a
    stand-in: this is the host-side writer only."""
    import argparse
    import types
    from svscope_b200 import SVscope, synth
    wins = []
    for k in range(7):
        w = synth.make_small_window(30 + k, body_len=60, sv_len=20, n_tumor=3, n_normal=3, n_carriers=2)
        w[4] = "chr%d\t%d\t%d" % (1 + k % 2, 1000 * (7 - k), 1000 * (7 - k) + 60)
        wins.append(w)
    synth.save_npz(str(tmp_path / "T.vs.N.TandemRepeat.batch0.npz"), wins)
    calls = []

    def fake_batch(chunk, **kw):
        calls.append([w[4] for w in chunk])
        if fail["on"] == len(calls):
            raise RuntimeError("interrupted")
        return types.SimpleNamespace(records=[w[4].split("\t") + ["s", "i", 1, "g", "j", 2, "NormalOutput"] for w in chunk])

    fail = {"on": 3}
    monkeypatch.setattr(SVscope, "localgraph_batch", fake_batch)
    monkeypatch.setattr(SVscope, "BATCH_WINDOWS", 2)
    args = argparse.Namespace(savedir=str(tmp_path), TSampleID="T", NSampleID="N", Continue=False)
    with pytest.raises(RuntimeError):
        SVscope.localGraph_npz(args)
    out = tmp_path / SVscope.raw_bed_name("T", "N")
    part = tmp_path / (SVscope.raw_bed_name("T", "N") + ".part0")
    assert not out.exists() and len(part.read_text().splitlines()) == 4     # two batches reached the disk
    fail["on"] = -1
    calls.clear()
    args.Continue = True
    SVscope.localGraph_npz(args)
    assert sorted(sum(calls, [])) == sorted(w[4] for w in wins[4:])           # only the missing windows
    lines = out.read_text().splitlines()
    assert len(lines) == 7 and not part.exists()
    keys = [(x.split("\t")[0], int(x.split("\t")[1])) for x in lines]
    assert keys == sorted(keys) and all(len(x.split("\t")) == 10 for x in lines)
    # a fresh run (no --Continue) rewrites the file from scratch
    calls.clear()
    args.Continue = False
    SVscope.localGraph_npz(args)
    assert len(sum(calls, [])) == 7 and len(out.read_text().splitlines()) == 7


def _npz_worker(rank, world, port, savedir):
    import argparse
    import types
    import torch.distributed as dist
    from svscope_b200 import SVscope
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    SVscope.localgraph_batch = lambda chunk, **kw: types.SimpleNamespace(
        records=[w[4].split("\t") + ["s", "i", 1, "g", "rank%d" % rank, 2, "NormalOutput"] for w in chunk])
    SVscope.BATCH_WINDOWS = 2
    SVscope.localGraph_npz(argparse.Namespace(savedir=savedir, TSampleID="T", NSampleID="N", Continue=False))
    dist.barrier()
    dist.destroy_process_group()


def _npz_worker_env_only(rank, world, port, savedir):
    """What `torchrun -m svscope_b200.SVscope localGraph_npz` gives a process: environment variables, no group."""
    import argparse
    import types
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from svscope_b200 import SVscope
    SVscope.localgraph_batch = lambda chunk, **kw: types.SimpleNamespace(
        records=[w[4].split("\t") + ["s", "i", 1, "g", "rank%d" % rank, 2, "NormalOutput"] for w in chunk])
    SVscope.BATCH_WINDOWS = 2
    SVscope.main(["localGraph_npz", "-s", savedir, "-t", "T", "-n", "N"])


def test_cli_under_torchrun_env_creates_the_process_group(tmp_path):
    """The documented multi-GPU launch sets only RANK / WORLD_SIZE / MASTER_*: the CLI has to create
    the (gloo) process group itself, shard the windows and write ONE complete Raw.bed."""
    import torch.multiprocessing as mp
    from svscope_b200 import SVscope
    wins = []
    for k in range(9):
        w = synth.make_small_window(70 + k, body_len=40 + 5 * k, sv_len=10, n_tumor=3, n_normal=3, n_carriers=2)
        w[4] = "chr%d\t%d\t%d" % (1 + k % 3, 500 * (9 - k), 500 * (9 - k) + 40)
        wins.append(w)
    synth.save_npz(str(tmp_path / "a.npz"), wins[:5])
    synth.save_npz(str(tmp_path / "b.npz"), wins[5:])
    ctx = mp.get_context("spawn")
    port = 33500 + os.getpid() % 2000
    procs = [ctx.Process(target=_npz_worker_env_only, args=(r, 2, port, str(tmp_path))) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    lines = (tmp_path / SVscope.raw_bed_name("T", "N")).read_text().splitlines()
    assert len(lines) == 9 and len({x.split("\t")[1] for x in lines}) == 9
    assert {x.split("\t")[7] for x in lines} == {"rank0", "rank1"}
    assert not [f for f in os.listdir(tmp_path) if ".part" in f]


def test_two_rank_raw_bed_writer_gloo(tmp_path):
    """N>1 path of localGraph_npz on CPU: two ranks shard the windows, append to their own
    parts, rank 0 merges and sorts (the GPU batch call is a stand-in)."""
    import torch.multiprocessing as mp
    from svscope_b200 import SVscope
    wins = []
    for k in range(9):
        w = synth.make_small_window(50 + k, body_len=40 + 5 * k, sv_len=10, n_tumor=3, n_normal=3, n_carriers=2)
        w[4] = "chr%d\t%d\t%d" % (1 + k % 3, 500 * (9 - k), 500 * (9 - k) + 40)
        wins.append(w)
    synth.save_npz(str(tmp_path / "b.npz"), wins)
    ctx = mp.get_context("spawn")
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_npz_worker, args=(r, 2, port, str(tmp_path))) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    lines = (tmp_path / SVscope.raw_bed_name("T", "N")).read_text().splitlines()
    assert len(lines) == 9
    keys = [(x.split("\t")[0], int(x.split("\t")[1])) for x in lines]
    assert keys == sorted(keys)
    assert {x.split("\t")[7] for x in lines} == {"rank0", "rank1"}
    assert not [f for f in os.listdir(tmp_path) if ".part" in f]


def test_callsomaticsv_writes_somatic_bed(tmp_path, monkeypatch, oracle):
    """callsomaticSV = localGraph_npz + the MisScore head of AlnFeature (src/SVscope.py:282-286):
    <T>.Somatic.bed holds MisScorePipe's columns plus ABSMisScore.  GPU calls are stand-ins
    (batch records from the golden Raw.bed, alignments from the oracle): host glue only."""
    import argparse
    import json
    import types
    from svscope_b200 import SVscope, PairwiseCompare as PC
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "misscore_pipe.json")))
    rows = [ln.split("\t") for ln in gold["raw_bed"].splitlines()]
    wins = []
    for k, r in enumerate(rows):
        w = synth.make_small_window(80 + k, body_len=40, sv_len=10, n_tumor=3, n_normal=3, n_carriers=2)
        w[4] = "\t".join(r[:3])
        wins.append(w)
    synth.save_npz(str(tmp_path / "b.npz"), wins)
    by_key = {"\t".join(r[:3]): r for r in rows}
    monkeypatch.setattr(SVscope, "localgraph_batch",
                        lambda chunk, **kw: types.SimpleNamespace(records=[by_key[w[4]] for w in chunk]))

    def fake_pairs(pairs, *a, stats=None, **kw):
        out = np.zeros((len(pairs), 4), np.int32)
        for k, (x, y) in enumerate(pairs):
            r = oracle.pairwise_first_alignment(x, y)
            out[k] = (r["score"], r["length"], r["matches"], r["length"] - r["matches"])
        return out

    monkeypatch.setattr(PC, "misscore_pairs", fake_pairs)
    args = argparse.Namespace(savedir=str(tmp_path), TSampleID="T", NSampleID="N", Continue=False)
    raw = SVscope.callsomaticSV(args)
    assert sorted(open(raw).read().splitlines()) == sorted(gold["raw_bed"].splitlines())
    got = [ln.split("\t") for ln in (tmp_path / "T.Somatic.bed").read_text().splitlines()]
    want = {g[3]: g for g in gold["rows"]}
    assert len(got) == len(want)
    for g in got:
        w = want[g[3]]
        assert g[:3] == [str(w[0]), str(w[1]), str(w[2])] and g[4:6] == [w[4], w[5]]
        assert int(g[6]) == w[6] and g[7] == w[7] and int(g[8]) == abs(w[6])


def test_tdscope_rescue_control_flow_matches_reference(oracle):
    """``SomTDDetector.TDscope`` (extraction callables + decision + DUP rescue, reference
    src/SomTDDetector.py:26-61) against records the reference's own TDscope produced for the same
    scenarios (oracle/gen_golden_tdscope.py); the decision callable is the oracle on both sides, so
    this pins the control flow: which extraction wins, which flag is written."""
    import json
    from oracle.gen_golden_tdscope import makers, scenarios
    from svscope_b200.SomTDDetector import TDscope
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "tdscope_cases.json")))["records"]
    flags = set()
    for name in scenarios():
        TDRecord, dm, dm2 = makers(name)
        np.random.seed(2023)
        rec = TDscope(TDRecord, dm, dm2, oracle.decision)
        assert [str(x) for x in rec] == gold[name], name
        flags.add(rec[-1])
    assert flags == {"NormalOutput|EMOutput", "NormalOutput", "UnspanedSV|EMOutput", "UnspannedSV|EMOutput",
                     "UnspanedSV", "UnspannedSV"}


def _fake_em_batch(oracle):
    """A numpy stand-in for svs_em_batch that follows the device contract of
    include/svscope_b200.h (schedule [M from labels] E, n_steps x (M, E); status = index of the
    M-step that met pi*N < 1 or NaN, nothing drawn on the 'device')."""
    def m_nodraw(gamma, X):
        N, nf = X.shape
        pi = gamma.sum(axis=0) / N
        if (pi * N < 1).any() or np.isnan(pi).any():
            return None, None
        tot = np.dot(gamma.T, np.ones((N, nf), dtype=np.int64))
        theta = np.dstack([np.dot(gamma.T, np.where(X == a, 1, 0)) / tot for a in range(5)])
        return pi, theta

    def run(ctx, Xs, tasks, want_theta=False):
        out = []
        for t in tasks:
            X = Xs[t.x_index]
            N, K = X.shape[0], t.K
            status, pi, theta, gamma = -1, t.pi, t.theta, None
            if t.labels is not None:
                g0 = np.zeros((N, K))
                g0[np.arange(N), np.asarray(t.labels)] = 1
                pi, theta = m_nodraw(g0, X)
                if pi is None:
                    status = 0
            if status < 0:
                gamma, _ = oracle.e_step(K, np.asarray(pi, float), np.asarray(theta, float), X)
                for it in range(1, t.n_steps + 1):
                    p2, t2 = m_nodraw(gamma, X)
                    if p2 is None:
                        status = it
                        break
                    pi, theta = p2, t2
                    gamma, _ = oracle.e_step(K, pi, theta, X)
            if status >= 0:
                out.append(dict(gamma=np.zeros((N, K)), pi=np.zeros(K), lik=np.zeros(N), theta=None, status=status))
            else:
                out.append(dict(gamma=gamma, pi=np.asarray(pi, float), status=-1,
                                lik=oracle.per_read_loglik(np.asarray(pi, float), np.asarray(theta, float), gamma, X),
                                theta=np.asarray(theta, float) if want_theta else None))
        return out
    return run


def test_em_host_driver_replays_the_reference_rng_order(oracle, monkeypatch):
    """batch.em_cluster_many (Ward start per K, device fits, host-side Dirichlet re-draws from the
    global RNG in the reference's order, NaN-BIC retries, K choice) with a numpy stand-in for the
    device call: equal to oracle.em_cluster - itself fuzzed against the reference's EMCluster - on
    random matrices, most of them small enough to hit the re-draw path."""
    import warnings
    monkeypatch.setattr(batch, "em_batch", _fake_em_batch(oracle))
    rng = np.random.default_rng(5)
    redraw = 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for it in range(150):
            N, nf = int(rng.integers(3, 22)), int(rng.integers(10, 60))
            X = np.tile(rng.integers(0, 4, nf), (N, 1))
            for _ in range(int(rng.integers(0, 3))):
                rows = rng.choice(N, size=int(rng.integers(1, max(2, N // 2))), replace=False)
                cols = rng.choice(nf, size=int(rng.integers(1, nf)), replace=False)
                X[np.ix_(rows, cols)] = rng.integers(0, 5)
            noise = rng.random((N, nf)) < float(rng.choice([0.0, 0.05, 0.3]))
            X[noise] = rng.integers(0, 5, int(noise.sum()))
            X = X.astype(np.int64)
            max_C = 9 if it % 3 else int(rng.integers(2, 9))       # EMCluster's max_C (the reference passes the default)
            want, info = oracle.em_cluster(X, max_C=max_C, reseed=True, return_info=True)
            got = batch.em_cluster_many(None, [X], [oracle.pairwise_identity(X)], [oracle.zero_param_num(X)],
                                        want_theta=True, reseed=True, max_C=max_C)[0]
            assert got["K"] == want[0] and np.array_equal(got["labels"], want[2]), it
            assert np.allclose(got["bics"], want[6], rtol=1e-9, equal_nan=True)
            assert np.allclose(got["gamma"], want[4], rtol=1e-9, atol=1e-300) and np.allclose(got["pi"], want[5], rtol=1e-9)
            assert np.allclose(got["theta"], want[3], rtol=1e-9, atol=1e-300)
            assert (got["n_redraws"] > 0) == (info["n_fallback"] > 0)
            redraw += got["n_redraws"] > 0
    assert redraw > 30


def _device_stand_ins(oracle, monkeypatch):
    """Replace every device wrapper used by batch.localgraph_batch with an oracle/numpy stand-in
    that follows the same contract (the GPU suite tests the real ones against the oracle)."""
    import types
    from svscope_b200 import poa_api

    def fake_upload(ctx, windows):
        seqs = [s for w in windows for s in w[0]] + [""]
        off = np.zeros(len(seqs) + 1, np.int64)
        off[1:] = np.cumsum([len(s) for s in seqs])
        return types.SimpleNamespace(seqs=seqs, off=off, nbytes=int(off[-1]))

    def fake_poa_groups(ctx, reads, groups, algorithm=1, want_msa=True, scores=None, as_array=False, strict=True):
        cons, msas = [], []
        for g in groups:
            c, m = oracle.poa([reads.seqs[i] for i in g], algorithm) if len(g) else ("", [])
            cons.append(c)
            if want_msa and as_array:
                msas.append(np.frombuffer("".join(m).encode(), np.uint8).reshape(len(m), -1) if m else np.zeros((0, 0), np.uint8))
            else:
                msas.append(m if want_msa else [])
        st = {k: 0.0 for k in poa_api.STAT_NAMES}
        st["status"] = np.zeros(len(groups), np.int32)
        return cons, msas, st

    def fake_msa_features(ctx, encs, drops, cutoffs):
        out = []
        for e, d, cut in zip(encs, drops, cutoffs):
            e = np.asarray(e).astype(np.int64)
            cols = np.flatnonzero(np.asarray(d) == 0)
            sub = e[:, cols]
            keep = np.zeros(e.shape[1], bool)
            keep[cols[oracle.find_non_same_site(sub, cutoff=cut)]] = True
            X = e[:, keep]
            ident = (X[:, None, :] == X[None, :, :]).sum(axis=2).astype(np.int32)
            out.append((keep, int(keep.sum()), oracle.zero_param_num(X) if X.size else 0, ident))
        return out

    class FakeJob:
        """stand-in for poa_api.PoaJob (svs_poa_submit / svs_poa_wait)"""
        def __init__(self, ctx, reads, groups, algorithm=1, want_msa=True, scores=None):
            self.args = (ctx, reads, groups, algorithm, want_msa, scores)

        def result(self, as_array=False, strict=True):
            ctx, reads, groups, algorithm, want_msa, scores = self.args
            return fake_poa_groups(ctx, reads, groups, algorithm, want_msa, scores, as_array, strict)

        def close(self):
            pass

    monkeypatch.setattr(batch, "upload_windows", fake_upload)
    monkeypatch.setattr(batch, "poa_groups", fake_poa_groups)
    monkeypatch.setattr(poa_api, "PoaJob", FakeJob)
    monkeypatch.setattr(batch, "msa_features", fake_msa_features)
    monkeypatch.setattr(batch, "em_batch", _fake_em_batch(oracle))


def test_batch_host_orchestration_equals_oracle_with_device_stand_ins(oracle, monkeypatch):
    """Everything batch.localgraph_batch does on the host - gates, the empty-read quirk, flank
    columns, feature cut-off, EM driver with per-window reseeding, cluster labelling, consensus
    grouping ('-' for empty clusters), record assembly, batching of several windows - against
    oracle.decision (itself fuzzed against the reference's Decision) on random windows with
    empty reads, a third tag and non-default cut-offs."""
    import warnings
    _device_stand_ins(oracle, monkeypatch)
    rng = np.random.default_rng(17)
    n_em = 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for b in range(10):
            wins = []
            for _ in range(6):
                nt, nn = int(rng.integers(2, 10)), int(rng.integers(2, 10))
                w = synth.make_sv_window(int(rng.integers(1 << 30)), int(rng.integers(120, 320)),
                                         "DEL" if rng.random() < 0.5 else "INS", int(rng.integers(10, 90)), nt, nn,
                                         int(rng.integers(0, nt + 1)), float(rng.choice([0.0, 0.03, 0.1])))
                seqs, ids = list(w[0]), np.array(w[1])
                r = rng.random()
                if r < 0.2:
                    for k in rng.choice(np.arange(1, len(seqs)), size=int(rng.integers(1, 3)), replace=False):
                        seqs[int(k)] = ""
                elif r < 0.3:
                    ids = np.array([x.replace("_normal|", "_other|") if rng.random() < 0.4 else x for x in ids])
                wins.append([seqs, ids, w[2], w[3], w[4]])
            kw = dict(readcutoff=int(rng.integers(2, 5)), hcutoff=int(rng.integers(2, 5)),
                      scutoff=float(rng.choice([0.05, 0.2]))) if b % 3 == 2 else {}
            got = batch.localgraph_batch(wins, ctx=object(), **kw).records
            for w, g in zip(wins, got):
                want = oracle.decision(w[4], w[0], w[1], w[2], w[3], **kw)
                assert [str(x) for x in want] == [str(x) for x in g]
                n_em += str(g[-1]).endswith("EMOutput")
    assert n_em >= 10


def test_window_beyond_the_mixture_limit_fails_alone(oracle, monkeypatch):
    """A window with more rows than the mixture kernel takes gets a flagged record; its neighbours in
    the batch still equal the oracle (one outlier must not fail the batch)."""
    import warnings
    _device_stand_ins(oracle, monkeypatch)
    monkeypatch.setattr(batch, "EM_MAX_READS", 12)
    small = synth.make_sv_window(5, 200, "DEL", 40, 5, 5, 3, 0.03)
    big = synth.make_sv_window(6, 200, "INS", 40, 9, 9, 4, 0.03)          # 18 reads > 12
    other = synth.make_sv_window(7, 240, "DEL", 50, 6, 5, 3, 0.03)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = batch.localgraph_batch([small, big, other], ctx=object())
        got = out.records
        for w, g in ((small, got[0]), (other, got[2])):
            want = oracle.decision(w[4], w[0], w[1], w[2], w[3])
            assert [str(x) for x in want] == [str(x) for x in g]
    assert got[1][9].endswith("|MixtureLimit18") and got[1][3] == "-" and got[1][5] == 0
    assert out.stats["poa_failed_windows"] == 1


def test_cell_header_table_keys_and_codes_selfcheck():
    """poa_cell.h: the 32-entry predecessor table equals the definitions it tabulates for every supported
    gap-parameter set, the shuffle-table key fold equals a direct 'first in-edge attaining the maximum' on
    random in-edge sets with frequent ties, and every code field survives make_code -> decoder."""
    from tests.emul.emul import cell_selfcheck
    bad, n = cell_selfcheck(seed=7, n_random=300)
    assert n > 100000 and bad == 0


def test_bench_line_contract_with_device_stand_ins():
    """bench.py's control flow and JSON line with stand-ins for the device stages (tests/tools/fake_bench.py):
    the keys the driver reads, the per-step batch shrinking to the time budget (with the second calibration
    step), and `config` being the identical dict in the ours arm and in --impl reference."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    fake = os.path.join(root, "tests", "tools", "fake_bench.py")
    out = subprocess.run([sys.executable, fake, "0.005", "--windows", "200", "--steps", "6", "--warmup", "3", "--budget-s", "3",
                          "--total-s", "0"],
                         capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert key in line, key
    assert line["steps"] == 6 and line["warmup"] == 3 and line["n_gpus"] == 1 and line["scaling"] == "weak"
    assert set(line["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert set(line["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    assert "model" not in line["config"] and line["config"]["workload"].startswith("configs[1]")
    # 6 steps of the full batch (1 s each with the stand-ins) do not fit 3 s: the step shrinks, and says so
    assert 64 <= line["run"]["windows_per_step"] < 200
    assert line["calibration"]["second_calibration_windows"] >= line["run"]["windows_per_step"]
    assert line["ms_per_step"] * 6 / 1e3 < 3.0 * 1.35
    ref = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--cpu-budget", "0.5",
                          "--windows", "200", "--steps", "6", "--warmup", "3"], capture_output=True, text=True, timeout=600, cwd=root)
    assert ref.returncode == 0, ref.stderr[-2000:]
    rline = json.loads(ref.stdout.strip().splitlines()[-1])
    assert rline["impl"] == "reference" and rline["config"] == line["config"]
    assert rline["metric"] == line["metric"] and rline["unit"] == line["unit"]
    assert rline["e2e"]["h2d_bytes_per_step"] == 0 and rline["cpu_baseline"]["kind"] == "port"
