"""BASELINE.json configs[4], scaled to a short run: isolated kernel sweep.
POA: per read length a small batch of identical-size windows (reference + 10 noisy reads);
edit distance: 256 pairs per length at 10 % divergence.  GCUPS on nominal cells, next to the
CPU oracle on the sizes it finishes quickly.  Writes gpurun_out/kernel_sweep.json."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from oracle import oracle as O
from svscope_b200 import synth, _lib
from svscope_b200.batch import edit_distance_matrices
from svscope_b200.poa_api import poa_groups

ctx = _lib.Context.default(0)
ctx.set_option("workers", 8)
out = {"poa": [], "edit_distance": []}
rng = np.random.default_rng(1)
for kb, nwin in [(1, 64), (2, 64), (5, 32), (10, 32), (20, 16)]:
    L = kb * 1000
    wins = []
    for w in range(nwin):
        base = synth._rand_seq(rng, L)
        wins.append([synth._to_str(base)] + [synth._to_str(synth.noisy_copy(rng, base, 0.05)) for _ in range(10)])
    seqs, groups = [], []
    for w in wins:
        groups.append(list(range(len(seqs), len(seqs) + len(w))))
        seqs += w
    reads = _lib.ReadSet(ctx, seqs)
    poa_groups(ctx, reads, groups[:2], want_msa=False)          # warm-up
    t0 = time.perf_counter()
    cons, _, st = poa_groups(ctx, reads, groups, want_msa=False)
    dt = time.perf_counter() - t0
    row = dict(read_kb=kb, windows=nwin, reads_per_window=10, nominal_cells=st["cells"], seconds=dt,
               gcups=st["cells"] / dt / 1e9, graph_rows_mean=st["rows"] / max(1, st["alignments"]))
    if kb <= 2:
        t0 = time.perf_counter()
        oc, _ = O.poa(wins[0], 1)
        ct = time.perf_counter() - t0
        row.update(cpu_gcups_one_core=O.poa.last_cells / ct / 1e9, cpu_equal=(oc == cons[0]))
    out["poa"].append(row)
    print(row, flush=True)
    reads.close()
for kb in [1, 2, 5, 10, 20, 50, 100]:
    L = kb * 1000
    npairs = 256
    seqs = []
    for p in range(npairs):
        a = synth._rand_seq(rng, L)
        seqs += [synth._to_str(a), synth._to_str(synth.noisy_copy(rng, a, 0.10))]
    reads = _lib.ReadSet(ctx, seqs)
    groups = [[2 * p, 2 * p + 1] for p in range(npairs)]
    edit_distance_matrices(ctx, reads, groups[:4])
    mats, st = edit_distance_matrices(ctx, reads, groups)
    row = dict(read_kb=kb, pairs=npairs, cells=st["cells"], kernel_ms=st["ms"], gcups=st["cells"] / st["ms"] / 1e6)
    if kb <= 10:
        t0 = time.perf_counter()
        d = O.levenshtein(seqs[0], seqs[1], bitparallel=True)
        ct = time.perf_counter() - t0
        row.update(cpu_gcups_one_core=len(seqs[0]) * len(seqs[1]) / ct / 1e9, cpu_equal=(int(mats[0][0, 1]) == d))
    out["edit_distance"].append(row)
    print(row, flush=True)
    reads.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/kernel_sweep.json", "w"), indent=1)
