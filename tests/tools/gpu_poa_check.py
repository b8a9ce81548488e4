"""First GPU bring-up: alignment pairs / MSA / consensus of the CUDA path vs the oracle."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from oracle import oracle as O
from svscope_b200 import synth, _lib
from svscope_b200.poa_api import align_pairs, poa_groups

ctx = _lib.Context(0)
for k, v in [("poa_threads", int(os.environ.get("POA_T", 256))), ("ring_rows", int(os.environ.get("RING", 12))),
             ("workers", int(os.environ.get("WORKERS", 4)))]:
    ctx.set_option(k, v)
rng = np.random.default_rng(0)
bad = 0
t0 = time.time()
for it in range(int(os.environ.get("NCASE", 60))):
    L = int(rng.integers(1, 400)); n = int(rng.integers(2, 10)); err = float(rng.choice([0.0, 0.05, 0.15]))
    base = synth._rand_seq(rng, L)
    if it % 3 == 0:
        mot = synth._rand_seq(rng, int(rng.integers(1, 6))); base = np.tile(mot, max(1, L // len(mot)))
    seqs = []
    for r in range(n):
        s = base.copy()
        for _ in range(int(rng.integers(0, 3))):
            p = int(rng.integers(0, len(s) + 1)); ln = int(rng.integers(1, 40))
            s = np.concatenate([s[:p], s[p + ln:]]) if rng.random() < 0.5 else np.concatenate([s[:p], synth._rand_seq(rng, ln), s[p:]])
        if len(s) == 0: s = synth._rand_seq(rng, 3)
        seqs.append(synth._to_str(synth.noisy_copy(rng, s, err)))
    if it % 7 == 0: seqs[int(rng.integers(1, n))] = ""
    o = O.PoaSession(1)
    ref = [o.add(s) for s in seqs]
    got = align_pairs(ctx, seqs)
    for k in range(n):
        if ref[k].shape != got[k].shape or not np.array_equal(ref[k], got[k]):
            bad += 1
            print("MISMATCH case", it, "seq", k, ref[k].shape, got[k].shape, flush=True)
            nn = min(len(ref[k]), len(got[k]))
            d = np.where((ref[k][:nn] != got[k][:nn]).any(axis=1))[0]
            if len(d): print(" first diff at", d[0], ref[k][max(0, d[0]-2):d[0]+3].tolist(), got[k][max(0, d[0]-2):d[0]+3].tolist())
            break
    o.close()
print("small cases bad =", bad, "time", round(time.time() - t0, 1), flush=True)

# windows through the batch API (MSA + consensus), several groups at once, multi-pass reads
wins = [synth.make_small_window(s, body_len=int(b), sv_len=int(b) // 4, n_tumor=6, n_normal=6, n_carriers=3)
        for s, b in [(1, 300), (2, 1200), (3, 2500), (4, 5000)]]
wins.append(synth.make_c3(seed=3, total_len=1500, n_tumor=6, n_normal=6, n_carriers=3))
allseqs, groups = [], []
for w in wins:
    groups.append(list(range(len(allseqs), len(allseqs) + len(w[0]))))
    allseqs += w[0]
reads = _lib.ReadSet(ctx, allseqs)
t0 = time.time()
cons, msas, st = poa_groups(ctx, reads, groups)
dt = time.time() - t0
print("batch stats", {k: round(v, 2) for k, v in st.items()}, "wall", round(dt, 2), flush=True)
for w, c, m in zip(wins, cons, msas):
    t1 = time.time(); oc, om = O.poa(w[0], 1); ot = time.time() - t1
    print("window L~%d: consensus %s msa %s (oracle %.1fs, %.3f GCUPS)" % (len(w[0][0]), c == oc, m == om, ot, O.poa.last_cells / ot / 1e9), flush=True)
    bad += (c != oc) + (m != om)
print("GCUPS device (dp kernel, summed launches):", st["cells"] / (st["dp_ms"] * 1e-3) / 1e9, flush=True)
print("RESULT", "PASS" if bad == 0 else "FAIL")
sys.exit(0 if bad == 0 else 1)
