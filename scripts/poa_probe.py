"""POA-only probe: one batch of configs[1]-like windows through svs_poa_batch for several kernel
shapes; prints wall time, nominal GCUPS and the per-phase thread-0 cycles per alignment."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from svscope_b200 import synth, _lib
from svscope_b200.poa_api import poa_groups

nwin = int(os.environ.get("NWIN", 8))
body = os.environ.get("BODY")
ctx = _lib.Context(0)
if os.environ.get("ARENA_MB"):
    ctx.set_option("arena_mb", int(os.environ["ARENA_MB"]))
if body:
    depth = int(os.environ.get("DEPTH", 30))
    wins = [synth.make_sv_window(100 + i, int(body), "DEL" if i % 2 else "INS", 300, depth, depth, max(3, depth // 3), 0.05) for i in range(nwin)]
else:
    wins = synth.make_c2(nwin)
seqs, groups = [], []
for w in wins:
    groups.append(list(range(len(seqs), len(seqs) + len(w[0]))))
    seqs += w[0]
reads = _lib.ReadSet(ctx, seqs)
ref = None
for cfg in os.environ.get("CFGS", "128,10,2;256,10,2;128,10,1").split(";"):
    t, ring, dp = [int(v) for v in cfg.split(",")]
    ctx.set_option("poa_threads", t); ctx.set_option("ring_rows", ring); ctx.set_option("dp_kernel", dp)
    for rep in range(int(os.environ.get("REPS", 1))):
        t0 = time.time()
        cons, msas, st = poa_groups(ctx, reads, groups, want_msa=True, as_array=True, strict=False)
        dt = time.time() - t0
        na = max(1.0, st["alignments"])
        ph = {k: st["cyc_" + k] / na / 1.9e6 for k in ("export", "dp", "traceback", "merge", "rank", "finish")}
        print("cfg T=%d ring=%d dp=%d: %.2fs wall, kernel %.0f ms, %d launches, %.1f nominal GCUPS, evaluated %.1f%%, failed %d, retries %d | ms/alignment (thread 0): %s"
              % (t, ring, dp, dt, st["dp_ms"], st["dp_launches"], st["cells"] / dt / 1e9, 100 * st["eval_cells"] / max(1.0, st["cells"]), st["failed_groups"], st["prune_retries"],
                 " ".join("%s %.2f" % kv for kv in ph.items())), flush=True)
        wl = max(1.0, st["wcyc_loop"] + st["wcyc_wait_end"])
        print("    warps in the DP: polling left %.1f%%, polling right/boundary %.1f%%, waiting at the end %.1f%%, working %.1f%%"
              % (100 * st["wcyc_wait_left"] / wl, 100 * st["wcyc_wait_right"] / wl, 100 * st["wcyc_wait_end"] / wl,
                 100 * (st["wcyc_loop"] - st["wcyc_wait_left"] - st["wcyc_wait_right"]) / wl), flush=True)
        if ref is None:
            ref = (cons, [m.tobytes() for m in msas])
        else:
            assert ref[0] == cons and ref[1] == [m.tobytes() for m in msas], "outputs differ between kernel shapes"
print("all shapes agree")
