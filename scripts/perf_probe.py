"""Performance probe at realistic sizes (not a bench): stage timings and DP throughput."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from svscope_b200 import synth, _lib
from svscope_b200.batch import localgraph_batch, upload_windows

nwin = int(os.environ.get("NWIN", 32))
ctx = _lib.Context(0)
for k in ("poa_threads", "ring_rows", "arena_mb", "poa_cols", "dp_kernel", "prune"):
    if k.upper() in os.environ:
        ctx.set_option(k, int(os.environ[k.upper()]))
t0 = time.time()
if os.environ.get("C1"):
    depth = int(os.environ.get("DEPTH", 30))
    wins = [synth.make_c1(seed=s, n_tumor=depth, n_normal=depth, n_carriers=depth // 2) for s in range(1, nwin + 1)]
else:
    wins = synth.make_c2(nwin, depth=int(os.environ.get("DEPTH", 30)))
print("gen", round(time.time() - t0, 1), "s; windows", len(wins), "mean len", np.mean([len(w[0][0]) for w in wins]), flush=True)
reads = upload_windows(ctx, wins)
for rep in range(int(os.environ.get("REPS", 1))):
    t0 = time.time()
    out = localgraph_batch(wins, ctx=ctx, reads=reads, edit_distance=bool(int(os.environ.get("ED", "1"))))
    dt = time.time() - t0
    st = out.stats
    print("rep", rep, "total %.2fs -> %.2f windows/s" % (dt, len(wins) / dt), flush=True)
    print(" timings", {k: round(v, 2) for k, v in out.timings.items()}, flush=True)
    print(" poa cells %.3e aligns %d dp_ms(sum) %.0f tb_ms(sum) %.0f launches %d h2d %.1fMB d2h %.1fMB exported_rows %.1f%%" % (
        st["poa_cells"], st["poa_alignments"], st["poa_dp_ms"], st["poa_tb_ms"], st["poa_dp_launches"],
        st["poa_h2d_bytes"] / 1e6, st["poa_d2h_bytes"] / 1e6, 100 * st["poa_exported_rows"] / max(1, st["poa_rows"])), flush=True)
    cyc = {k: st["poa_cyc_" + k] for k in ("export", "dp", "traceback", "merge", "rank", "finish")}
    tot = max(1.0, sum(cyc.values()))
    print(" window-kernel phase shares (thread-0 cycles): " + ", ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in cyc.items()),
          "| retries %d failed %d" % (st["poa_prune_retries"], st["poa_failed_windows"]), flush=True)
    poa_t = out.timings["poa_msa"] + out.timings["poa_consensus"]
    print(" POA wall GCUPS %.1f ; ED cells %.3e in %.1f ms -> %.0f GCUPS" % (
        st["poa_cells"] / poa_t / 1e9, st["ed_cells"], st["ed_ms"], st["ed_cells"] / max(st["ed_ms"], 1e-9) / 1e6), flush=True)
    print(" EMOutput", sum(r[-1].endswith("EMOutput") for r in out.records), "of", len(wins), "redraw windows", st["em_redraw_windows"], flush=True)
