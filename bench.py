#!/usr/bin/env python
"""bench.py — localGraph windows/s on synthetic tumor/normal long-read windows.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--windows 1000] [--impl ours|reference]

One "step" = one pass of the localGraph hot path (window MSA by partial-order alignment,
feature selection, mixture-model clustering with BIC selection, per-cluster consensus POA,
read-by-read edit-distance matrix, 10-field records) over one batch of windows drawn from
BASELINE.json configs[1]: INS/DEL windows at 30x depth (30 tumor + 30 normal reads), 5-15 kb.
N>1 (torchrun, one rank per GPU): every rank gets its own batch of the same distribution
(weak scaling), no collective on the data path; `value` = windows of all ranks / max time.

`value`   device-resident: the reads are uploaded to HBM before the timed region.
`e2e`     the same step through the public batch call with host buffers (read upload and all
          result copies inside the timed region).
Batch size of a step: the configs[1] batch (1000 windows) unless K steps of it would not fit
`--budget-s` seconds / the whole process would not fit `--total-s` (the driver allows 870 s
per run); then every step takes the first n windows of the batch, n chosen from one untimed
full-size step and one untimed step of n windows, and the JSON line says so
(`run.windows_per_step`).  `config` holds the workload only and is the same dict in both arms;
what is specific to this implementation and run is under `run`.  Warm-up steps run on a 64-window slice (context, arena, kernels).
`--impl reference`  times the CPU path (oracle port of the reference: pyspoa is not
          installable offline) on a bounded sample with all host cores, once.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
_T_IMPORT = time.time()
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # before any CUDA context (see svscope_b200/__init__.py)
if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":      # NCCL would print its version on stdout, next to the JSON line
    os.environ["NCCL_DEBUG"] = "WARN"

METRIC = "localGraph windows/sec"
UNIT = "windows/s"
WORKLOAD = "configs[1]: synthetic INS/DEL windows, 30 tumor + 30 normal reads, 5-15 kb, 5% error"
CONFIGS1_BATCH = 1000
# dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of the window kernel
# (profiles/r02_poa_window_kernel_128x8_2cta_ncu_full.txt: 2 windows x 61 reads of 3 kb, 120 alignments)
NCU_TRAFFIC = {"bytes": 1.478912e6 + 1.219072e6, "alignments": 120.0}


# ------------------------------------------------------------------------------------------
def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.device), "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.strip().split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                pw.append(float(r[3]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        busy = [v for v in sm if v > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_median": statistics.median(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}


def dist_setup(n_gpus):
    import torch
    rank = env_int("RANK", 0)
    world = env_int("WORLD_SIZE", 1)
    local = env_int("LOCAL_RANK", 0)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        backend = "nccl" if torch.cuda.is_available() else "gloo"
        if torch.cuda.is_available():
            torch.cuda.set_device(local)
        dist.init_process_group(backend)
    return rank, world, local


def barrier_sync():
    import torch
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def _reduce(x: float, op_name: str) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return x
    t = torch.tensor([x], dtype=torch.float64, device="cuda" if torch.cuda.is_available() else "cpu")
    dist.all_reduce(t, op=getattr(dist.ReduceOp, op_name))
    return float(t.item())


def max_over_ranks(x):
    return _reduce(x, "MAX")


def min_over_ranks(x):
    return _reduce(x, "MIN")


def sum_over_ranks(x):
    return _reduce(x, "SUM")


def process_age_s():
    """Seconds since this process started (interpreter start-up and the first `import torch` included)."""
    try:
        import psutil
        return max(0.0, time.time() - psutil.Process().create_time())
    except Exception:
        return time.time() - _T_IMPORT


def usable_cores(gb_per_process=10.0):
    """Host threads the CPU arm uses: the cores this process may run on, capped so that one oracle process per
    core (five int32 matrices of one alignment: up to ~8 GB for a configs[1] window) fits the free memory."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    try:
        avail_kb = next(int(l.split()[1]) for l in open("/proc/meminfo") if l.startswith("MemAvailable"))
        n = min(n, max(1, int(avail_kb / 1048576.0 / gb_per_process)))
    except Exception:
        pass
    return max(1, n)


def workload_config(n_windows, edit_distance, world):
    """`config` of BOTH arms (ours and --impl reference print the same dict): the workload only."""
    return {"workload": WORKLOAD, "windows_per_gpu": n_windows, "configs1_batch": CONFIGS1_BATCH,
            "reads_per_window": 60, "edit_distance_matrix": bool(edit_distance),
            "parallelism": f"dp{world}: every rank its own batch of windows, no collective on the data path",
            "l2": "GPU arm: inputs larger than L2 (reads + traceback codes >> 126 MB per step); CPU arm: not applicable"}


def make_batch(n_windows, rank):
    from svscope_b200 import synth
    return [synth.make_c2_window(rank * n_windows + i) for i in range(n_windows)]


# ------------------------------------------------------------------------------------------
# CPU path (oracle port of the reference) on a bounded, cost-stratified sample
# ------------------------------------------------------------------------------------------
def model_cells(window, new_node_rate=0.035):
    """Nominal DP cells of the window MSA, sum over reads of (|V| + 1)(L + 1), with the graph
    growing by `new_node_rate` nodes per aligned base (substitutions + insertions of the 5 %
    error model).  Used to place windows in cost strata and, in the reference arm, to turn
    the measured CPU cell rate into windows/s; the ours arm prints the residual of this model
    against the cells the device counted."""
    seqs = window[0]
    V = float(len(seqs[0]))
    cells = 0.0
    for s in seqs[1:]:
        cells += (V + 1.0) * (len(s) + 1.0)
        V += new_node_rate * len(s)
    return cells


def _cpu_worker(args):
    """One window on one core until the deadline: window MSA read by read (cells counted by the
    oracle engine), then a slice of the read-by-read Levenshtein matrix."""
    seqs, budget_s, ed_budget_s = args
    from oracle import oracle as O
    t0 = time.perf_counter()
    sess = O.PoaSession(1)
    cells, n_done = 0.0, 0
    for s in seqs:
        sess.add(s)
        cells += float(O.lib().spo_last_cells(sess._h))
        n_done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    t_poa = time.perf_counter() - t0
    sess.close()
    t1 = time.perf_counter()
    ed_cells, k = 0.0, 0
    reads = seqs[1:]
    while time.perf_counter() - t1 < ed_budget_s and k + 1 < len(reads):
        O.levenshtein(reads[k], reads[k + 1], bitparallel=True)
        ed_cells += float(len(reads[k])) * float(len(reads[k + 1]))
        k += 1
    t_ed = time.perf_counter() - t1
    return dict(cells=cells, t_poa=t_poa, alignments=n_done, complete=n_done == len(seqs), ed_cells=ed_cells, t_ed=t_ed)


def cpu_rates(windows, budget_s=20.0, cores=None):
    """Scalar-port rates on `cores` windows spread evenly over the cost-sorted batch, one process each."""
    import multiprocessing as mp
    cores = min(cores or usable_cores(), len(windows))
    costs = np.array([model_cells(w) for w in windows])
    order = np.argsort(costs)
    picks = [int(order[int((k + 0.5) * len(order) / cores)]) for k in range(cores)]
    jobs = [(list(windows[i][0]), budget_s, max(2.0, 0.15 * budget_s)) for i in picks]
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(_cpu_worker, jobs, chunksize=1)
    wall = time.perf_counter() - t0
    poa_rate = sum(r["cells"] for r in res) / max(1e-9, sum(r["t_poa"] for r in res))      # cells / core-second
    ed_rate = sum(r["ed_cells"] for r in res) / max(1e-9, sum(r["t_ed"] for r in res))
    return dict(cores=cores, poa_cells_per_core_s=poa_rate, ed_cells_per_core_s=ed_rate, wall_s=wall,
                alignments=sum(r["alignments"] for r in res), windows_completed=sum(r["complete"] for r in res),
                sample_cost_quantiles=[float(costs[i] / costs.mean()) for i in picks])


def cpu_baseline(windows, poa_cells_per_window, ed_cells_per_window, budget_s, with_ed=True):
    r = cpu_rates(windows, budget_s)
    t_window = poa_cells_per_window / r["poa_cells_per_core_s"]
    t_with_ed = t_window + ed_cells_per_window / max(1.0, r["ed_cells_per_core_s"])
    value_no_ed = r["cores"] / t_window
    value_ed = r["cores"] / t_with_ed
    sample = (f"{r['cores']} windows spread evenly over the cost-sorted batch (cost {min(r['sample_cost_quantiles']):.2f}x-"
              f"{max(r['sample_cost_quantiles']):.2f}x of the mean), one process per window for {budget_s:.0f} s each on {r['cores']} cores: "
              f"{r['alignments']} alignments at {r['poa_cells_per_core_s'] / 1e9:.3f} GCUPS per core (scalar five-matrix engine), "
              f"bit-parallel Levenshtein at {r['ed_cells_per_core_s'] / 1e9:.2f} GCUPS per core; windows/s = cores / "
              f"(DP cells per window / rate + edit-distance cells per window / rate); features and mixture model (~2 % of the "
              f"CPU time of a window) not charged")
    return dict(value=value_ed if with_ed else value_no_ed, unit=UNIT, cores=r["cores"], kind="port", sample=sample,
                value_without_edit_distance=value_no_ed, value_with_edit_distance=value_ed,
                value_reference_pool6=min(6, r["cores"]) / (t_with_ed if with_ed else t_window),
                value_one_core=1.0 / (t_with_ed if with_ed else t_window),
                poa_gcups_per_core=r["poa_cells_per_core_s"] / 1e9, ed_gcups_per_core=r["ed_cells_per_core_s"] / 1e9,
                poa_cells_per_window=poa_cells_per_window, ed_cells_per_window=ed_cells_per_window, wall_s=r["wall_s"],
                note="CPU restatement of pyspoa (scalar; the real pyspoa 0.2.1 SIMD engine cannot be installed offline and would "
                     "be several times faster) + bit-parallel Levenshtein; the reference itself runs no Levenshtein "
                     "(src/DecisionMaker.py:76-84 is commented out): value_without_edit_distance is the reference-faithful path; "
                     "value_reference_pool6 = the same rates on min(6, cores) processes, the pool size the reference itself "
                     "uses (src/SVscope.py:158-161)")


# ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--windows", type=int, default=CONFIGS1_BATCH, help="windows per rank and step (configs[1]: 1000)")
    ap.add_argument("--budget-s", type=float, default=480.0,
                    help="seconds the K timed steps may take; the per-step batch shrinks to fit (0: never shrink)")
    ap.add_argument("--total-s", type=float, default=740.0,
                    help="seconds the whole process should take (the driver allows 870 s per run); the timed budget "
                         "shrinks if start-up, warm-up and calibration took long (0: ignore)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-edit-distance", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    ap.add_argument("--poa-threads", type=int, default=0)
    ap.add_argument("--ring-rows", type=int, default=0)
    ap.add_argument("--chunks", type=int, default=3)
    args = ap.parse_args()

    if args.impl == "reference":
        return main_reference(args)

    import torch
    rank, world, local = dist_setup(args.gpus)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: svscope_b200 has no CPU fallback (use --impl reference "
                         "for the CPU baseline)")
    from svscope_b200 import _lib
    from svscope_b200.batch import localgraph_batch, upload_windows
    ctx = _lib.Context(local)
    if args.poa_threads:
        ctx.set_option("poa_threads", args.poa_threads)
    if args.ring_rows:
        ctx.set_option("ring_rows", args.ring_rows)
    ed = not args.no_edit_distance

    t_gen = time.perf_counter()
    windows_all = make_batch(args.windows, rank)
    t_gen = time.perf_counter() - t_gen

    def run(wins, reads):
        return localgraph_batch(wins, ctx=ctx, reads=reads, edit_distance=ed, chunks=args.chunks)

    # ---- warm-up: W steps on the 64 cheapest windows (context, arena, kernel images, allocator pools) ----
    slice_w = sorted(windows_all, key=model_cells)[:64]
    slice_reads = upload_windows(ctx, slice_w)
    t_w = time.perf_counter()
    for _ in range(max(args.warmup, 1)):
        run(slice_w, slice_reads)
    t_w = time.perf_counter() - t_w
    # the edit-distance kernel alone on the same slice (in a step it runs beside the window kernels and its
    # event time is that of a kernel sharing the SMs): the figure its own roofline fraction is quoted on
    myers_alone = None
    if ed:
        from svscope_b200.batch import edit_distance_matrices
        torch.cuda.synchronize()
        sbase = np.concatenate([[0], np.cumsum([len(w[0]) for w in slice_w])])
        _, myers_alone = edit_distance_matrices(ctx, slice_reads, [list(range(int(sbase[i]) + 1, int(sbase[i + 1])))
                                                                   for i in range(len(slice_w))])
    slice_reads.close()

    # ---- per-step batch: the configs[1] batch, or its first n windows if K steps would not fit ------
    # The driver allows 870 s for the whole process.  The timed steps get what `--total-s` leaves after the
    # time already spent (process start, imports, batch generation, warm-up, calibration) and the work that
    # follows them (the host-buffer step, the CPU sample at N=1), at most `--budget-s`.
    n_step = len(windows_all)
    calib = None
    budget_s = args.budget_s
    if args.budget_s > 0 and args.steps > 1:
        reads_all = upload_windows(ctx, windows_all)
        barrier_sync()
        t_c = time.perf_counter()
        run(windows_all, reads_all)
        torch.cuda.synchronize()
        t_full = max_over_ranks(time.perf_counter() - t_c)
        calib = {"full_batch_windows": len(windows_all), "full_batch_step_s": t_full,
                 "full_batch_windows_per_s_per_gpu": len(windows_all) / t_full}
        if args.total_s > 0:
            after = 15.0 + (args.cpu_budget + 8.0 if (world == 1 and not args.no_cpu_baseline) else 0.0)
            avail = args.total_s - process_age_s() - after - 0.5 * t_full      # 0.5 t_full: the second calibration step
            share = args.steps / (args.steps + (0.0 if args.no_e2e else 1.1))    # the host-buffer step follows
            budget_s = min_over_ranks(max(30.0, min(budget_s, avail * share)))
        if t_full * args.steps > budget_s:
            n_step = int(len(windows_all) * budget_s / (t_full * args.steps) / 1.04)
            n_step = int(min_over_ranks(float(max(64, min(len(windows_all), n_step)))))
        if n_step < len(windows_all):
            # a step of n windows is not n/1000 of the full step (fewer windows per SM, longer tail): one
            # untimed step at the chosen size, then shrink once more if it says so
            reads_all.close()
            reads_all = None
            sub = windows_all[:n_step]
            reads_sub = upload_windows(ctx, sub)
            barrier_sync()
            t_c = time.perf_counter()
            run(sub, reads_sub)
            torch.cuda.synchronize()
            t_sub = max_over_ranks(time.perf_counter() - t_c)
            reads_sub.close()
            calib.update(second_calibration_windows=n_step, second_calibration_step_s=t_sub)
            if t_sub * args.steps > 1.02 * budget_s:
                n_step = int(min_over_ranks(float(max(64, int(n_step * budget_s / (t_sub * args.steps))))))
        calib["timed_budget_s"] = budget_s
    else:
        reads_all = upload_windows(ctx, windows_all)
    windows = windows_all[:n_step]
    reads = reads_all if reads_all is not None else upload_windows(ctx, windows)     # resident in HBM before timing
    read_bytes = reads.nbytes

    # ---- timed region ------------------------------------------------------------------------------
    t_before_timed = process_age_s()
    sampler = ClockSampler(local)
    barrier_sync()
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    agg, stage = {}, {}
    out = None
    for _ in range(args.steps):
        out = run(windows, reads)
        for k, v in out.stats.items():
            agg[k] = agg.get(k, 0.0) + v
        for k, v in out.timings.items():
            stage[k] = stage.get(k, 0.0) + v
    torch.cuda.synchronize()
    ev1.record()
    ev1.synchronize()
    barrier_sync()
    wall = time.perf_counter() - t0
    dev_s = ev0.elapsed_time(ev1) / 1e3
    clocks = sampler.stop() if rank == 0 else None
    t_max = max_over_ranks(max(dev_s, 1e-9))
    t_min = min_over_ranks(max(dev_s, 1e-9))
    total_windows = sum_over_ranks(float(len(windows) * args.steps))
    value = total_windows / t_max

    # ---- end to end: the same step with host buffers (upload + every result copy inside) -----------
    e2e = None
    if not args.no_e2e:
        barrier_sync()
        t1 = time.perf_counter()
        o2 = localgraph_batch(windows, ctx=ctx, reads=None, edit_distance=ed, chunks=args.chunks)
        torch.cuda.synchronize()
        barrier_sync()
        t_e2e = max_over_ranks(time.perf_counter() - t1)
        s2 = o2.stats
        h2d = read_bytes + s2["poa_h2d_bytes"] + s2.get("feat_h2d_bytes", 0) + s2.get("em_h2d_bytes", 0)
        d2h = (s2["poa_d2h_bytes"] + s2.get("poa_copy_bytes", 0) + s2.get("feat_d2h_bytes", 0) + s2.get("em_d2h_bytes", 0)
               + s2.get("ed_d2h_bytes", 0))
        assert o2.records == out.records      # resident and host-buffer paths give the same records
        e2e = {"value": sum_over_ranks(float(len(windows))) / t_e2e, "unit": UNIT,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": 1,
               "note": "one step after the timed steps: reads uploaded from host memory, MSA / consensus / matrices copied back"}

    if rank != 0:
        return
    steps = args.steps
    st = {k: v / steps for k, v in agg.items()}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    alu = ctx.int_alu_probe()
    # ---- roofline of the dominant kernel (poa_window_kernel): integer ALU --------------------------
    # Launch durations by CUDA events on the launch streams overlap (the kernels of the sub-batches queue
    # behind one another), so the kernel's own time is taken from its in-kernel clocks: cycles every
    # CTA spent on its windows, summed, / resident CTAs / SM clock = the launch time at perfect packing.
    n_launch = max(1.0, st["poa_dp_launches"])
    cta_cycles = sum(st.get("poa_cyc_" + k, 0.0) for k in ("export", "dp", "traceback", "merge", "rank", "finish"))
    sm_hz = float((clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)) * 1e6
    resident = float(ctx.get_option("sm_count")) * max(1, 512 // int(ctx.get_option("poa_threads")))
    kern_s = max(1e-9, cta_cycles / resident / sm_hz)
    event_s = st["poa_dp_ms"] / 1e3
    ops_per_cell = 18.0                                 # 8*indeg+10 integer add/max per cell at in-degree 1
    peak_gcups = alu["addmax"] * 2.0 / ops_per_cell     # fused add+max counts as two algorithmic ops
    gcups_nominal = st["poa_cells"] / kern_s / 1e9
    gcups_eval = st["poa_eval_cells"] / kern_s / 1e9
    algo_bytes = st["poa_algo_bytes"]
    roofline = {
        "bound": "int_alu", "kernel": f"poa_window_kernel<{ctx.get_option('poa_threads')},8>",
        "achieved": gcups_nominal, "peak": peak_gcups, "unit": "GCUPS", "frac": gcups_nominal / peak_gcups,
        "achieved_evaluated_cells": gcups_eval, "frac_evaluated_cells": gcups_eval / peak_gcups,
        "evaluated_fraction": st["poa_eval_cells"] / max(1.0, st["poa_cells"]),
        "traffic": NCU_TRAFFIC["bytes"] / NCU_TRAFFIC["alignments"] * (st["poa_alignments"] / n_launch),
        "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture (2 windows, 120 alignments of "
                        "3 kb reads: the working set of so small a launch stays in the 126 MB L2) scaled to the alignments of one "
                        "launch here; at bench size the traceback codes (1-2 B per evaluated cell) do go to HBM: see hbm.implementation_bytes",
        "launches_per_step": n_launch, "avg_launch_ms": kern_s / n_launch * 1e3,
        "kernel_seconds_per_step": kern_s, "event_seconds_per_step_overlapping": event_s,
        "peak_source": "measured in this run: fused add+max issue rate (svs_int_alu_probe) x 2 / 18 ops per cell",
        "probe_gops": alu, "ops_per_cell": ops_per_cell,
        "note": "achieved = NOMINAL cells (sum (|V|+1)(L+1), what the CPU engine fills) / kernel seconds, where kernel seconds = "
                "in-kernel cycle counters of all CTAs / resident CTAs / SM clock under load (the CUDA-event durations of the "
                "launches overlap, because the sub-batch kernels queue behind one another: event_seconds_per_step_overlapping); "
                "the time includes the in-kernel graph phases (export, traceback, merge, rank order: poa.phase_share); "
                "exact pruning evaluates evaluated_fraction of the nominal cells",
        "hbm": {"bound": "hbm", "achieved": algo_bytes / kern_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": algo_bytes / kern_s / 1e9 / hbm_peak, "peak_source": peak_src,
                "algorithmic_bytes_per_step": algo_bytes,
                "implementation_bytes": 1.35 * st["poa_eval_cells"],
                "note": "algorithmic bytes = read + rank-ordered graph + alignment path (SURVEY 8d); implementation_bytes = "
                        "traceback codes written (1 B per evaluated cell of single-predecessor rows, 2 B otherwise)"}}
    ed_ms = max(st.get("ed_ms", 0.0), 1e-9)
    myers_peak = alu["xor"] / 0.53
    kernels = {
        "myers_kernel": {"bound": "int_alu", "unit": "GCUPS", "peak": myers_peak,
                         "achieved": (myers_alone["cells"] / max(myers_alone["ms"], 1e-9) / 1e6) if myers_alone else None,
                         "frac": (myers_alone["cells"] / max(myers_alone["ms"], 1e-9) / 1e6 / myers_peak) if myers_alone else None,
                         "achieved_in_step": st.get("ed_cells", 0.0) / ed_ms / 1e6,
                         "frac_in_step": st.get("ed_cells", 0.0) / ed_ms / 1e6 / myers_peak,
                         "note": "cells = L_i * L_j per pair; ~17 64-bit logic ops per 64-cell word step = 0.53 32-bit ops per cell "
                                 "against the measured LOP3 issue rate; achieved / frac: the kernel ALONE on the device (all read "
                                 "pairs of the 64 warm-up windows, CUDA events on its stream); *_in_step: its event time inside a "
                                 "step, where it shares the SMs with the window kernels (not a kernel rate)"},
        "em_kernel": {"bound": "latency", "note": "FP64, one CTA per (window, K): < 2 % of a step, occupancy/latency bound; "
                                                  "no roofline fraction claimed"}}
    cyc = {k: st.get("poa_cyc_" + k, 0.0) for k in ("export", "dp", "traceback", "merge", "rank", "finish")}
    cyc_tot = max(1.0, sum(cyc.values()))
    wl = max(1.0, st.get("poa_wcyc_loop", 0.0) + st.get("poa_wcyc_wait_end", 0.0))
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_max / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32 (POA DP, edit distance u32 bit-vectors), f64 (mixture model)", "data": "synthetic",
        "config": workload_config(len(windows_all), ed, world),
        "run": {"windows_per_step": len(windows),
                "step_batch_note": ("the configs[1] batch" if len(windows) == CONFIGS1_BATCH else
                                    f"--windows {len(windows_all)}: a smaller batch of the same distribution" if calib is None or len(windows) == len(windows_all) else
                                    f"first {len(windows)} windows of the {len(windows_all)}-window batch per step, so that "
                                    f"{args.steps} steps fit {budget_s:.0f} s (the full batch takes "
                                    f"{calib['full_batch_step_s']:.1f} s per step: calibration)"),
                "warmup_note": f"{max(args.warmup, 1)} warm-up steps on the 64 cheapest windows of the batch ({t_w:.1f} s)",
                "host_work_per_rank": "one Python process (+ one thread driving the edit-distance kernels)",
                "poa_threads": ctx.get_option("poa_threads"), "ring_rows": ctx.get_option("ring_rows"),
                "sub_batches": int(out.stats.get("sub_batches", 1)),
                "seconds_before_timed_region": round(t_before_timed, 1)},
        "calibration": calib,
        "e2e": e2e,
        "gpu_launches": int(agg.get("poa_dp_launches", 0) + agg.get("aux_launches", 0)),
        "clocks": clocks,
        "roofline": roofline,
        "kernels": kernels,
        "stage_seconds_per_step": {k: round(v / steps, 3) for k, v in stage.items()},
        "rank_time_skew": t_max / max(t_min, 1e-9),
        "wall_s_timed": wall, "gen_s": t_gen,
        "poa": {"cells_per_step": st["poa_cells"], "evaluated_cells_per_step": st["poa_eval_cells"],
                "alignments_per_step": st["poa_alignments"],
                "exported_row_frac": st["poa_exported_rows"] / max(1.0, st["poa_rows"]),
                "pruning_retries_per_step": st.get("poa_prune_retries", 0.0),
                "failed_windows_per_step": st.get("poa_failed_windows", 0.0),
                "phase_share": {k: v / cyc_tot for k, v in cyc.items()},
                "dp_warps": {"working": (st.get("poa_wcyc_loop", 0.0) - st.get("poa_wcyc_wait_left", 0.0)
                                         - st.get("poa_wcyc_wait_right", 0.0)) / wl,
                             "waiting_for_handover": st.get("poa_wcyc_wait_left", 0.0) / wl,
                             "waiting_at_row_start": st.get("poa_wcyc_wait_right", 0.0) / wl,
                             "waiting_at_end": st.get("poa_wcyc_wait_end", 0.0) / wl,
                             "note": "lane-0 clocks of every warp; waiting_at_row_start = left neighbour not yet one row behind or "
                                     "right neighbour more than 32 rows behind (the strips ramp up and down one after the other: "
                                     "profiles/r02_dp_phase_profile_summary.txt)"},
                "host_ms_per_step": 0.0},
        "edit_distance": {"cells_per_step": st.get("ed_cells", 0.0), "kernel_ms_per_step": st.get("ed_ms", 0.0)},
        "em_output_windows": sum(r[-1].endswith("EMOutput") for r in out.records),
        "em_redraw_fraction": st.get("em_redraw_windows", 0.0) / max(1.0, st.get("windows_em", 1.0)),
    }
    if not args.no_cpu_baseline and world == 1:
        poa_cpw = st["poa_cells"] / len(windows)
        ed_cpw = st.get("ed_cells", 0.0) / len(windows)
        cb = cpu_baseline(windows, poa_cpw, ed_cpw, args.cpu_budget, with_ed=ed)
        model = float(np.mean([model_cells(w) for w in windows]))
        cb["cost_model_residual"] = {"model_msa_cells_per_window": model,
                                     "counted_cells_per_window_msa_and_consensus": poa_cpw,
                                     "note": "the model covers the window MSA only; the reference arm adds the measured "
                                             "consensus share"}
        line["cpu_baseline"] = cb
    print(json.dumps(line))


def _shutdown():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


# consensus POA cells / window-MSA cells and edit-distance cells per window of the configs[1] batch,
# counted on the device (profiles/r02_bench_*.json: poa.cells_per_step against model_cells)
CONSENSUS_SHARE = 0.23


def main_reference(args):
    """The CPU path of the reference on this box's host cores (rank 0 only), timed ONCE on a bounded sample."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    windows = make_batch(args.windows, 0)
    ed = not args.no_edit_distance
    poa_cpw = float(np.mean([model_cells(w) for w in windows])) * (1.0 + CONSENSUS_SHARE)
    ed_cpw = float(np.mean([sum(len(a) * len(b) for i, a in enumerate(w[0][1:]) for b in w[0][i + 2:]) for w in windows[:50]]))
    cb = cpu_baseline(windows, poa_cpw, ed_cpw, args.cpu_budget, with_ed=ed)
    value = cb["value"]
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": env_int("WORLD_SIZE", 1),
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["wall_s"] * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32/f64 (CPU)",
            "data": "synthetic",
            "config": workload_config(args.windows, ed, env_int("WORLD_SIZE", 1)),
            "run": {"sample_note": "one bounded sample, timed once (the driver's steps/warmup are echoed, not repeated); "
                                   "cells per window from the growth model of bench.model_cells x (1 + consensus share)"},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "value_without_edit_distance",
                                                "value_with_edit_distance", "value_reference_pool6", "value_one_core",
                                                "poa_gcups_per_core", "ed_gcups_per_core")},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference = negi2331026/SVScope Python path; its spoa.poa lives in the pyspoa wheel that cannot be "
                    "installed offline, so the timed code is the oracle port (scalar five-matrix POA restatement, bit-parallel "
                    "Levenshtein), one process per window on all host cores"}
    print(json.dumps(line))


if __name__ == "__main__":
    try:
        main()
    finally:
        _shutdown()
