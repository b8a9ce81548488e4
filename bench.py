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
`--impl reference`  times the CPU path (oracle port of the reference: pyspoa is not
          installable offline) on a bounded sample with all host cores.
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
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # before any CUDA context (see svscope_b200/__init__.py)

NCU_TRAFFIC_BYTES_PER_LAUNCH = 15.59e9   # profiles/r01_poa_persistent_kernel_final_ncu_full.txt (one launch, 296 alignments)
METRIC = "localGraph windows/sec"
UNIT = "windows/s"
WORKLOAD = "configs[1]: synthetic INS/DEL windows, 30 tumor + 30 normal reads, 5-15 kb, 5% error"


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
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        busy = [v for v in sm if v > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


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


def max_over_ranks(x: float) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return x
    t = torch.tensor([x], dtype=torch.float64, device="cuda" if torch.cuda.is_available() else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x: float) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return x
    t = torch.tensor([x], dtype=torch.float64, device="cuda" if torch.cuda.is_available() else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def make_batch(n_windows, rank):
    from svscope_b200 import synth
    return [synth.make_c2_window(rank * n_windows + i) for i in range(n_windows)]


# ------------------------------------------------------------------------------------------
# CPU path (oracle port of the reference) on a bounded sample
# ------------------------------------------------------------------------------------------
def _cpu_one(w):
    from oracle import oracle as O
    t0 = time.perf_counter()
    rec = O.decision(w[4], w[0], w[1], w[2], w[3])
    O.levenshtein_matrix(w[0][1:], bitparallel=True)
    return rec[-1], time.perf_counter() - t0


def cpu_sample(windows, budget_s=25.0, cores=None):
    """Cheapest windows first until the cost model predicts ~budget_s per core."""
    from svscope_b200 import synth
    cores = cores or os.cpu_count() or 1
    costs = np.array([synth.window_cost(w) for w in windows])
    order = np.argsort(costs)
    rate = 1.5e8  # cost units per core-second (measured: ~0.1-0.2 GCUPS scalar five-matrix DP)
    picked, load = [], 0.0
    for i in order:
        if picked and (load + costs[i]) / rate > budget_s * cores:
            break
        picked.append(int(i))
        load += costs[i]
        if len(picked) >= 4 * cores:
            break
    return picked, costs


def run_cpu(windows, budget_s, cores=None):
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    picked, costs = cpu_sample(windows, budget_s, cores)
    sample = [windows[i] for i in picked]
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(min(cores, len(sample))) as pool:
        res = pool.map(_cpu_one, sample, chunksize=1)
    dt = time.perf_counter() - t0
    raw = len(sample) / dt
    scale = float(costs[picked].mean() / costs.mean())   # sample is cheaper than the batch average
    return dict(value=raw * scale, raw_windows_per_s=raw, seconds=dt, n=len(sample), cores=min(cores, len(sample)),
                cost_scale=scale,
                sample=f"{len(sample)} cheapest of {len(windows)} windows (mean cost {scale:.2f}x of the batch mean; "
                       f"value = sample windows/s x that ratio), one process per window on {min(cores, len(sample))} cores, "
                       f"{dt:.1f} s wall")


# ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--windows", type=int, default=1000, help="windows per rank and step (configs[1]: 1000)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-edit-distance", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=25.0)
    ap.add_argument("--workers", type=int, default=0)
    ap.add_argument("--poa-threads", type=int, default=0)
    ap.add_argument("--ring-rows", type=int, default=0)
    ap.add_argument("--poa-cols", type=int, default=0)
    ap.add_argument("--streams", type=int, default=0)
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
    ncpu = os.cpu_count() or 8
    workers = args.workers or max(2, min(12, (ncpu // max(1, world)) - 1))
    ctx.set_option("workers", workers)
    if args.poa_threads:
        ctx.set_option("poa_threads", args.poa_threads)
    if args.ring_rows:
        ctx.set_option("ring_rows", args.ring_rows)
    if args.poa_cols:
        ctx.set_option("poa_cols", args.poa_cols)
    if args.streams:
        ctx.set_option("streams", args.streams)
    ed = not args.no_edit_distance

    t_gen = time.perf_counter()
    windows = make_batch(args.windows, rank)
    t_gen = time.perf_counter() - t_gen
    reads = upload_windows(ctx, windows)                      # resident in HBM before timing
    read_bytes = reads.nbytes

    def step():
        return localgraph_batch(windows, ctx=ctx, reads=reads, edit_distance=ed)

    def e2e_step():
        """Same step through the public batch call with host buffers: the reads are uploaded from
        page-locked host memory and every result is copied back inside the timed region."""
        barrier_sync()
        t0 = time.perf_counter()
        o2 = localgraph_batch(windows, ctx=ctx, reads=None, edit_distance=ed)
        torch.cuda.synchronize()
        barrier_sync()
        t_e2e = max_over_ranks(time.perf_counter() - t0)
        s2 = o2.stats
        h2d = read_bytes + s2["poa_h2d_bytes"] + s2.get("feat_h2d_bytes", 0) + s2.get("em_h2d_bytes", 0)
        d2h = s2["poa_d2h_bytes"] + s2.get("feat_d2h_bytes", 0) + s2.get("em_d2h_bytes", 0) + s2.get("ed_d2h_bytes", 0)
        return o2, {"value": sum_over_ranks(float(args.windows)) / t_e2e, "unit": UNIT,
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "note": "timed on the last warm-up step" if args.warmup > 0 else "timed after the timed steps"}

    # warm-up steps; the last one doubles as the end-to-end measurement (it is itself preceded by
    # warm-up steps), which keeps the default run within minutes
    e2e, o2 = None, None
    for w in range(args.warmup):
        if w == args.warmup - 1 and not args.no_e2e:
            o2, e2e = e2e_step()
        else:
            out = step()
    sampler = ClockSampler(local)
    barrier_sync()
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    agg = {}
    for _ in range(args.steps):
        out = step()
        for k, v in out.stats.items():
            agg[k] = agg.get(k, 0.0) + v
    torch.cuda.synchronize()
    ev1.record()
    ev1.synchronize()
    barrier_sync()
    wall = time.perf_counter() - t0
    dev_s = ev0.elapsed_time(ev1) / 1e3
    clocks = sampler.stop() if rank == 0 else None
    t_max = max_over_ranks(max(dev_s, 1e-9))
    total_windows = sum_over_ranks(float(args.windows * args.steps))
    value = total_windows / t_max
    if e2e is None and not args.no_e2e:
        o2, e2e = e2e_step()
    if o2 is not None:
        assert o2.records == out.records      # resident and host-buffer paths give the same records

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
    dp_s = st["poa_dp_ms"] / 1e3
    n_launch = max(1.0, st["poa_dp_launches"])
    algo_bytes_per_launch = st["poa_algo_bytes"] / n_launch
    avg_launch_s = dp_s / n_launch
    achieved_gbs = algo_bytes_per_launch / avg_launch_s / 1e9
    roofline = {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved_gbs / hbm_peak, "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH, "peak_source": peak_src,
                "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of ONE launch (296 alignments, launch 14 of a depth-8 probe: "
                                "profiles/r01_poa_persistent_kernel_final_ncu_full.txt); almost all of it is "
                                "traceback codes (1-2 B per DP cell), which are implementation traffic, not algorithmic bytes",
                "kernel": "poa_persistent_kernel<256,8>", "launches_per_step": n_launch,
                "avg_launch_ms": avg_launch_s * 1e3,
                "note": "algorithmic bytes = read + rank-ordered graph + alignment path (SURVEY 8d); the kernel is "
                        "integer-ALU bound, see roofline_alu; launches of different worker streams overlap, so the "
                        "per-launch event time is an upper bound of the exclusive time"}
    alu = ctx.int_alu_probe()
    ops_per_cell = 18.0   # 8*indeg+10 integer add/max per cell at in-degree 1
    poa_wall = max(1e-9, (out.timings["poa_msa"] + out.timings["poa_consensus"]))
    gcups = st["poa_cells"] / poa_wall / 1e9
    peak_gcups = alu["addmax"] * 2.0 / ops_per_cell   # fused add+max counts as two algorithmic ops
    roofline_alu = {"bound": "int_alu", "achieved": gcups, "peak": peak_gcups, "unit": "GCUPS", "frac": gcups / peak_gcups,
                    "probe_gops": alu, "ops_per_cell": ops_per_cell,
                    "note": "NOMINAL cells = sum (|V|+1)(L+1) over alignments (what the CPU engine fills); exact pruning evaluates "
                            "about 30 % of them; achieved over the wall time of the two POA stages (last timed step); "
                            "peak = measured fused add+max issue rate x 2 / 18 ops per cell"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_max / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32 (POA DP, edit distance u32 bit-vectors), f64 (mixture model)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "windows_per_gpu": args.windows, "reads_per_window": 60,
                   "edit_distance_matrix": ed, "l2": "inputs larger than L2 (reads + traceback codes >> 126 MB per step)",
                   "parallelism": f"windows sharded over {world} GPU(s), no collective", "host_workers": workers,
                   "poa_threads": ctx.get_option("poa_threads"), "poa_cols": ctx.get_option("poa_cols"),
                   "ring_rows": ctx.get_option("ring_rows")},
        "e2e": e2e,
        "gpu_launches": int(agg.get("poa_dp_launches", 0) + agg.get("poa_tb_launches", 0) + agg.get("aux_launches", 0)),
        "clocks": clocks,
        "roofline": roofline,
        "roofline_alu": roofline_alu,
        "stage_seconds_last_step": {k: round(v, 3) for k, v in out.timings.items()},
        "wall_s_timed": wall, "gen_s": t_gen,
        "poa": {"cells_per_step": st["poa_cells"], "alignments_per_step": st["poa_alignments"],
                "exported_row_frac": st["poa_exported_rows"] / max(1.0, st["poa_rows"]),
                "pruning_retries_per_step": st.get("poa_prune_retries", 0.0),
                "host_ms_per_step": {k: st.get("poa_" + k, 0.0) for k in
                                     ("host_wait_ms", "host_merge_ms", "host_plan_ms", "host_pack_ms", "launch_ms")}},
        "edit_distance": {"cells_per_step": st["ed_cells"], "kernel_ms_per_step": st["ed_ms"],
                          "gcups": st["ed_cells"] / max(st["ed_ms"], 1e-9) / 1e6},
        "em_output_windows": sum(r[-1].endswith("EMOutput") for r in out.records),
    }
    # informational, outside the timed region: the step after the Raw.bed,
    # MisScore alignments of the records this step produced
    try:
        from svscope_b200 import PairwiseCompare as PC
        mpairs = [p for r in out.records if str(r[9]) == "NormalOutput|EMOutput"
                  for p in PC._record_pairs(str(r[3]), str(r[6]))]
        mst = {}
        t_m = time.perf_counter()
        PC.misscore_pairs(mpairs, ctx=ctx, stats=mst)
        t_m = time.perf_counter() - t_m
        line["misscore_after_step"] = {"pairs": len(mpairs), "cells": mst.get("cells", 0.0),
                                       "kernel_ms": mst.get("kernel_ms", 0.0), "wall_s": t_m,
                                       "gcups_kernel": mst.get("cells", 0.0) / max(mst.get("kernel_ms", 0.0), 1e-9) / 1e6,
                                       "note": "not part of `value`: PairwiseCompare.MisScorePipe's alignments "
                                               "(svs_misscore_pairs) on the records of the last step"}
    except Exception as exc:  # informational only
        line["misscore_after_step"] = {"error": repr(exc)}
    if not args.no_cpu_baseline and world == 1:
        cb = run_cpu(windows, args.cpu_budget)
        line["cpu_baseline"] = {"value": cb["value"], "unit": UNIT, "cores": cb["cores"], "kind": "port",
                                "sample": cb["sample"], "raw_sample_windows_per_s": cb["raw_windows_per_s"],
                                "note": "CPU restatement of pyspoa (real pyspoa 0.2.1 SIMD engine unavailable offline) + "
                                        "numpy port of ReadsCluster/DecisionMaker + bit-parallel Levenshtein"}
    print(json.dumps(line))


def _shutdown():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


def main_reference(args):
    """The CPU path of the reference on this box's host cores (rank 0 only)."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    windows = make_batch(args.windows, 0)
    cores = os.cpu_count() or 1
    times, last = [], None
    for i in range(args.warmup + args.steps):
        budget = args.cpu_budget if i >= args.warmup else min(args.cpu_budget, 8.0)
        last = run_cpu(windows, budget, cores)
        if i >= args.warmup:
            times.append(last)
    value = float(np.mean([t["value"] for t in times]))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": env_int("WORLD_SIZE", 1),
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean([t["seconds"] for t in times])) * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32/f64 (CPU)",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "windows_per_gpu": args.windows, "reads_per_window": 60,
                       "edit_distance_matrix": True},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": last["cores"], "kind": "port", "sample": last["sample"]},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference = negi2331026/SVScope Python path; its spoa.poa lives in the pyspoa wheel that cannot be "
                    "installed offline, so the timed code is the oracle port (scalar five-matrix POA restatement + "
                    "numpy mixture model + bit-parallel Levenshtein), one process per window on all host cores"}
    print(json.dumps(line))


if __name__ == "__main__":
    try:
        main()
    finally:
        _shutdown()
