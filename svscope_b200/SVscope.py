"""Entry points of the localGraph stage with the reference's names and outputs
(src/SVscope.py: localGraph :118-183, localGraph_npz :185-239, callsomaticSV :341-356).

    python -m svscope_b200.SVscope localGraph_npz -s <dir with *.npz> -t TUMOR -n NORMAL [-C]

The 10-column ``<T>.vs.<N>.TandemRepeat.Raw.bed`` is written exactly as the reference does
(tab-joined ``str()`` of the record fields, then ``sort -k1,1 -k2,2n``), so the downstream
stages (AlnFeature, OutVCF) consume it unchanged.  The per-window process pool of the
reference (Pool of at most 6 workers, :158-161) is replaced by GPU batching; with several
GPUs (one process per GPU, torchrun) the windows are sharded and gathered on the host."""
from __future__ import annotations

import argparse
import logging
import os
import re
import time

import numpy as np

from . import shard as _shard
from .batch import localgraph_batch
from .synth import window_cost

logging.basicConfig(level=logging.INFO, format="%(asctime)s - %(levelname)s - %(message)s")

BATCH_WINDOWS = 1024


def raw_bed_name(TSampleID: str, NSampleID: str) -> str:
    return "%s.vs.%s.TandemRepeat.Raw.bed" % ("-".join(TSampleID.split(",")), "-".join(NSampleID.split(",")))


def _format(rec) -> str:
    return "\t".join([str(x) for x in rec]) + "\n"


def write_raw_bed(path: str, records, append: bool = False) -> None:
    with open(path, "a" if append else "w") as f:
        for rec in records:
            f.write(_format(rec))
    os.system("sort -k1,1 -k2,2n {p} -o {p}".format(p=path))


def _ranks():
    """(rank, world size).  Under torchrun (WORLD_SIZE > 1 in the environment) the process group
    is created here if the caller has not done so: without it every process would believe it is
    rank 0 of 1, compute all windows and fight over the same part file."""
    import torch.distributed as dist
    if not dist.is_available():
        return 0, 1
    if not dist.is_initialized() and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group("gloo")     # host-side gathers and barriers only: no collective on the data path
        import atexit
        atexit.register(lambda: dist.is_initialized() and dist.destroy_process_group())
    if dist.is_initialized() and dist.get_world_size() > 1:
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def _my_indices(windows):
    rank, world = _ranks()
    if world == 1:
        return list(range(len(windows)))
    return _shard.my_shard([window_cost(w) for w in windows], rank, world)


def iter_batches(windows, mine, batch_windows: int = BATCH_WINDOWS, **kw):
    """Yields (window indices, records) of one GPU batch after the other."""
    for b in range(0, len(mine), batch_windows):
        idx = mine[b:b + batch_windows]
        yield idx, localgraph_batch([windows[i] for i in idx], **kw).records


def run_windows(windows, batch_windows: int = BATCH_WINDOWS, **kw):
    """Records for a list of windows, sharded over the ranks of the current process group
    (the full list on rank 0, None on the other ranks)."""
    mine = _my_indices(windows)
    local = []
    for _, recs in iter_batches(windows, mine, batch_windows, **kw):
        local.extend(recs)
    if _ranks()[1] == 1:
        return local
    return _shard.gather_records(mine, local, len(windows), group=_shard.host_group())


def _finished_records(paths):
    done = set()
    for p in paths:
        if os.path.exists(p):
            with open(p) as fh:
                done.update("\t".join(x.strip().split("\t")[0:3]) for x in fh.readlines() if x.strip())
    return done


def localGraph_npz(args):
    """Reference :185-239: all ``*.npz`` batches in ``args.savedir`` -> Raw.bed; ``--Continue``
    skips windows whose TDRecord is already in the output (:195-200, :213).

    The records of every GPU batch are appended at once to ``<Raw.bed>.part<rank>`` (flushed to
    disk), and the parts are merged into the sorted Raw.bed at the end, so an interrupted run
    loses at most the batch in flight: ``--Continue`` also counts what the parts hold."""
    import glob
    import torch.distributed as dist
    logging.info("Local Graph : Start working")
    t0 = time.time()
    rank, world = _ranks()
    group = _shard.host_group() if world > 1 else None
    path = os.path.join(args.savedir, raw_bed_name(args.TSampleID, args.NSampleID))
    resume = bool(getattr(args, "Continue", False))
    if rank == 0 and not resume:
        for stale in glob.glob(glob.escape(path) + ".part*"):
            os.remove(stale)
    if world > 1:
        dist.barrier(group=group)
    finished = _finished_records([path] + sorted(glob.glob(glob.escape(path) + ".part*"))) if resume else set()
    if world > 1:
        dist.barrier(group=group)  # every rank has read the parts before anyone appends
    # pass 1: (file, row, cost) of every unfinished window - sequence lengths only, one file in memory at a time
    files = [name for name in sorted(os.listdir(args.savedir)) if re.search("npz", name)]
    meta = []
    for fi, name in enumerate(files):
        dat = np.load(os.path.join(args.savedir, name), allow_pickle=True)["DatSet"]
        for i in range(dat.shape[0]):
            row = dat[i]
            if finished and row[4] in finished:
                continue
            meta.append((fi, i, window_cost(row)))
        del dat
    mine = set(_shard.my_shard([m[2] for m in meta], rank, world)) if world > 1 else set(range(len(meta)))
    mine_by_file = {}
    for k, (fi, i, _) in enumerate(meta):
        if k in mine:
            mine_by_file.setdefault(fi, []).append(i)
    # pass 2: stream my windows, file by file, in full GPU batches (no rank ever holds all windows)
    with open(path + ".part%d" % rank, "a") as part:
        pending = []

        def flush(rows):
            recs = localgraph_batch(rows).records
            part.write("".join(_format(r) for r in recs))
            part.flush()
            os.fsync(part.fileno())

        for fi in sorted(mine_by_file):
            dat = np.load(os.path.join(args.savedir, files[fi]), allow_pickle=True)["DatSet"]
            for i in mine_by_file[fi]:
                pending.append(list(dat[i]))
                if len(pending) >= BATCH_WINDOWS:
                    flush(pending)
                    pending = []
            del dat
        if pending:
            flush(pending)
    if world > 1:
        dist.barrier(group=group)
    if rank == 0:
        parts = sorted(glob.glob(glob.escape(path) + ".part*"))
        with open(path, "a" if resume else "w") as out:
            for p in parts:
                with open(p) as fh:
                    out.write(fh.read())
        os.system("sort -k1,1 -k2,2n {p} -o {p}".format(p=path))
        for p in parts:
            os.remove(p)
    logging.info(f"Local Graph : work finished with {(time.time() - t0) / 3600} hour")
    return path


def localGraph(args):
    """Reference :118-183 extracts every window from the BAM files with pysam before the
    decision.  That feeder is outside this package; dump the windows with the reference's
    SomTDDetector_AimDatFetch.py (npz rows) and run ``localGraph_npz``."""
    raise NotImplementedError(
        "localGraph needs the pysam/BAM feeder (src/DataScanner.py:222-247), which is out of scope of the "
        "accelerated path; use localGraph_npz on npz batches written by SomTDDetector_AimDatFetch.py")


def somatic_bed(rawBedFile: str, savedir: str, TSampleID: str) -> str:
    """The MisScore head of the reference's ``AlnFeature`` (src/SVscope.py:282-286):
    ``PairwiseCompare.MisScorePipe(rawBedFile).drop_duplicates()`` plus the ``ABSMisScore``
    column, written to ``<savedir>/<T>.Somatic.bed`` (tab-separated, no header, no index).  The
    alignments run on the GPU (svs_misscore_pairs); the rest of AlnFeature (bed.gz/sqlite
    coverage and mapQ features, random forest) is the reference's and reads this file."""
    from . import PairwiseCompare
    df = PairwiseCompare.MisScorePipe(rawBedFile).drop_duplicates()
    df['ABSMisScore'] = df['MisScore'].apply(lambda x: abs(x))
    out = os.path.join(savedir, '%s.Somatic.bed' % TSampleID)
    df.to_csv(out, sep="\t", index=False, header=None)
    return out


def callsomaticSV(args):
    """Reference :341-356 = localGraph + AlnFeature.  The localGraph half (npz batches in
    ``args.savedir``) and the MisScore head of AlnFeature run here; the Raw.bed path is returned
    for the reference's remaining AlnFeature / OutVCF stages."""
    path = localGraph_npz(args)
    if _ranks()[0] == 0 and os.path.exists(path) and os.path.getsize(path) > 0:
        somatic_bed(path, args.savedir, args.TSampleID)
    return path


def main(argv=None):
    parser = argparse.ArgumentParser(prog="svscope_b200.SVscope", description=__doc__)
    sub = parser.add_subparsers(dest="cmd", required=True)
    for name, fn in (("localGraph_npz", localGraph_npz), ("callsomaticSV", callsomaticSV), ("localGraph", localGraph)):
        p = sub.add_parser(name)
        p.add_argument("-s", "--savedir", required=True)
        p.add_argument("-t", "--TSampleID", required=True)
        p.add_argument("-n", "--NSampleID", required=True)
        p.add_argument("-p", "--thread", default="1")
        p.add_argument("-o", "--offset", type=int, default=50)
        p.add_argument("-q", "--mapQ", type=int, default=5)
        p.add_argument("-C", "--Continue", action="store_true")
        p.add_argument("-w", "--windowBed")
        p.add_argument("-T", "--Tumorbam")
        p.add_argument("-N", "--Normalbam")
        p.add_argument("-r", "--Reference")
        p.set_defaults(func=fn)
    args = parser.parse_args(argv)
    return args.func(args)


if __name__ == "__main__":
    main()
