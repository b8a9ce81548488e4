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

BATCH_WINDOWS = 512


def raw_bed_name(TSampleID: str, NSampleID: str) -> str:
    return "%s.vs.%s.TandemRepeat.Raw.bed" % ("-".join(TSampleID.split(",")), "-".join(NSampleID.split(",")))


def write_raw_bed(path: str, records, append: bool = False) -> None:
    with open(path, "a" if append else "w") as f:
        for rec in records:
            f.write("\t".join([str(x) for x in rec]) + "\n")
    os.system("sort -k1,1 -k2,2n {p} -o {p}".format(p=path))


def run_windows(windows, batch_windows: int = BATCH_WINDOWS, **kw):
    """Records for a list of windows, sharded over the ranks of the current process group."""
    import torch.distributed as dist
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    if distributed:
        mine = _shard.my_shard([window_cost(w) for w in windows], dist.get_rank(), dist.get_world_size())
    else:
        mine = list(range(len(windows)))
    local = []
    for b in range(0, len(mine), batch_windows):
        chunk = [windows[i] for i in mine[b:b + batch_windows]]
        local.extend(localgraph_batch(chunk, **kw).records)
    if not distributed:
        return local
    return _shard.gather_records(mine, local, len(windows), group=_shard.host_group())


def localGraph_npz(args):
    """Reference :185-239: all ``*.npz`` batches in ``args.savedir`` -> Raw.bed; ``--Continue``
    skips windows whose TDRecord is already in the output."""
    logging.info("Local Graph : Start working")
    t0 = time.time()
    path = os.path.join(args.savedir, raw_bed_name(args.TSampleID, args.NSampleID))
    finished = set()
    if getattr(args, "Continue", False) and os.path.exists(path):
        with open(path) as fh:
            finished = {"\t".join(x.strip().split("\t")[0:3]) for x in fh.readlines()}
    windows = []
    for name in sorted(os.listdir(args.savedir)):
        if not re.search("npz", name):
            continue
        dat = np.load(os.path.join(args.savedir, name), allow_pickle=True)["DatSet"]
        for i in range(dat.shape[0]):
            row = list(dat[i])
            if finished and row[4] in finished:
                continue
            windows.append(row)
    records = run_windows(windows)
    if records is not None:
        write_raw_bed(path, records, append=bool(finished))
    logging.info(f"Local Graph : work finished with {(time.time() - t0) / 3600} hour")
    return path


def localGraph(args):
    """Reference :118-183 extracts every window from the BAM files with pysam before the
    decision.  That feeder is outside this package; dump the windows with the reference's
    SomTDDetector_AimDatFetch.py (npz rows) and run ``localGraph_npz``."""
    raise NotImplementedError(
        "localGraph needs the pysam/BAM feeder (src/DataScanner.py:222-247), which is out of scope of the "
        "accelerated path; use localGraph_npz on npz batches written by SomTDDetector_AimDatFetch.py")


def callsomaticSV(args):
    """Reference :341-356 = localGraph + AlnFeature.  Only the localGraph half is accelerated;
    when ``args.savedir`` holds npz batches it is run here and the Raw.bed path is returned for
    the reference's unchanged AlnFeature / OutVCF stages."""
    return localGraph_npz(args)


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
