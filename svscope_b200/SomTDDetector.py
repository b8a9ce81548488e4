"""Per-window drivers with the reference's signatures (src/SomTDDetector.py:26-73)."""
from __future__ import annotations

import logging
import re
import time

import numpy as np

from .DecisionMaker import Decision

logging.basicConfig(level=logging.INFO, format="%(asctime)s - %(levelname)s - %(message)s")


def TDscope_npz(TDRecord, sequenceList, ReadIDs, flank_5, flank_3):
    """npz variant: no BAM access, default window flag (reference :63-73)."""
    t0 = time.time()
    record = Decision(TDRecord, sequenceList, ReadIDs, flank_5, flank_3)
    logging.info(f"pipeline for region {TDRecord} finished Take {time.time() - t0}s")
    return record


def TDscope(TDRecord, DataMaker, DataMaker2, DecisionMaker):
    """Extraction + decision + DUP rescue (reference :26-61).  ``DataMaker``/``DataMaker2`` are
    supplied by the caller (the pysam feeder is not part of this package)."""
    sequenceList, ReadIDs, flank_5, flank_3, TDRecord, flag = DataMaker(TDRecord)
    sv_type = TDRecord.strip().split("\t")[3].split(",")[0]
    record = DecisionMaker(TDRecord, sequenceList, ReadIDs, flank_5, flank_3, flag)
    if record[-1].split("|")[-1] != "EMOutput" and sv_type == "DUP":
        rescan = DataMaker2(TDRecord)
        seq5, ids5, f55, f35, TDRecord, flag5 = rescan[0]
        seq3, ids3, f53, f33, TDRecord, flag3 = rescan[1]
        rec5 = DecisionMaker(TDRecord, seq5, ids5, f55, f35, flag5)
        if rec5[-1].split("|")[-1] == "EMOutput":
            return rec5
        rec3 = DecisionMaker(TDRecord, seq3, ids3, f53, f33, flag3)
        if rec3[-1].split("|")[-1] == "EMOutput":
            return rec3
        if len([x for x in np.setdiff1d(ids5, ReadIDs) if re.search("_tumor", x)]) >= 3:
            record[-1] = flag5
        elif len([x for x in np.setdiff1d(ids3, ReadIDs) if re.search("_tumor", x)]) >= 3:
            record[-1] = flag3
    return record
