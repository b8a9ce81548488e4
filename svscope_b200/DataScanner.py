"""Host-side mirror of the MSA / feature-selection part of src/DataScanner.py (lines 124-220).

Same names and argument meaning as the reference so that its callers (and tests written
against it) keep working; the column statistics run on the device.  The BAM extraction half
of the reference file (pysam: FetchTDsubSeq, DataMaker, DataMaker2, :50-122, :222-325) is the
feeder of the hot path and is out of scope here."""
from __future__ import annotations

import numpy as np

from . import batch as _batch
from ._lib import Context
from .spoa import poa


def SeqEncoder(seqinput):
    """A,T,C,G,- -> 0,1,2,3,4 (reference :124-129); KeyError on any other symbol."""
    if len(seqinput) == 0:
        return np.array([])
    return _batch.encode_msa([seqinput if isinstance(seqinput, str) else "".join(seqinput)])[0].astype(np.int64)


def SeqDecoder(seqinput):
    """Drop gap code 4, map 0..3 back to A,T,C,G (reference :131-137)."""
    row = np.asarray(seqinput).astype(np.int64)
    return _batch.DEC_LUT[row[row != 4]].tobytes().decode()


def SeqAligner(seqList):
    """Reference :139-144 calls ``poa(seqList)`` with the default local mode, which is off the
    localGraph path and not implemented on the device."""
    raise NotImplementedError("SeqAligner (local-alignment POA) is not on the localGraph hot path")


def CallMargin(msa, flank_5, flank_3):
    """Columns of the reference row that spell the 5' and 3' flanks (reference :146-165)."""
    return _batch.margin_columns("".join(msa[0]) if not isinstance(msa[0], str) else msa[0], flank_5, flank_3)


def FindNonSameSite(seqencode_New_Sub, cutoff=3):
    """Columns whose second-largest symbol count is >= cutoff (reference :167-179)."""
    mat = np.asarray(seqencode_New_Sub)
    if mat.ndim != 2 or mat.shape[1] == 0:
        return np.zeros(0, np.int64)
    ctx = Context.default()
    keep, _, _, _ = _batch.msa_features(ctx, [mat], [np.zeros(mat.shape[1], np.uint8)], [float(cutoff)])[0]
    return np.where(keep)[0]


def MSAFeatureSelection(sequenceList, flank_5, flank_3, readIDList, hcutoff=3, scutoff=0.05):
    """MSA of reference + reads, encoded matrix, feature matrix, read ids (reference :181-220,
    including its handling of fully deleted reads :198-209)."""
    ctx = Context.default()
    readIDList = np.asarray(readIDList)
    lens = np.array([len(x) for x in sequenceList[1:]])
    _, msa = poa(sequenceList, 1)
    enc = _batch.encode_msa(msa).astype(np.int64)
    nonempty = np.flatnonzero(lens != 0)
    if nonempty.size != lens.size:
        kept = list(readIDList[nonempty])
        enc = np.concatenate([enc, np.full((len(kept), enc.shape[1]), 4, enc.dtype)], axis=0)
        readIDList = np.array(kept + kept)
    drop = np.zeros(enc.shape[1], np.uint8)
    drop[_batch.margin_columns(msa[0], flank_5, flank_3)] = 1
    cutoff = float(max([hcutoff, enc.shape[0] * scutoff]))
    keep, _, _, _ = _batch.msa_features(ctx, [enc[1:]], [drop], [cutoff])[0]
    return enc, enc[1:][:, keep], readIDList


def DataMaker(*args, **kwargs):
    raise NotImplementedError(
        "DataMaker reads BAM/FASTA through pysam (reference src/DataScanner.py:222-247); the extraction "
        "stage is outside the accelerated path - feed windows in the npz row format (localGraph_npz)")


DataMaker2 = DataMaker
