"""MisScore of the Raw.bed records on the GPU, with the reference's names
(src/PairwiseCompare.py; SURVEY.md §8f row F1 — the step right after localGraph, called by
``AlnFeature`` at src/SVscope.py:282).

The reference aligns every (somatic consensus, germline consensus) pair of an ``EMOutput``
record with ``Bio.pairwise2.align.globalms(seq1, seq2, 1, 0, -1, -1)[0]`` and takes
``MisScore = len(match line) - match line.count("|")`` (:19-30).  Here all pairs of a file go
to one ``svs_misscore_pairs`` call (one CTA per pair); the per-record reduction
(:54-64) and the allele-frequency string (:66-75) stay on the host as in the reference.
There is no CPU fallback."""
from __future__ import annotations

import argparse
import os
import re

import numpy as np

from . import _lib

COLUMNS = ['chrom', 'start', 'end', 'somSeqList', 'somSupportReadID', 'someventCount', 'germSeqList',
           'germSupportReadID', 'germeventCount', 'flag']
OUT_COLUMNS = ['chrom', 'start', 'end', 'window', 'somSupportReadID', 'germSupportReadID', 'MisScore', 'AF']


def misscore_pairs(pairs, match: int = 1, mismatch: int = 0, open: int = -1, extend: int = -1,
                   want_lines: bool = False, ctx: "_lib.Context | None" = None, stats: dict | None = None):
    """First ``globalms`` alignment of every (seqA, seqB) pair.  Returns an int32 array [n, 4]
    (score, alignment columns, '|' columns, MisScore) and, with ``want_lines``, the match lines."""
    ctx = ctx or _lib.Context.default()
    pairs = list(pairs)
    n = len(pairs)
    out = np.zeros((n, 4), np.int32)
    if n == 0:
        return (out, []) if want_lines else out
    for v in (match, mismatch, open, extend):
        if int(v) != v:
            raise ValueError("svs_misscore_pairs takes integer scores")
    index = {}
    for a, b in pairs:
        for s in (a, b):
            if s not in index:
                index[s] = len(index)
    reads = _lib.ReadSet(ctx, list(index))
    ia = np.array([index[a] for a, _ in pairs], np.int64)
    ib = np.array([index[b] for _, b in pairs], np.int64)
    lines = line_off = None
    if want_lines:
        cap = np.array([len(a) + len(b) for a, b in pairs], np.int64)
        line_off = np.zeros(n + 1, np.int64)
        line_off[1:] = np.cumsum(cap)
        lines = np.zeros(max(int(line_off[-1]), 1), np.uint8)
    st = np.zeros(4, np.float64)
    try:
        ctx.check(_lib.load().svs_misscore_pairs(ctx._h, reads._h, _lib.ptr(ia), _lib.ptr(ib), n, int(match),
                                                 int(mismatch), int(open), int(extend), _lib.ptr(out),
                                                 _lib.ptr(lines), _lib.ptr(line_off), _lib.ptr(st), 4))
    finally:
        reads.close()
    if stats is not None:
        stats.update(cells=float(st[0]), kernel_ms=float(st[1]), launches=int(st[2]), trace_bytes=float(st[3]),
                     pairs=n)
    if want_lines:
        text = [lines[line_off[k]:line_off[k] + out[k, 1]].tobytes().decode() for k in range(n)]
        return out, text
    return out


def AligmentScore(SomConsensus, GerConsensus, cutoff=0):
    """:19-30 (batch of one).  Empty input raises IndexError, as ``globalms(...)[0]`` does."""
    if not SomConsensus or not GerConsensus:
        raise IndexError("list index out of range")
    if cutoff == 0:
        return int(misscore_pairs([(str(SomConsensus), str(GerConsensus))])[0, 3])
    _, lines = misscore_pairs([(str(SomConsensus), str(GerConsensus))], want_lines=True)
    alig = lines[0]
    TD_alig = alig[cutoff:len(alig) - cutoff]
    return len(TD_alig) - TD_alig.count("|")


def smaller_absolute_value(a, b):
    """:32-36 — ``a`` only if strictly smaller in absolute value (ties go to ``b``)."""
    return a if abs(a) < abs(b) else b


def Mismatch_abs(callLine):
    """:38-52.  The reference compares every length difference with a constant that is never
    updated, so each entry ends up as len(somatic consensus) - len(LAST germline consensus);
    that is what is returned here (one number, or ';'-joined for several somatic consensus)."""
    last_germ = callLine['germSeqList'].split(';')[-1]
    diffs = [smaller_absolute_value(10 ** 21, len(som) - len(last_germ)) for som in callLine['somSeqList'].split(';')]
    return ';'.join(str(d) for d in diffs)


def _record_pairs(somSeqList: str, germSeqList: str):
    return [(Som, Ger) for Som in somSeqList.split(';') for Ger in germSeqList.split(';')]


def _reduce_record(pairs, scores):
    """:57-63 on precomputed alignment scores of the record's pairs (same order)."""
    MisScore = 1000000000000000000000
    for (Som, Ger), score in zip(pairs, scores):
        score = int(score)
        if len(Som) < len(Ger):
            score = (-1) * score
        MisScore = smaller_absolute_value(MisScore, score)
    return MisScore


def CalculateMisscore(callLine):
    """:54-64 for one record (mapping with 'somSeqList' and 'germSeqList')."""
    pairs = _record_pairs(callLine['somSeqList'], callLine['germSeqList'])
    for Som, Ger in pairs:
        if not Som or not Ger:
            raise IndexError("list index out of range")
    return _reduce_record(pairs, misscore_pairs(pairs)[:, 3])


def CallAlleleFreq(SomaticTD):
    """:66-75: per somatic cluster, reads / (all somatic reads + germline-cluster reads).  The
    reference filters the germline reads with ``re.search('_tumor|', x)``, whose empty
    alternative matches every ID, so ALL germline-cluster reads count - kept as is."""
    som_counts = np.array([len(ids.split(",")) for ids in SomaticTD['somSupportReadID'].split(";")])
    germ_reads = [rid for ids in SomaticTD['germSupportReadID'].split(";") for rid in ids.split(",")]
    counted = [rid for rid in germ_reads if re.search('_tumor|', rid)]
    total = np.sum(som_counts) + len(counted)
    return ";".join(str(x) for x in som_counts / total)


def MisScorePipe(filepath, stats: dict | None = None):
    """:77-88: Raw.bed -> DataFrame ['chrom','start','end','window','somSupportReadID',
    'germSupportReadID','MisScore','AF'] of the 'NormalOutput|EMOutput' records; all alignments
    of the file in one GPU batch."""
    import pandas as pd
    df = pd.read_csv(filepath, sep="\t", header=None)
    df.columns = COLUMNS
    somDf = df.loc[df['flag'] == 'NormalOutput|EMOutput'].copy()
    SomaticRes = pd.DataFrame(columns=OUT_COLUMNS)
    if somDf.shape[0] > 0:
        somDf['window'] = somDf['chrom'] + '_' + somDf['start'].astype('str') + '-' + somDf['end'].astype('str')
        per_record = [_record_pairs(s, g) for s, g in zip(somDf['somSeqList'], somDf['germSeqList'])]
        flat = [p for rec in per_record for p in rec]
        for Som, Ger in flat:
            if not Som or not Ger:
                raise IndexError("list index out of range")
        scores = misscore_pairs(flat, stats=stats)[:, 3]
        mis, k = [], 0
        for rec in per_record:
            mis.append(_reduce_record(rec, scores[k:k + len(rec)]))
            k += len(rec)
        somDf['MisScore'] = mis
        somDf['AF'] = [CallAlleleFreq(row) for _, row in somDf.iterrows()]
        SomaticRes = somDf[OUT_COLUMNS]
    return SomaticRes


def main(args):
    """:90-97: <workDir>/<sample>/<sample>.vs.<sample>.TandemRepeat.Raw.bed -> <outputDir>/<sample>.Somatic.bed"""
    sample = args.sampleID
    raw_bed = os.path.join(args.workDir, sample, "%s.vs.%s.TandemRepeat.Raw.bed" % (sample, sample))
    os.makedirs(args.outputDir, exist_ok=True)
    MisScorePipe(raw_bed).to_csv(os.path.join(args.outputDir, '%s.Somatic.bed' % sample), sep="\t", header=None,
                                 index=False)


if __name__ == "__main__":
    cli = argparse.ArgumentParser(description=__doc__)
    cli.add_argument("-w", "--workDir", required=True, help="directory holding <sampleID>/...Raw.bed")
    cli.add_argument("-s", "--sampleID", required=True)
    cli.add_argument("-o", "--outputDir", required=True)
    main(cli.parse_args())
