"""ORACLE — TEST INFRASTRUCTURE ONLY.  **Parity unpinned.**

Pure-Python restatement of the part of Biopython's ``Bio.pairwise2`` that the reference's
MisScore step runs:

    src/PairwiseCompare.py:19-30   AligmentScore:
        alignment = pairwise2.align.globalms(seq1, seq2, 1, 0, -1, -1)[0]
        alig = format_alignment(*alignment).split('\\n')[1];  MisScore = len(alig) - alig.count("|")

Biopython is a third-party dependency of the reference (``from Bio import pairwise2``,
PairwiseCompare.py:8-11; version not pinned by README.md) and is absent from this image and
from /root/reference, and the reference holds no vectors for this step.  The functions below
restate the published algorithm of ``pairwise2`` (Gotoh fill with the five-bit trace encoding
``_make_score_matrix_fast``, the stack-driven ``_recover_alignments`` with its dead-end rule
and ``_find_gap_open``, ``_finish_backtrace``, ``_clean_alignments``, ``format_alignment``) from
recollection, literally and with strings, so it is slow: small cases only.  The C version
(oracle/misscore_oracle.c) is checked against this one in tests/test_oracle_golden.py.

MisScore depends on WHICH co-optimal alignment comes first (for match 1 / mismatch 0 / gap -1,
MisScore = (lenA + lenB)/2 - score - gaps/2), so the traversal order below is the contract.
"""
from __future__ import annotations

MAX_ALIGNMENTS = 1000
_PRECISION = 1000


def rint(x, precision=_PRECISION):
    return int(x * precision + 0.5)


def calc_affine_penalty(length, open, extend, penalize_extend_when_opening):
    if length <= 0:
        return 0
    penalty = open + extend * length
    if not penalize_extend_when_opening:
        penalty -= extend
    return penalty


def _make_score_matrix_fast(sequenceA, sequenceB, match, mismatch, open_A, extend_A, open_B, extend_B,
                            penalize_extend_when_opening=False, penalize_end_gaps=(True, True)):
    first_A_gap = calc_affine_penalty(1, open_A, extend_A, penalize_extend_when_opening)
    first_B_gap = calc_affine_penalty(1, open_B, extend_B, penalize_extend_when_opening)
    lenA, lenB = len(sequenceA), len(sequenceB)
    score_matrix, trace_matrix = [], []
    for i in range(lenA + 1):
        score_matrix.append([None] * (lenB + 1))
        trace_matrix.append([None] * (lenB + 1))
    for i in range(lenA + 1):
        score_matrix[i][0] = calc_affine_penalty(i, open_B, extend_B, penalize_extend_when_opening) \
            if penalize_end_gaps[1] else 0
    for i in range(lenB + 1):
        score_matrix[0][i] = calc_affine_penalty(i, open_A, extend_A, penalize_extend_when_opening) \
            if penalize_end_gaps[0] else 0
    col_score = [0]
    for i in range(1, lenB + 1):
        col_score.append(calc_affine_penalty(i, 2 * open_B, extend_B, penalize_extend_when_opening))
    for row in range(1, lenA + 1):
        row_score = calc_affine_penalty(row, 2 * open_A, extend_A, penalize_extend_when_opening)
        for col in range(1, lenB + 1):
            nogap_score = score_matrix[row - 1][col - 1] + \
                (match if sequenceA[row - 1] == sequenceB[col - 1] else mismatch)
            if not penalize_end_gaps[0] and row == lenA:
                row_open = score_matrix[row][col - 1]
                row_extend = row_score
            else:
                row_open = score_matrix[row][col - 1] + first_A_gap
                row_extend = row_score + extend_A
            row_score = max(row_open, row_extend)
            if not penalize_end_gaps[1] and col == lenB:
                col_open = score_matrix[row - 1][col]
                col_extend = col_score[col]
            else:
                col_open = score_matrix[row - 1][col] + first_B_gap
                col_extend = col_score[col] + extend_B
            col_score[col] = max(col_open, col_extend)
            best_score = max(nogap_score, col_score[col], row_score)
            score_matrix[row][col] = best_score
            # 1 = open gap in seqA, 2 = match/mismatch, 4 = open gap in seqB,
            # 8 = extend gap in seqA, 16 = extend gap in seqB
            row_score_rint = rint(row_score)
            col_score_rint = rint(col_score[col])
            row_trace_score = 0
            col_trace_score = 0
            if rint(row_open) == row_score_rint:
                row_trace_score += 1
            if rint(row_extend) == row_score_rint:
                row_trace_score += 8
            if rint(col_open) == col_score_rint:
                col_trace_score += 4
            if rint(col_extend) == col_score_rint:
                col_trace_score += 16
            trace_score = 0
            best_score_rint = rint(best_score)
            if rint(nogap_score) == best_score_rint:
                trace_score += 2
            if row_score_rint == best_score_rint:
                trace_score += row_trace_score
            if col_score_rint == best_score_rint:
                trace_score += col_trace_score
            trace_matrix[row][col] = trace_score
    return score_matrix, trace_matrix, score_matrix[lenA][lenB]


def _finish_backtrace(sequenceA, sequenceB, ali_seqA, ali_seqB, row, col, gap_char):
    if row:
        ali_seqA += sequenceA[row - 1::-1]
    if col:
        ali_seqB += sequenceB[col - 1::-1]
    if row > col:
        ali_seqB += gap_char * (len(ali_seqA) - len(ali_seqB))
    elif col > row:
        ali_seqA += gap_char * (len(ali_seqB) - len(ali_seqA))
    return ali_seqA, ali_seqB


def _find_gap_open(sequenceA, sequenceB, ali_seqA, ali_seqB, end, row, col, col_gap, gap_char, score_matrix,
                   trace_matrix, in_process, gap_fn, target, index, direction):
    dead_end = False
    target_score = score_matrix[row][col]
    for n in range(target):
        if direction == "col":
            col -= 1
            ali_seqA += gap_char
            ali_seqB += sequenceB[col:col + 1]
        else:
            row -= 1
            ali_seqA += sequenceA[row:row + 1]
            ali_seqB += gap_char
        actual_score = score_matrix[row][col] + gap_fn(index, n + 1)
        if rint(actual_score) == rint(target_score) and n > 0:
            if not trace_matrix[row][col]:
                break
            else:
                in_process.append((ali_seqA[:], ali_seqB[:], end, row, col, col_gap, trace_matrix[row][col]))
        if not trace_matrix[row][col]:
            dead_end = True
    return ali_seqA, ali_seqB, row, col, in_process, dead_end


def _recover_alignments(sequenceA, sequenceB, score, score_matrix, trace_matrix, gap_char, gap_A_fn, gap_B_fn,
                        max_alignments=MAX_ALIGNMENTS):
    lenA, lenB = len(sequenceA), len(sequenceB)
    tracebacks = []
    in_process = []
    row, col = lenA, lenB  # global alignment, end gaps penalised: the only start is the corner
    end = None
    ali_seqA, ali_seqB = sequenceA[0:0], sequenceB[0:0]
    in_process += [(ali_seqA, ali_seqB, end, row, col, False, trace_matrix[row][col])]
    while in_process and len(tracebacks) < max_alignments:
        dead_end = False
        ali_seqA, ali_seqB, end, row, col, col_gap, trace = in_process.pop()
        while (row > 0 or col > 0) and not dead_end:
            cache = (ali_seqA[:], ali_seqB[:], end, row, col, col_gap)
            if not trace:
                if col and col_gap:
                    dead_end = True
                else:
                    ali_seqA, ali_seqB = _finish_backtrace(sequenceA, sequenceB, ali_seqA, ali_seqB, row, col,
                                                           gap_char)
                break
            elif trace % 2 == 1:  # open gap in seqA
                trace -= 1
                if col_gap:
                    dead_end = True
                else:
                    col -= 1
                    ali_seqA += gap_char
                    ali_seqB += sequenceB[col:col + 1]
                    col_gap = False
            elif trace % 4 == 2:  # match/mismatch
                trace -= 2
                row -= 1
                col -= 1
                ali_seqA += sequenceA[row:row + 1]
                ali_seqB += sequenceB[col:col + 1]
                col_gap = False
            elif trace % 8 == 4:  # open gap in seqB
                trace -= 4
                row -= 1
                ali_seqA += sequenceA[row:row + 1]
                ali_seqB += gap_char
                col_gap = True
            elif trace in (8, 24):  # extend gap in seqA
                trace -= 8
                if col_gap:
                    dead_end = True
                else:
                    col_gap = False
                    x = _find_gap_open(sequenceA, sequenceB, ali_seqA, ali_seqB, end, row, col, col_gap, gap_char,
                                       score_matrix, trace_matrix, in_process, gap_A_fn, col, row, "col")
                    ali_seqA, ali_seqB, row, col, in_process, dead_end = x
            elif trace == 16:  # extend gap in seqB
                trace -= 16
                col_gap = True
                x = _find_gap_open(sequenceA, sequenceB, ali_seqA, ali_seqB, end, row, col, col_gap, gap_char,
                                   score_matrix, trace_matrix, in_process, gap_B_fn, row, col, "row")
                ali_seqA, ali_seqB, row, col, in_process, dead_end = x
            if trace:  # another path to follow
                cache += (trace,)
                in_process.append(cache)
            trace = trace_matrix[row][col]
        if not dead_end:
            tracebacks.append((ali_seqA[::-1], ali_seqB[::-1], score, 0, end))
    # _clean_alignments: drop duplicates keeping the order, set `end`
    unique = []
    for align in tracebacks:
        if align not in unique:
            unique.append(align)
    out = []
    for seqA, seqB, sc, begin, end in unique:
        end = len(seqA) if end is None else end
        if begin >= end:
            continue
        out.append((seqA, seqB, sc, begin, end))
    return out


def globalms(sequenceA, sequenceB, match, mismatch, open, extend, max_alignments=MAX_ALIGNMENTS):
    """``pairwise2.align.globalms(seqA, seqB, match, mismatch, open, extend)``: list of
    (seqA, seqB, score, begin, end); the reference takes element 0."""
    if not sequenceA or not sequenceB:
        return []
    if open > 0 or extend > 0:
        raise ValueError("Gap penalties should be non-positive.")
    score_matrix, trace_matrix, best = _make_score_matrix_fast(sequenceA, sequenceB, match, mismatch,
                                                               open, extend, open, extend)

    def gap_fn(index, length):
        return calc_affine_penalty(length, open, extend, False)

    return _recover_alignments(sequenceA, sequenceB, best, score_matrix, trace_matrix, "-", gap_fn, gap_fn,
                               max_alignments)


def match_line(align1, align2):
    """Second line of ``format_alignment`` for a global alignment of plain strings."""
    out = []
    for a, b in zip(align1, align2):
        if a == b:
            out.append("|")
        elif a.strip() == "-" or b.strip() == "-":
            out.append(" ")
        else:
            out.append(".")
    return "".join(out)


def aligment_score(SomConsensus, GerConsensus, cutoff=0, max_alignments=1):
    """PairwiseCompare.py:19-30.  ``max_alignments=1`` stops after the first traceback, which is
    element 0 of the full list (the list is built in discovery order)."""
    alignment = globalms(SomConsensus, GerConsensus, 1, 0, -1, -1, max_alignments=max_alignments)[0]
    alig = match_line(alignment[0], alignment[1])
    TD_alig = alig[cutoff:len(alig) - cutoff]
    return len(TD_alig) - TD_alig.count("|")
