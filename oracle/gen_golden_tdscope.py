"""Generate tests/golden/tdscope_cases.json: the REFERENCE's ``SomTDDetector.TDscope``
(src/SomTDDetector.py:26-61, imported unmodified; stubs as in oracle/gen_golden.py) driven with
synthetic ``DataMaker`` / ``DataMaker2`` callables and ``DecisionMaker = oracle.decision``, over the
branches of its DUP rescue.  The test replays the same scenarios through
``svscope_b200.SomTDDetector.TDscope``.

    python oracle/gen_golden_tdscope.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from gen_golden import import_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "tdscope_cases.json")


def scenarios():
    """name -> (TDRecord, first window spec, rescan specs).  A spec is (seed, n_carriers,
    extra tumor read ids, flag); n_carriers = 0 gives a window without a somatic cluster."""
    rec_dup = "chr1\t5000\t5300\tDUP,x\t7"
    rec_del = "chr1\t5000\t5300\tDEL\t7"
    em, no = 4, 0
    return {
        "del_em": (rec_del, (11, em, 0, "NormalOutput"), None),
        "del_no_em_no_rescue": (rec_del, (12, no, 0, "NormalOutput"), ((13, em, 0, "UnspanedSV"), (14, em, 0, "UnspannedSV"))),
        "dup_em_first": (rec_dup, (15, em, 0, "NormalOutput"), ((16, em, 0, "UnspanedSV"), (17, em, 0, "UnspannedSV"))),
        "dup_rescue_5": (rec_dup, (18, no, 0, "NormalOutput"), ((19, em, 0, "UnspanedSV"), (20, em, 0, "UnspannedSV"))),
        "dup_rescue_3": (rec_dup, (21, no, 0, "NormalOutput"), ((22, no, 0, "UnspanedSV"), (23, em, 0, "UnspannedSV"))),
        "dup_flag_5": (rec_dup, (24, no, 0, "NormalOutput"), ((25, no, 3, "UnspanedSV"), (26, no, 3, "UnspannedSV"))),
        "dup_flag_3": (rec_dup, (27, no, 0, "NormalOutput"), ((28, no, 2, "UnspanedSV"), (29, no, 4, "UnspannedSV"))),
        "dup_nothing": (rec_dup, (30, no, 0, "NormalOutput"), ((31, no, 1, "UnspanedSV"), (32, no, 2, "UnspannedSV"))),
    }


def build_window(spec, TDRecord, base_ids=None):
    from svscope_b200 import synth
    seed, n_carriers, extra, flag = spec
    w = synth.make_small_window(seed, body_len=260, sv_len=70, n_tumor=7, n_normal=7, n_carriers=n_carriers)
    ids = list(w[1])
    if base_ids is not None:          # rescans share the read names of the first extraction ...
        ids = list(base_ids)
        for k in range(extra):        # ... except `extra` tumor reads that only the rescan sees
            ids[k] = ids[k].replace("|", "|new%d_" % k)
    return [w[0], np.array(ids), w[2], w[3], TDRecord, flag]


def makers(name):
    TDRecord, first, rescans = scenarios()[name]
    w0 = build_window(first, TDRecord)

    def DataMaker(rec):
        return tuple(w0)

    def DataMaker2(rec):
        return [tuple(build_window(s, TDRecord, base_ids=w0[1])) for s in rescans]

    return TDRecord, DataMaker, DataMaker2


def main():
    import_reference()
    import SomTDDetector as REF
    from oracle import oracle as O
    out = {}
    for name in scenarios():
        TDRecord, dm, dm2 = makers(name)
        np.random.seed(2023)
        rec = REF.TDscope(TDRecord, dm, dm2, O.decision)
        out[name] = [str(x) for x in rec]
        print(name, rec[-1], rec[5], rec[8])
    json.dump({"generator": "oracle/gen_golden_tdscope.py", "records": out}, open(OUT, "w"), indent=1)


if __name__ == "__main__":
    main()
