"""``Decision`` with the reference's signature (src/DecisionMaker.py:110): one window in, the
10-field record out.  It is the batch pipeline with a batch of one."""
from __future__ import annotations

import numpy as np

from .batch import localgraph_batch


def Decision(TDRecord, sequenceList, ReadIDs, flank_5, flank_3, windowFlag="NormalOutput", Tlabel="tumor",
             readcutoff=3, hcutoff=3, scutoff=0.05):
    out = localgraph_batch([[list(sequenceList), np.asarray(ReadIDs), flank_5, flank_3, TDRecord]],
                           windowFlags=[windowFlag], Tlabel=Tlabel, readcutoff=readcutoff, hcutoff=hcutoff,
                           scutoff=scutoff, reseed=False)
    return out.records[0]
