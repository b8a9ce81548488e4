"""BASELINE configs[2] at FULL size through the CPU emulation of the product's alignment code (tests/emul: the shared cell
header poa_cell.h, the pruning bands and the device-graph code, EmuSession(prune=1)) - about 100 minutes on one core.
The digest of all alignment pairs is compared with the one the oracle's row-checkpoint engine produced
(tests/golden/c3_full.json: alignments_sha256), i.e. the product's cell arithmetic, exact pruning and graph code
agree with the oracle at this size without a GPU (log: profiles/r02_emul_c3_full.log).

    python tests/tools/emul_c3_full.py
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from svscope_b200 import synth                   # noqa: E402
from tests.emul.emul import EmuSession, lib      # noqa: E402


def main():
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "c3_full.json")))
    seqs = synth.make_c3(seed=gold["seed"])[0]
    e = EmuSession(ring_rows=8, prune=1)
    t0 = time.time()
    sha = hashlib.sha256()
    for k, s in enumerate(seqs):
        p = e.add(s)
        sha.update(np.ascontiguousarray(p, dtype=np.int32).tobytes())
        print("read", k, "nodes", lib().emu_num_nodes(e.h), "kept %.3f" % e.kept_fraction(), "retries", e.retries(),
              "t %.0f" % (time.time() - t0), flush=True)
    print("alignments_sha256", sha.hexdigest())
    print("oracle           ", gold["alignments_sha256"])
    print("EQUAL" if sha.hexdigest() == gold["alignments_sha256"] else "DIFFERENT")


if __name__ == "__main__":
    main()
