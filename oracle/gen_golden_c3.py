"""Golden digests for BASELINE configs[2] at FULL size (svscope_b200.synth.make_c3: 120 reads of 20 kb,
10 % error, tandem-repeat expansion) made by the CPU oracle's row-checkpoint engine (spoa_oracle.cpp:
BlockedEngine - the flat engine would need five ~150k x 20k int32 matrices).  About an hour on one core.

    python oracle/gen_golden_c3.py [seed]      ->  tests/golden/c3_full.json
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from oracle import oracle as O
    from svscope_b200 import synth
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    w = synth.make_c3(seed=seed)
    seqs = w[0]
    s = O.PoaSession(1, block_rows=-1)
    t0 = time.time()
    scores, kept, recomputed, aln_sha = [], 0, 0, hashlib.sha256()
    for k, x in enumerate(seqs):
        pairs = s.add(x)
        aln_sha.update(np.ascontiguousarray(pairs, dtype=np.int32).tobytes())
        st = s.blocked_stats
        kept = max(kept, st["kept_rows"])
        recomputed += st["recomputed_blocks"]
        scores.append(int(s.score) if len(pairs) else None)
        g_nodes = int(O.lib().spo_num_nodes(s._h))
        print(f"read {k}/{len(seqs)} len {len(x)} nodes {g_nodes} kept_rows {st['kept_rows']} "
              f"recomputed {st['recomputed_blocks']} cells {s.cells:.3e} t {time.time() - t0:.0f} s", flush=True)
    cons = s.consensus()
    msa = s.msa()
    out = dict(generator="oracle/gen_golden_c3.py", engine="row-checkpoint (BlockedEngine, ~2 GB blocks)", seed=seed,
               numpy=np.__version__, n_seqs=len(seqs), ref_len=len(seqs[0]),
               input_sha256=hashlib.sha256("\n".join(seqs).encode()).hexdigest(),
               nodes=int(O.lib().spo_num_nodes(s._h)), edges=int(O.lib().spo_num_edges(s._h)), cells=int(s.cells),
               scores=scores, alignments_sha256=aln_sha.hexdigest(),
               msa_cols=len(msa[0]), msa_sha256=hashlib.sha256("\n".join(msa).encode()).hexdigest(),
               consensus_len=len(cons), consensus_sha256=hashlib.sha256(cons.encode()).hexdigest(),
               max_kept_rows=kept, recomputed_blocks=recomputed, seconds=round(time.time() - t0, 1))
    s.close()
    with open(os.path.join(ROOT, "tests", "golden", "c3_full.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote tests/golden/c3_full.json:", out["nodes"], "nodes,", out["msa_cols"], "columns,", out["seconds"], "s")


if __name__ == "__main__":
    main()
