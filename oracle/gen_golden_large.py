"""Full-size golden records made with the CPU oracle (minutes of CPU): configs[0] (C1: 30+30 reads
~10 kb, 2 kb somatic DEL) and a scaled configs[2] (tandem-repeat INS, 10 % error).  Written to
tests/golden/large_*.json: the 10-field record, sha256 of the window MSA and the consensus."""
import hashlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O
from svscope_b200 import synth

def one(name, w):
    t0 = time.time()
    cons, msa = O.poa(w[0], 1)
    rec = O.decision(w[4], w[0], w[1], w[2], w[3])
    out = dict(record=[str(x) for x in rec], consensus_sha=hashlib.sha256(cons.encode()).hexdigest(),
               msa_sha=hashlib.sha256("\n".join(msa).encode()).hexdigest(), msa_cols=len(msa[0]), seconds=time.time() - t0)
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", name + ".json"), "w"))
    print(name, out["record"][5], out["record"][8], out["record"][9], out["msa_cols"], round(out["seconds"], 1), "s", flush=True)

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "c3"):
        one("large_c3_scaled", synth.make_c3(seed=3, total_len=6000, n_tumor=20, n_normal=20, n_carriers=10, err=0.10))
    if which in ("all", "c1"):
        one("large_c1", synth.make_c1(seed=1))
