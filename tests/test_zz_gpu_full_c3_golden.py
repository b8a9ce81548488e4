"""BASELINE configs[2] at FULL size against the CPU oracle: 120 reads of 20 kb at 10 % error with a tandem-repeat
expansion, aligned on the device (memory tiers, deep graph), compared with the digests the oracle's
row-checkpoint engine produced (oracle/gen_golden_c3.py -> tests/golden/c3_full.json; the engine itself is
checked against the flat five-matrix engine in tests/test_oracle_golden.py).  Collected last on purpose
(`pytest -x`): it re-uses the alignment of tests/test_gpu_poa.py's property test."""
import hashlib
import json
import os

import pytest

pytestmark = pytest.mark.gpu


def test_full_size_configs2_window_equals_oracle_digests(c3_full_gpu, golden_dir):
    path = os.path.join(golden_dir, "c3_full.json")
    if not os.path.exists(path):
        pytest.skip("tests/golden/c3_full.json not generated (oracle/gen_golden_c3.py, about an hour of CPU)")
    g = json.load(open(path))
    seqs, cons, msas, st = (c3_full_gpu[k] for k in ("seqs", "cons", "msas", "st"))
    assert hashlib.sha256("\n".join(seqs).encode()).hexdigest() == g["input_sha256"]
    assert st["failed_groups"] == 0
    msa = msas[0]
    assert len(msa[0]) == g["msa_cols"]
    assert hashlib.sha256("\n".join(msa).encode()).hexdigest() == g["msa_sha256"]          # bit-exact MSA
    assert len(cons[0]) == g["consensus_len"]
    assert hashlib.sha256(cons[0].encode()).hexdigest() == g["consensus_sha256"]           # bit-exact consensus
