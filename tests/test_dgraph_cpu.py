"""Device-resident graph code (svscope_b200/csrc/poa_dgraph.h) executed on the CPU: merge of an
alignment, rank order, rank-ordered export (predecessors, flags, column-0 scores and codes,
path-length intervals), MSA and consensus must equal the host graph class the oracle tests pin,
after every read, for any number of cooperating threads."""
import numpy as np

from svscope_b200 import synth
from tests.emul import emul
from tests.tools.fuzz_emul import make_group


def test_device_graph_equals_host_graph_on_adversarial_groups():
    rng = np.random.default_rng(2024)
    for it in range(250):
        g = make_group(rng)
        nt, ring = [(1, 4), (128, 2), (7, 12), (256, 10)][it % 4]
        assert emul.dgraph_check(g, ring_rows=ring, n_threads=nt) == "", g


def test_device_graph_equals_host_graph_on_sv_windows():
    for seed, (body, sv, typ) in enumerate([(900, 200, "INS"), (700, 150, "DEL")]):
        w = synth.make_sv_window(seed + 3, body, typ, sv, 8, 8, 4, 0.08)
        assert emul.dgraph_check(list(w[0]), ring_rows=10, n_threads=128) == ""


def test_device_graph_handles_empty_and_single_sequences():
    assert emul.dgraph_check(["", "ACGT", "", "ACGGT"]) == ""
    assert emul.dgraph_check(["ACGT"]) == ""
    assert emul.dgraph_check(["", ""]) == ""
    assert emul.dgraph_check(["A", "C", "G", "T", "A"]) == ""
