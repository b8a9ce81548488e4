import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def c3_full_gpu():
    """BASELINE configs[2] at FULL size (synth.make_c3(seed=3): 120 reads of 20 kb, 10 % error) aligned once on
    the device with the default CTA shape; shared by the property test (tests/test_gpu_poa.py) and the golden
    test (tests/test_zz_gpu_full_c3_golden.py)."""
    if os.environ.get("SVS_SKIP_FULL_C3"):
        pytest.skip("SVS_SKIP_FULL_C3 set")
    from svscope_b200 import _lib, synth
    from svscope_b200._lib import ReadSet
    from svscope_b200.poa_api import poa_groups
    ctx = _lib.Context.default(0)
    seqs = synth.make_c3(seed=3)[0]
    reads = ReadSet(ctx, seqs)
    try:
        ctx.set_option("poa_threads", 384)
        cons, msas, st = poa_groups(ctx, reads, [list(range(len(seqs)))])
    finally:
        reads.close()
    return dict(seqs=seqs, cons=cons, msas=msas, st=st)
