"""The product's warp-pipelined DP SOURCE (svscope_b200/csrc/poa_dp2.cuh: bands, exact-size code rows, row loop
with the hand-over between warps, strips, exported rows, warp traceback) executed on the CPU through
tests/emul/cuda_shim.h - one OS thread per warp, one user-level context per lane - against the oracle.
What the GPU tests check on the device is checked here for the control flow of the kernel source itself:
several strips per alignment, warps without columns in a strip, pruning bands and their retry, predecessor
rows that left the on-chip ring (ring depth 1-3), multi-predecessor rows."""
import numpy as np
import pytest

from svscope_b200 import synth


def _check_group(oracle, seqs, **kw):
    from tests.emul.emul import EmuSession
    o, e = oracle.PoaSession(1), EmuSession(**kw)
    try:
        for k, s in enumerate(seqs):
            a, b = o.add(s), e.add(s)
            assert a.shape == b.shape and np.array_equal(a, b), (k, len(s), kw)
        assert e.consensus() == o.consensus()
        return e.warp_retries()
    finally:
        e.close()
        o.close()


@pytest.mark.parametrize("threads,ring,body,sv,nreads,prune", [
    (128, 3, 300, 60, 4, 0),       # one strip, two warps
    (128, 1, 500, 120, 4, 1),      # ring of one row: almost every non-adjacent predecessor comes from global memory
    (128, 2, 1300, 200, 4, 1),     # two balanced strips of 2.6 warps: the fourth warp has no columns
    (128, 3, 2400, 500, 4, 1),     # three strips, bands
    (256, 8, 2600, 700, 3, 1),     # 256 threads, two strips
    (256, 4, 1500, 300, 3, 0),     # unpruned, one strip of six warps
    (384, 8, 3300, 600, 3, 1),     # the product's default shape (twelve warps, ring of 8 rows), two strips, bands
])
def test_dp2_source_on_cpu_equals_oracle(oracle, threads, ring, body, sv, nreads, prune):
    w = synth.make_sv_window(100 + body + threads, body, "DEL" if body % 200 else "INS", sv, nreads, nreads,
                             max(1, nreads // 2), 0.05)
    _check_group(oracle, w[0], ring_rows=ring, warp_threads=threads, warp_prune=prune)


def test_dp2_source_on_cpu_adversarial_groups(oracle):
    """Tandem repeats, large indels, noisy and unrelated reads (the pruning guess fails: retry), 10 % error."""
    rng = np.random.default_rng(5)
    retries = 0
    for it in range(6):
        L = int(rng.integers(700, 1500))
        base = synth._rand_seq(rng, L)
        if it % 2 == 0:
            mot = synth._rand_seq(rng, int(rng.integers(2, 9)))
            base[200:500] = np.tile(mot, 300 // len(mot) + 1)[:300]
        seqs = []
        for r in range(5):
            s = base.copy()
            if rng.random() < 0.6:
                p, ln = int(rng.integers(50, L - 300)), int(rng.integers(20, 250))
                s = np.concatenate([s[:p], s[p + ln:]]) if rng.random() < 0.5 else \
                    np.concatenate([s[:p], synth._rand_seq(rng, ln), s[p:]])
            seqs.append(synth._to_str(synth.noisy_copy(rng, s, float(rng.choice([0.03, 0.1])))))
        if it == 3:
            seqs[3] = synth._to_str(synth._rand_seq(rng, 900))       # unrelated read
        retries += _check_group(oracle, seqs, ring_rows=int(rng.integers(1, 4)), warp_threads=128, warp_prune=1)
    assert retries >= 1      # the retry path ran


def test_dp2_source_on_cpu_no_cyclic_wait_after_an_exported_row(golden_dir):
    """Regression (found by tests/tools/fuzz_dp2_cpu.py, seed 11): warp 3 asked for the fence of an exported row of
    warp 2 (dp2_stage_global) while warp 2, whose next row with cells came 40 pruned rows later, waited for warp 3
    to come within 32 rows - fences were only issued every 8 visited rows.  The end-of-batch fence in poa_dp2.cuh
    breaks the cycle.  Run in a child process: a cyclic wait would otherwise hang the suite."""
    import json
    import os
    import subprocess
    import sys
    case = os.path.join(golden_dir, "dp2_deadlock_case.json")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import json, sys, numpy as np\n"
        "sys.path.insert(0, %r)\n"
        "from oracle import oracle as O\n"
        "from tests.emul.emul import EmuSession\n"
        "d = json.load(open(%r))\n"
        "for kw in (d['kw'], dict(d['kw'], warp_threads=256), dict(d['kw'], ring_rows=1)):\n"
        "    o, e = O.PoaSession(1), EmuSession(**kw)\n"
        "    for s in d['seqs']:\n"
        "        assert np.array_equal(o.add(s), e.add(s))\n"
        "print('ok')\n" % (root, case))
    env = dict(os.environ, SVS_EMU_WATCHDOG="90")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=400, env=env)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), (out.returncode, out.stderr[-1500:])
