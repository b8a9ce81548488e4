"""Runs bench.main() with stand-ins for the device stages: control flow, time budget and JSON assembly of
bench.py on a box without a GPU (tests/test_host_logic.py::test_bench_line_contract_with_device_stand_ins).
No timing printed by this script means anything.

    python tests/tools/fake_bench.py <seconds per window> [bench.py arguments]
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
torch.cuda.is_available = lambda: True
torch.cuda.synchronize = lambda *a, **k: None
class Ev:
    def __init__(self, enable_timing=True): self.t=None
    def record(self): self.t=time.perf_counter()
    def synchronize(self): pass
    def elapsed_time(self, o): return (o.t-self.t)*1e3
torch.cuda.Event = Ev
from svscope_b200 import _lib, batch
class Ctx:
    def __init__(self, local): self.o=dict(poa_threads=384, ring_rows=8, sm_count=148)
    def set_option(self,k,v): self.o[k]=v
    def get_option(self,k): return self.o[k]
    def int_alu_probe(self): return dict(add=6e5,max=2.8e4,xor=3.5e4,addmax=1.8e4)
_lib.Context = Ctx
class RS:
    nbytes=123
    def close(self): pass
batch.upload_windows = lambda ctx, w: RS()
SCALE = float(sys.argv[1]) if len(sys.argv)>1 and not sys.argv[1].startswith('-') else 0.002
if len(sys.argv)>1 and not sys.argv[1].startswith('-'): del sys.argv[1]
def fake(wins, ctx=None, reads=None, edit_distance=False, chunks=3, **k):
    time.sleep(SCALE * len(wins) + 0.1)     # + 0.1 s: a step of n windows is not n/N of the full step
    keys=["poa_algo_bytes","poa_alignments","poa_cells","poa_copy_bytes","poa_d2h_bytes","poa_dp_launches","poa_dp_ms","poa_eval_cells","poa_exported_rows","poa_h2d_bytes","poa_rows","ed_cells","ed_ms","windows_em","aux_launches"]+["poa_cyc_"+x for x in ("export","dp","traceback","merge","rank","finish")]
    st={k:float(len(wins))*10 for k in keys}; st["sub_batches"]=3
    return batch.BatchOutput(records=[["c","1","2","-","-",0,"-","-",0,"NormalOutput"] for _ in wins], timings=dict(total=1.0), stats=st)
batch.localgraph_batch = fake
batch.edit_distance_matrices = lambda ctx, reads, groups: ([], dict(cells=1e12, ms=100.0, bytes=0, pairs=len(groups)))
import bench
bench.make_batch = lambda n, r: [([ "ACGT"*10 ]*5, None, None, None, None) for _ in range(n)]
bench.cpu_baseline = lambda *a, **k: dict(value=1.0, note="fake")
class S:
    def __init__(s,d): pass
    def start(s): pass
    def stop(s): return {"sm_mhz":1965.0,"sm_max_mhz":1965.0,"reasons":[],"samples":1}
bench.ClockSampler = S
bench.main()
