"""ctypes wrapper for the CPU emulation of the GPU alignment path (test infrastructure)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(HERE, "_build", "libpoa_emul.so")
_lib = None


def build(force=False):
    srcs = [os.path.join(HERE, "poa_emul.cpp"), os.path.join(HERE, "dgraph_emul.cpp"),
            os.path.join(HERE, "poa_graph.cpp"), os.path.join(HERE, "dp2_threads.cpp"), os.path.join(HERE, "cuda_shim.cpp")]
    deps = srcs + [os.path.join(HERE, "poa_graph.h"), os.path.join(HERE, "cuda_shim.h")] + [
        os.path.join(ROOT, "svscope_b200", "csrc", h) for h in ("poa_cell.h", "poa_dgraph.h", "poa_task.h", "poa_dp2.cuh")]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-w", "-pthread"] + srcs + ["-o", LIB], check=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        L.emu_new.restype = ctypes.c_void_p
        L.emu_new.argtypes = [ctypes.c_int]
        L.emu_free.argtypes = [ctypes.c_void_p]
        L.emu_add.restype = ctypes.c_int64
        L.emu_add.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int64]
        L.emu_last_alignment.restype = ctypes.c_int64
        L.emu_last_alignment.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
        L.emu_num_nodes.restype = ctypes.c_int64
        L.emu_num_nodes.argtypes = [ctypes.c_void_p]
        L.emu_rank_to_node.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.emu_consensus.restype = ctypes.c_int64
        L.emu_consensus.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
        L.emu_msa_dims.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.emu_msa.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.emu_set_prune.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.emu_set_warp.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        L.emu_warp_retries.restype = ctypes.c_int
        L.emu_warp_retries.argtypes = [ctypes.c_void_p]
        L.emu_set_dyn.argtypes = [ctypes.c_void_p, ctypes.c_double]
        L.emu_set_dyn_ext.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.emu_static_fraction.restype = ctypes.c_double
        L.emu_static_fraction.argtypes = [ctypes.c_void_p]
        L.emu_overflows.restype = ctypes.c_int
        L.emu_overflows.argtypes = [ctypes.c_void_p]
        L.emu_retries.restype = ctypes.c_int
        L.emu_retries.argtypes = [ctypes.c_void_p]
        L.emu_kept_fraction.restype = ctypes.c_double
        L.emu_kept_fraction.argtypes = [ctypes.c_void_p]
        L.emu_cell_selfcheck.restype = ctypes.c_int64
        L.emu_cell_selfcheck.argtypes = [ctypes.c_uint32, ctypes.c_int64, ctypes.c_void_p]
        L.dgraph_emul_check.restype = ctypes.c_int
        L.dgraph_emul_check.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_char_p, ctypes.c_int]
        _lib = L
    return _lib


def dgraph_check(seqs, ring_rows=4, n_threads=128, serial_rank=False):
    """Runs the device-resident graph code (poa_dgraph.h) on the CPU over the sequence group and
    compares every array with the host graph after every read.  Returns '' or the first difference."""
    enc = [s.encode() for s in seqs]
    off = np.zeros(len(enc) + 1, np.int64)
    off[1:] = np.cumsum([len(b) for b in enc])
    msg = ctypes.create_string_buffer(512)
    rc = lib().dgraph_emul_check(b"".join(enc), off.ctypes.data, len(enc), ring_rows, n_threads,
                                 1 if serial_rank else 0, msg, 512)
    return "" if rc == 0 else (msg.value.decode() or "mismatch")


def cell_selfcheck(seed=1, n_random=2000):
    """(mismatches, comparisons) of the shared cell header's self-check (poa_emul.cpp emu_cell_selfcheck)."""
    n = ctypes.c_int64()
    bad = lib().emu_cell_selfcheck(seed, n_random, ctypes.byref(n))
    return int(bad), int(n.value)


class EmuSession:
    def __init__(self, ring_rows=4, prune=0, dyn=None, dyn_ext=None, warp_threads=0, warp_prune=0):
        self.h = lib().emu_new(ring_rows)
        if warp_threads:
            # the product's warp-pipelined DP source on that many OS threads (dp2_threads.cpp)
            lib().emu_set_warp(self.h, int(warp_threads), int(warp_prune))
        if dyn_ext is not None:
            lib().emu_set_dyn_ext(self.h, int(dyn_ext))
        if prune:
            lib().emu_set_prune(self.h, prune)
        if dyn is not None:
            lib().emu_set_dyn(self.h, float(dyn))

    def retries(self):
        return lib().emu_retries(self.h)

    def warp_retries(self):
        return lib().emu_warp_retries(self.h)

    def kept_fraction(self):
        return lib().emu_kept_fraction(self.h)

    def add(self, seq):
        b = seq.encode()
        n = lib().emu_add(self.h, b, len(b))
        nodes = np.empty(n, np.int32)
        pos = np.empty(n, np.int32)
        lib().emu_last_alignment(self.h, nodes.ctypes.data, pos.ctypes.data, n)
        return np.stack([nodes, pos], axis=1)

    def rank_to_node(self):
        n = lib().emu_num_nodes(self.h)
        out = np.empty(n, np.int32)
        lib().emu_rank_to_node(self.h, out.ctypes.data)
        return out

    def consensus(self):
        cap = lib().emu_num_nodes(self.h) + 1
        buf = ctypes.create_string_buffer(cap)
        n = lib().emu_consensus(self.h, buf, cap)
        return buf.raw[:n].decode()

    def msa(self):
        r, c = ctypes.c_int64(), ctypes.c_int64()
        lib().emu_msa_dims(self.h, ctypes.byref(r), ctypes.byref(c))
        if r.value == 0 or c.value == 0:
            return ["" for _ in range(r.value)]
        buf = ctypes.create_string_buffer(r.value * c.value)
        lib().emu_msa(self.h, buf)
        return [buf.raw[i * c.value:(i + 1) * c.value].decode() for i in range(r.value)]

    def close(self):
        if self.h:
            lib().emu_free(self.h)
            self.h = None


# --------------------------------------------------------------------------------------
# MisScore kernel emulation (misscore_emul.cpp)
MIS_LIB = os.path.join(HERE, "_build", "libmisscore_emul.so")
_mis = None


def mis_lib():
    global _mis
    if _mis is None:
        src = os.path.join(HERE, "misscore_emul.cpp")
        deps = [src] + [os.path.join(ROOT, "svscope_b200", "csrc", h) for h in ("misscore_cell.h", "misscore_tb.h")]
        if not os.path.exists(MIS_LIB) or any(os.path.getmtime(d) > os.path.getmtime(MIS_LIB) for d in deps):
            os.makedirs(os.path.dirname(MIS_LIB), exist_ok=True)
            subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-w", src, "-o", MIS_LIB], check=True)
        L = ctypes.CDLL(MIS_LIB)
        L.mis_emul.restype = ctypes.c_int
        L.mis_emul.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_int, ctypes.c_int,
                               ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                               ctypes.c_void_p]
        _mis = L
    return _mis


def misscore_emul(a: str, b: str, match=1, mismatch=0, gap=1, threads=256, cols=16, want_line=False):
    """dict(score, length, matches, status[, line]) from the emulated kernel."""
    ab, bb = a.encode(), b.encode()
    out = np.zeros(4, np.int32)
    line = ctypes.create_string_buffer(len(ab) + len(bb) + 1) if want_line else None
    rc = mis_lib().mis_emul(ab, len(ab), bb, len(bb), match, mismatch, gap, threads, cols, out.ctypes.data, line)
    res = dict(score=int(out[0]), length=int(out[1]), matches=int(out[2]), status=int(rc))
    if want_line:
        res["line"] = line.raw[:res["length"]].decode()
    return res
