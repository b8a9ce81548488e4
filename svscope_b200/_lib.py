"""ctypes binding of libsvscope_b200.so (C ABI in include/svscope_b200.h).

There is no CPU fallback: importing this module never fails, but any call into the library
raises ``RuntimeError`` when the shared object has not been built or no CUDA device is
usable."""
from __future__ import annotations

import ctypes
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# SVS_LIB: an instrumented build of the same library (scripts/dp_phase_profile.sh); never a different implementation
LIB_PATH = os.environ.get("SVS_LIB") or os.path.join(_HERE, "_C", "libsvscope_b200.so")

_lib = None
_lock = threading.Lock()

c_i64p = ctypes.POINTER(ctypes.c_int64)
c_vp = ctypes.c_void_p

SYMBOLS = {
    "svs_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(c_vp)]),
    "svs_destroy": (None, [c_vp]),
    "svs_last_error": (ctypes.c_char_p, [c_vp]),
    "svs_version": (ctypes.c_char_p, []),
    "svs_set_option": (ctypes.c_int, [c_vp, ctypes.c_char_p, ctypes.c_int64]),
    "svs_get_option": (ctypes.c_int64, [c_vp, ctypes.c_char_p]),
    "svs_int_alu_probe": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int]),
    "svs_reads_upload": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.c_int64, ctypes.POINTER(c_vp)]),
    "svs_reads_free": (None, [c_vp]),
    "svs_poa_batch": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, ctypes.c_int64] + [ctypes.c_int] * 8 + [ctypes.POINTER(c_vp)]),
    "svs_poa_submit": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, ctypes.c_int64] + [ctypes.c_int] * 8 + [ctypes.POINTER(c_vp)]),
    "svs_poa_wait": (ctypes.c_int, [c_vp]),
    "svs_poa_result_status": (ctypes.c_int, [c_vp, c_vp]),
    "svs_poa_result_sizes": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp]),
    "svs_poa_result_copy": (ctypes.c_int, [c_vp, c_vp, c_vp]),
    "svs_poa_result_stats": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int]),
    "svs_poa_result_free": (None, [c_vp]),
    "svs_poa_align_pairs": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.c_int64, c_vp, c_vp, ctypes.c_int64, c_vp, c_vp]),
    "svs_msa_features": (ctypes.c_int, [c_vp, ctypes.c_int64] + [c_vp] * 12),
    "svs_em_batch": (ctypes.c_int, [c_vp, ctypes.c_int64] + [c_vp] * 7 + [ctypes.c_int32, c_vp, ctypes.c_int32] + [c_vp] * 9),
    "svs_edit_distance_matrix": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, ctypes.c_int64, c_vp, c_vp, c_vp, ctypes.c_int]),
    "svs_edit_distance_pairs": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, ctypes.c_int64, c_vp, c_vp, ctypes.c_int]),
    "svs_misscore_pairs": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, ctypes.c_int64] + [ctypes.c_int] * 4 + [c_vp] * 4 + [ctypes.c_int]),
}


def load():
    """Load the shared library (no device needed) and declare every exported symbol."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} is missing: build it with `python -m svscope_b200.csrc.build` "
                    "(there is no CPU fallback)")
            L = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SYMBOLS.items():
                fn = getattr(L, name)
                fn.restype = res
                fn.argtypes = args
            _lib = L
    return _lib


def ptr(a):
    return None if a is None else a.ctypes.data


class SvsError(RuntimeError):
    pass


class Context:
    """One CUDA device + scratch arena.  ``Context.default()`` is shared per (process, device)."""

    _defaults = {}

    def __init__(self, device: int = 0, **options):
        L = load()
        h = c_vp()
        rc = L.svs_create(int(device), ctypes.byref(h))
        if rc != 0 or not h.value:
            raise RuntimeError(
                f"svs_create(device={device}) failed (code {rc}): no usable CUDA device; "
                "svscope_b200 has no CPU fallback")
        self._h = h
        self.device = device
        for k, v in options.items():
            self.set_option(k, v)

    @classmethod
    def default(cls, device: int | None = None) -> "Context":
        if device is None:
            device = int(os.environ.get("SVS_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        if device not in cls._defaults:
            cls._defaults[device] = cls(device)
        return cls._defaults[device]

    def check(self, rc: int):
        if rc != 0:
            raise SvsError(f"svscope_b200 error {rc}: {load().svs_last_error(self._h).decode()}")

    def set_option(self, key: str, value: int):
        self.check(load().svs_set_option(self._h, key.encode(), int(value)))

    def get_option(self, key: str) -> int:
        return int(load().svs_get_option(self._h, key.encode()))

    def int_alu_probe(self):
        out = np.zeros(4, np.float64)
        self.check(load().svs_int_alu_probe(self._h, ptr(out), 4))
        return dict(add=float(out[0]), max=float(out[1]), xor=float(out[2]), addmax=float(out[3]))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            load().svs_destroy(self._h)
            self._h = c_vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ReadSet:
    """Sequences resident in device memory (svs_reads_upload)."""

    def __init__(self, ctx: Context, seqs):
        self.ctx = ctx
        if isinstance(seqs, tuple):  # (uint8 buffer, int64 offsets)
            buf, off = seqs
        else:
            enc = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
            off = np.zeros(len(enc) + 1, np.int64)
            if enc:
                off[1:] = np.cumsum([len(b) for b in enc])
            joined = b"".join(enc)
            buf = np.frombuffer(joined, np.uint8) if joined else np.zeros(0, np.uint8)
        self.buf = np.ascontiguousarray(buf, dtype=np.uint8)
        self.off = np.ascontiguousarray(off, dtype=np.int64)
        self.n = self.off.shape[0] - 1
        h = c_vp()
        pad = self.buf if self.buf.size else np.zeros(1, np.uint8)
        ctx.check(load().svs_reads_upload(ctx._h, ptr(pad), ptr(self.off), self.n, ctypes.byref(h)))
        self._h = h

    @property
    def nbytes(self) -> int:
        return int(self.off[-1])

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            load().svs_reads_free(self._h)
            self._h = c_vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
