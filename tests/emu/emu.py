"""ctypes loader for the TEST-ONLY host harness (tests/emu/emu.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libemu.so")
        srcs = [os.path.join(_HERE, "emu.cpp"),
                os.path.join(_HERE, "..", "..", "pangenspark_b200", "csrc", "rlz_core.cuh")]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, srcs[0]])
        L = C.CDLL(so)
        L.emu_parse.restype = C.c_int
        L.emu_free.argtypes = [C.c_void_p]
        L.emu_long_hits.restype = C.c_uint64
        _LIB = L
    return _LIB


def parse(ref: bytes, sa: np.ndarray, rows, k: int, chunk: int, min_cap: int = 64, foreign_cap: int = 1 << 20, use_uniq: int = 1):
    L = lib()
    d = np.frombuffer(bytes(ref), dtype=np.uint8)
    sa = np.ascontiguousarray(sa, dtype=np.int64)
    rows = [np.frombuffer(bytes(r), dtype=np.uint8) for r in rows]
    H = len(rows)
    ptrs = (C.c_void_p * H)(*[r.ctypes.data if r.size else None for r in rows])
    cols = np.array([r.size for r in rows], dtype=np.uint64)
    outp = (C.c_void_p * H)()
    npairs = np.zeros(H, dtype=np.uint64)
    m = np.zeros(H, dtype=np.uint64)
    rounds = C.c_int(0)
    rc = L.emu_parse(d.ctypes.data_as(C.c_void_p), C.c_uint64(d.size), sa.ctypes.data_as(C.c_void_p), C.c_int(k),
                     C.c_uint32(chunk), C.c_uint32(min_cap), C.c_uint32(foreign_cap), C.c_int(use_uniq), C.c_int(H), ptrs, cols.ctypes.data_as(C.c_void_p), outp,
                     npairs.ctypes.data_as(C.c_void_p), m.ctypes.data_as(C.c_void_p), C.byref(rounds))
    if rc != 0:
        return rc, None, None, rounds.value
    res = []
    for h in range(H):
        arr = np.ctypeslib.as_array(C.cast(outp[h], C.POINTER(C.c_int64)), shape=(int(npairs[h]), 2)).copy()
        L.emu_free(outp[h])
        res.append(arr)
    return 0, res, m.astype(np.int64), rounds.value


def long_hits() -> int:
    """phrases the harness took from the extents across chunks so far (long-repeat path of phrase_capped_l)"""
    return int(lib().emu_long_hits())
