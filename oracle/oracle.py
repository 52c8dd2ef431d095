"""ctypes front-end of the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module (see drlz_oracle.c for the reference file:line map).  PARITY UNPINNED: the reference
ships no golden vectors and cannot run here; this restates DistributedRLZ.scala:108-242.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("drlz_oracle.c", "synth.c")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        u8p, i64p, u64p = C.c_void_p, C.c_void_p, C.c_void_p
        L.orc_sa_build.argtypes = [u8p, C.c_uint64, i64p]; L.orc_sa_build.restype = C.c_int
        L.orc_lcp_build.argtypes = [u8p, C.c_uint64, i64p, i64p]; L.orc_lcp_build.restype = C.c_int
        L.orc_encode_literal.argtypes = [u8p, C.c_uint64, i64p, u8p, C.c_uint64, i64p, C.c_uint64, u64p]
        L.orc_encode_literal.restype = C.c_uint64
        L.orc_encode_a3.argtypes = [u8p, C.c_uint64, i64p, i64p, u8p, C.c_uint64, i64p, C.c_uint64]
        L.orc_encode_a3.restype = C.c_uint64
        L.orc_ms_all.argtypes = [u8p, C.c_uint64, i64p, i64p, u8p, C.c_uint64, i64p, i64p]
        L.orc_ms_all.restype = None
        L.orc_strip_gaps.argtypes = [u8p, C.c_uint64, u8p]; L.orc_strip_gaps.restype = C.c_uint64
        L.orc_decode.argtypes = [u8p, C.c_uint64, i64p, C.c_uint64, u8p, C.c_uint64]
        L.orc_decode.restype = C.c_int64
        L.orc_encode_rows_mt.argtypes = [u8p, C.c_uint64, i64p, C.c_void_p, u64p, C.c_uint64, C.c_int,
                                         C.c_void_p, u64p, u64p, u64p, C.c_void_p]
        L.orc_encode_rows_mt.restype = C.c_int
        L.orc_free.argtypes = [C.c_void_p]; L.orc_free.restype = None
        L.orc_max_threads.restype = C.c_int
        L.syn_u.argtypes = [C.c_uint64, C.c_uint64]; L.syn_u.restype = C.c_uint64
        L.syn_reference.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, u8p]; L.syn_reference.restype = None
        L.syn_columns.argtypes = [C.c_uint64, C.c_uint64, C.c_double]; L.syn_columns.restype = C.c_uint64
        L.syn_row.argtypes = [C.c_uint64, u8p, C.c_uint64, C.c_uint64, C.c_double, C.c_double, C.c_double, u8p]
        L.syn_row.restype = C.c_uint64
        L.syn_rows.argtypes = [C.c_uint64, u8p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_double, C.c_double,
                               C.c_double, u8p, C.c_uint64]
        L.syn_rows.restype = None
        _LIB = L
    return _LIB


def _u8(a) -> np.ndarray:
    if isinstance(a, (bytes, bytearray, memoryview)):
        a = np.frombuffer(bytes(a), dtype=np.uint8)
    return np.ascontiguousarray(a, dtype=np.uint8)


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def sa_build(d) -> np.ndarray:
    d = _u8(d)
    sa = np.empty(d.size, dtype=np.int64)
    if lib().orc_sa_build(_p(d), d.size, _p(sa)) != 0:
        raise MemoryError("orc_sa_build")
    return sa


def lcp_build(d, sa) -> np.ndarray:
    d = _u8(d)
    lcp = np.zeros(d.size, dtype=np.int64)
    if lib().orc_lcp_build(_p(d), d.size, _p(sa), _p(lcp)) != 0:
        raise MemoryError("orc_lcp_build")
    return lcp


def strip_gaps(row) -> np.ndarray:
    row = _u8(row)
    out = np.empty(row.size + 1, dtype=np.uint8)
    m = lib().orc_strip_gaps(_p(row), row.size, _p(out))
    return out[:m].copy()


def encode_literal(d, sa, x):
    """A1 (defines parity).  Returns (pairs[np,2] int64, would_throw)."""
    d, x = _u8(d), _u8(x)
    cap = max(16, x.size // 16 + 16)
    wt = C.c_uint64(0)
    while True:
        pairs = np.empty((cap, 2), dtype=np.int64)
        n = lib().orc_encode_literal(_p(d), d.size, _p(sa), _p(x), x.size, _p(pairs), cap, C.byref(wt))
        if n <= cap:
            return pairs[:n].copy(), int(wt.value)
        cap = int(n)


def encode_a3(d, sa, lcp, x) -> np.ndarray:
    d, x = _u8(d), _u8(x)
    cap = max(16, x.size // 16 + 16)
    while True:
        pairs = np.empty((cap, 2), dtype=np.int64)
        n = lib().orc_encode_a3(_p(d), d.size, _p(sa), _p(lcp), _p(x), x.size, _p(pairs), cap)
        if n <= cap:
            return pairs[:n].copy()
        cap = int(n)


def ms_all(d, sa, lcp, x):
    d, x = _u8(d), _u8(x)
    ms = np.zeros(x.size, dtype=np.int64)
    pos = np.zeros(x.size, dtype=np.int64)
    if x.size:
        lib().orc_ms_all(_p(d), d.size, _p(sa), _p(lcp), _p(x), x.size, _p(ms), _p(pos))
    return ms, pos


def decode(d, pairs) -> bytes:
    d = _u8(d)
    pairs = np.ascontiguousarray(pairs, dtype=np.int64).reshape(-1, 2)
    cap = int(pairs[:, 1].sum() + (pairs[:, 1] == 0).sum()) + 1
    out = np.empty(cap, dtype=np.uint8)
    m = lib().orc_decode(_p(d), d.size, _p(pairs), pairs.shape[0], _p(out), cap)
    if m < 0:
        raise ValueError("undecodable phrase stream")
    return out[:m].tobytes()


def encode_rows_mt(d, sa, rows, nthreads: int = 0):
    """Parse MSA rows (gapped) the way local[N] would: one row per task.
    Returns (list of pairs arrays, m per row, would_throw per row, seconds)."""
    d = _u8(d)
    rows = [_u8(r) for r in rows]
    H = len(rows)
    ptrs = (C.c_void_p * H)(*[r.ctypes.data for r in rows])
    cols = np.array([r.size for r in rows], dtype=np.uint64)
    outp = (C.c_void_p * H)()
    npairs = np.zeros(H, dtype=np.uint64)
    m = np.zeros(H, dtype=np.uint64)
    wt = np.zeros(H, dtype=np.uint64)
    sec = C.c_double(0)
    rc = lib().orc_encode_rows_mt(_p(d), d.size, _p(sa), ptrs, _p(cols), H, nthreads, outp, _p(npairs), _p(m),
                                  _p(wt), C.byref(sec))
    res = []
    for h in range(H):
        if outp[h]:
            arr = np.ctypeslib.as_array(C.cast(outp[h], C.POINTER(C.c_int64)), shape=(int(npairs[h]), 2)).copy()
            lib().orc_free(outp[h])
        else:
            arr = np.zeros((0, 2), dtype=np.int64)
        res.append(arr)
    if rc != 0:
        raise MemoryError("orc_encode_rows_mt")
    return res, m.astype(np.int64), wt.astype(np.int64), float(sec.value)


def max_threads() -> int:
    return int(lib().orc_max_threads())


def gap_runs(row) -> np.ndarray:
    """Gap-position side table of one MSA row as [runs, 2] int64 (first column, length): the maximal runs of the
    column list ScoreMatrix.scala:91-93 collects (`for i if row(i) == '-' yield i`), run lengths as its
    `consecutive` helper counts them (:63-79)."""
    row = _u8(row)
    g = (row == ord("-")).astype(np.int8)
    edge = np.diff(np.concatenate([[0], g, [0]]))
    starts, ends = np.flatnonzero(edge == 1), np.flatnonzero(edge == -1)
    return np.stack([starts, ends - starts], axis=1).astype(np.int64).reshape(-1, 2)


def lz_bytes(pairs) -> bytes:
    """DistributedRLZ.scala:263-284: int64 LE pos, int64 LE len per phrase."""
    return np.ascontiguousarray(pairs, dtype="<i8").tobytes()


# ---------------------------------------------------------------------------------------------
# synthetic data; see synth.c
# ---------------------------------------------------------------------------------------------
CONFIGS = {
    # name: (n, H, p_snp, p_ins, p_del)
    "cfg1": (20_000_000, 4, 1e-3, 1e-5, 1e-5),
    "cfg2": (48_000_000, 50, 1e-3, 1e-5, 1e-5),
    "cfg3": (249_000_000, 100, 1e-3, 1e-5, 1e-5),
    "cfg4": (249_000_000, 1000, 1e-3, 1e-5, 1e-5),
    "cfg5": (3_100_000_000, 8, 8e-3, 1e-3, 1e-3),
}


def seed_for(cfg: int) -> int:
    return 0xD1CE0000 + cfg


def syn_reference(seed: int, n: int, repeats: int = 0) -> np.ndarray:
    d = np.empty(n, dtype=np.uint8)
    lib().syn_reference(seed, n, repeats, _p(d))
    return d


def syn_columns(seed: int, n: int, p_ins: float) -> int:
    return int(lib().syn_columns(seed, n, p_ins))


def syn_rows(seed: int, d: np.ndarray, h0: int, H: int, p_snp: float, p_ins: float, p_del: float,
             out: np.ndarray | None = None) -> np.ndarray:
    """Rows h0..h0+H-1 as a [H, M] uint8 array (row 0 = the reference with gaps)."""
    M = syn_columns(seed, d.size, p_ins)
    if out is None:
        out = np.empty((H, M), dtype=np.uint8)
    assert out.shape == (H, M) and out.dtype == np.uint8 and out.flags.c_contiguous
    lib().syn_rows(seed, _p(d), d.size, h0, H, p_snp, p_ins, p_del, _p(out), M)
    return out
