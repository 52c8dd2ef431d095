"""ctypes binding of libdrlz.so (include/drlz.h).

This is plumbing for the tests, the bench and the Python mirror of the reference's `DistributedRLZ.main`
(distributed_rlz.py).  There is no CPU fallback: if the CUDA library is missing or no device is present the
calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libdrlz.so")

ERRORS = {0: "OK", -1: "E_ARG", -2: "E_CUDA", -3: "E_NOMEM", -4: "E_NOSPACE", -5: "E_ALPHABET", -6: "E_STATE",
          -7: "E_LIMIT", -8: "E_INTERNAL"}

SYMBOLS = [
    "drlz_create", "drlz_destroy", "drlz_last_error", "drlz_set_stream", "drlz_set_option", "drlz_host_alloc",
    "drlz_host_free", "drlz_build_index", "drlz_index_get", "drlz_index_info", "drlz_index_timings",
    "drlz_index_reserve", "drlz_index_buffers", "drlz_index_commit", "drlz_parse", "drlz_parse_batch",
    "drlz_parse_batch_device", "drlz_parse_timings", "drlz_matching_stats", "drlz_copy_to_host", "drlz_submit", "drlz_wait",
    "drlz_parse_batch_gaps", "drlz_submit_gaps",
]


class DrlzError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERRORS.get(code, code)}: {msg}")
        self.code = code


def build_library(force: bool = False) -> str:
    """Compile libdrlz.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    srcs = [os.path.join(_PKG, "csrc", f) for f in os.listdir(os.path.join(_PKG, "csrc"))]
    srcs += [os.path.join(_PKG, "..", "include", "drlz.h"), os.path.join(_PKG, "Makefile")]
    stale = not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _PKG, "-s", "lib/libdrlz.so"])
    return LIB_PATH


_LIB = None


def load_library() -> C.CDLL:
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(f"{LIB_PATH} missing: run __graft_entry__.build() (no CPU fallback exists)")
        L = C.CDLL(LIB_PATH)
        vp, u64, i32 = C.c_void_p, C.c_uint64, C.c_int
        L.drlz_create.argtypes = [C.POINTER(vp), i32]; L.drlz_create.restype = i32
        L.drlz_destroy.argtypes = [vp]; L.drlz_destroy.restype = None
        L.drlz_last_error.argtypes = [vp]; L.drlz_last_error.restype = C.c_char_p
        L.drlz_set_stream.argtypes = [vp, vp]; L.drlz_set_stream.restype = i32
        L.drlz_set_option.argtypes = [vp, C.c_char_p, C.c_int64]; L.drlz_set_option.restype = i32
        L.drlz_host_alloc.argtypes = [C.c_size_t]; L.drlz_host_alloc.restype = vp
        L.drlz_host_free.argtypes = [vp]; L.drlz_host_free.restype = None
        L.drlz_build_index.argtypes = [vp, vp, u64]; L.drlz_build_index.restype = i32
        L.drlz_index_get.argtypes = [vp, i32, vp, u64]; L.drlz_index_get.restype = i32
        L.drlz_index_info.argtypes = [vp, vp]; L.drlz_index_info.restype = i32
        L.drlz_index_timings.argtypes = [vp, vp]; L.drlz_index_timings.restype = i32
        L.drlz_index_reserve.argtypes = [vp, u64, i32]; L.drlz_index_reserve.restype = i32
        L.drlz_index_buffers.argtypes = [vp, vp, vp]; L.drlz_index_buffers.restype = i32
        L.drlz_index_commit.argtypes = [vp]; L.drlz_index_commit.restype = i32
        L.drlz_parse.argtypes = [vp, vp, u64, i32, vp, vp, vp, u64, vp]; L.drlz_parse.restype = i32
        L.drlz_parse_batch.argtypes = [vp, vp, vp, C.c_uint32, i32, vp, vp, vp, u64, vp]; L.drlz_parse_batch.restype = i32
        L.drlz_parse_batch_device.argtypes = [vp, vp, vp, vp, C.c_uint32, i32, vp, vp, vp]
        L.drlz_parse_batch_device.restype = i32
        L.drlz_parse_timings.argtypes = [vp, vp]; L.drlz_parse_timings.restype = i32
        L.drlz_copy_to_host.argtypes = [vp, vp, vp, u64]; L.drlz_copy_to_host.restype = i32
        L.drlz_submit.argtypes = [vp, vp, vp, C.c_uint32, i32, vp, vp, vp, u64, vp, vp]; L.drlz_submit.restype = i32
        L.drlz_wait.argtypes = [vp, u64]; L.drlz_wait.restype = i32
        L.drlz_parse_batch_gaps.argtypes = [vp, vp, vp, C.c_uint32, i32, vp, vp, vp, u64, vp, vp, u64, vp]
        L.drlz_parse_batch_gaps.restype = i32
        L.drlz_submit_gaps.argtypes = [vp, vp, vp, C.c_uint32, i32, vp, vp, vp, u64, vp, vp, u64, vp, vp]
        L.drlz_submit_gaps.restype = i32
        L.drlz_matching_stats.argtypes = [vp, vp, u64, vp, vp]; L.drlz_matching_stats.restype = i32
        _LIB = L
    return _LIB


def _u8(a) -> np.ndarray:
    if isinstance(a, (bytes, bytearray, memoryview)):
        a = np.frombuffer(bytes(a), dtype=np.uint8)
    return np.ascontiguousarray(a, dtype=np.uint8)


class PinnedBuffer:
    """cudaHostAlloc'd memory exposed as a numpy array (for the streaming path)."""

    def __init__(self, nbytes: int):
        self._lib = load_library()
        self.nbytes = int(nbytes)
        self.ptr = self._lib.drlz_host_alloc(max(1, self.nbytes))
        if not self.ptr:
            raise MemoryError("drlz_host_alloc")
        self.array = np.ctypeslib.as_array(C.cast(self.ptr, C.POINTER(C.c_uint8)), shape=(max(1, self.nbytes),))[: self.nbytes]

    def free(self):
        if self.ptr:
            self.array = None
            self._lib.drlz_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Drlz:
    """One context per GPU (one process per GPU)."""

    def __init__(self, device: int = 0):
        self._lib = load_library()
        h = C.c_void_p()
        rc = self._lib.drlz_create(C.byref(h), device)
        if rc != 0:
            raise DrlzError(rc, "drlz_create failed (no CUDA device? there is no CPU fallback)")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.drlz_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise DrlzError(rc, self._lib.drlz_last_error(self._h).decode())

    def set_option(self, key: str, value: int):
        self._ck(self._lib.drlz_set_option(self._h, key.encode(), int(value)))

    def set_stream(self, cuda_stream_ptr: int | None):
        self._ck(self._lib.drlz_set_stream(self._h, C.c_void_p(cuda_stream_ptr or 0)))

    # ---- index ------------------------------------------------------------------------------------
    def build_index(self, ref):
        ref = _u8(ref)
        self._ck(self._lib.drlz_build_index(self._h, ref.ctypes.data_as(C.c_void_p), ref.size))

    def index_info(self):
        a = (C.c_uint64 * 4)()
        self._ck(self._lib.drlz_index_info(self._h, a))
        kinfo = int(a[1])           # k | log2(bits per symbol) << 8 | alphabet size << 16
        return {"n": int(a[0]), "k": kinfo & 0xFF, "bits": 1 << ((kinfo >> 8) & 0xFF), "sigma": (kinfo >> 16) & 0x1FF, "long_repeats": bool((kinfo >> 28) & 1), "kinfo": kinfo,
                "sa_rounds": int(a[2]), "bytes": int(a[3])}

    def index_timings(self):
        a = (C.c_double * 5)()
        self._ck(self._lib.drlz_index_timings(self._h, a))
        return {"total_ms": a[0], "pack_ms": a[1], "sa_ms": a[2], "table_ms": a[3], "lcp_ms": a[4]}

    def index_get(self, what: str) -> np.ndarray:
        info = self.index_info()
        n, k = info["n"], info["k"]
        spw = 64 // info["bits"]
        kinds = {"sa": (0, np.uint32, n), "isa": (1, np.uint32, n), "lcp": (2, np.uint32, n),
                 "tab": (3, np.uint32, info["sigma"] ** k + 1), "text": (4, np.uint64, (n + spw - 1) // spw), "uniq": (5, np.uint16, n)}
        code, dt, cnt = kinds[what]
        out = np.empty(cnt, dtype=dt)
        self._ck(self._lib.drlz_index_get(self._h, code, out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    def index_reserve(self, n: int, k: int):
        self._ck(self._lib.drlz_index_reserve(self._h, n, k))

    def index_buffers(self):
        ptrs = (C.c_void_p * 3)()
        nb = (C.c_uint64 * 3)()
        self._ck(self._lib.drlz_index_buffers(self._h, ptrs, nb))
        return [(int(ptrs[i] or 0), int(nb[i])) for i in range(3)]

    def index_commit(self):
        self._ck(self._lib.drlz_index_commit(self._h))

    # ---- parse ------------------------------------------------------------------------------------
    def parse_batch(self, rows, strip_gaps: bool = True, want_gapless: bool = False, pairs_cap: int | None = None,
                    out: np.ndarray | None = None, gapless_out: list | None = None):
        """rows: list of uint8 arrays / bytes (host).  Returns (list of [np,2] int64 arrays, m[H], gapless list|None).
        gapless_out: one uint8 array per row (at least as long as the row; e.g. views of pinned memory) to receive the
        gapless sequences instead of freshly allocated arrays."""
        rows = [_u8(r) for r in rows]
        H = len(rows)
        ptrs = (C.c_void_p * max(1, H))(*[r.ctypes.data if r.size else None for r in rows])
        cols = np.array([r.size for r in rows], dtype=np.uint64)
        m = np.zeros(max(1, H), dtype=np.uint64)
        npairs = np.zeros(max(1, H), dtype=np.uint64)
        gl = None
        glp = None
        if want_gapless:
            gl = gapless_out if gapless_out is not None else [np.empty(max(1, r.size), dtype=np.uint8) for r in rows]
            if len(gl) != H or any(g.size < r.size for g, r in zip(gl, rows)):
                raise ValueError("gapless_out: one array per row, at least as long as the row")
            glp = (C.c_void_p * max(1, H))(*[g.ctypes.data for g in gl])
        cap = pairs_cap if pairs_cap is not None else max(1024, int(cols.sum()) // 64 + 64 * H)
        while True:
            buf = out if out is not None and out.size >= 2 * cap else np.empty(2 * cap, dtype=np.int64)
            rc = self._lib.drlz_parse_batch(self._h, ptrs, cols.ctypes.data_as(C.c_void_p), H, int(strip_gaps), glp,
                                            m.ctypes.data_as(C.c_void_p), buf.ctypes.data_as(C.c_void_p), cap,
                                            npairs.ctypes.data_as(C.c_void_p))
            if rc == -4 and pairs_cap is None:
                cap = int(npairs[:H].sum()) + 16
                out = None
                continue
            self._ck(rc)
            break
        res, o = [], 0
        for h in range(H):
            c = int(npairs[h])
            res.append(buf[2 * o: 2 * (o + c)].reshape(-1, 2))
            o += c
        m = m[:H].astype(np.int64)
        if gl is not None:
            gl = [g[: int(m[h])] for h, g in enumerate(gl)]
        return res, m, gl

    def parse_batch_gaps(self, rows, strip_gaps: bool = True, gaps_cap: int | None = None):
        """parse_batch plus the gap-position side table: returns (pairs list, m[H], list of [runs,2] int64 arrays
        (first column, length) of every run of '-' per row)."""
        rows = [_u8(r) for r in rows]
        H = len(rows)
        ptrs = (C.c_void_p * max(1, H))(*[r.ctypes.data if r.size else None for r in rows])
        cols = np.array([r.size for r in rows], dtype=np.uint64)
        m = np.zeros(max(1, H), dtype=np.uint64)
        npairs = np.zeros(max(1, H), dtype=np.uint64)
        ngaps = np.zeros(max(1, H), dtype=np.uint64)
        cap = max(1024, int(cols.sum()) // 64 + 64 * H)
        gcap = gaps_cap if gaps_cap is not None else max(1024, int(cols.sum()) // 256 + 64 * H)
        while True:
            buf = np.empty(2 * cap, dtype=np.int64)
            gbuf = np.empty(2 * max(1, gcap), dtype=np.int64)
            rc = self._lib.drlz_parse_batch_gaps(self._h, ptrs, cols.ctypes.data_as(C.c_void_p), H, int(strip_gaps), None,
                                                 m.ctypes.data_as(C.c_void_p), buf.ctypes.data_as(C.c_void_p), cap,
                                                 npairs.ctypes.data_as(C.c_void_p), gbuf.ctypes.data_as(C.c_void_p), gcap,
                                                 ngaps.ctypes.data_as(C.c_void_p))
            if rc == -4:
                cap = max(cap, int(npairs[:H].sum()) + 16)
                gcap = max(gcap, int(ngaps[:H].sum()) + 16)
                continue
            self._ck(rc)
            break
        res, gaps, o, go = [], [], 0, 0
        for h in range(H):
            c, g = int(npairs[h]), int(ngaps[h])
            res.append(buf[2 * o: 2 * (o + c)].reshape(-1, 2)); o += c
            gaps.append(gbuf[2 * go: 2 * (go + g)].reshape(-1, 2)); go += g
        return res, m[:H].astype(np.int64), gaps

    def submit(self, rows, pairs_cap: int, strip_gaps: bool = True):
        """Async form: returns a handle; call .wait() for (pairs list, m).  Buffers are kept alive by the handle."""
        rows = [_u8(r) for r in rows]
        H = len(rows)
        st = {"rows": rows, "ptrs": (C.c_void_p * max(1, H))(*[r.ctypes.data if r.size else None for r in rows]),
              "cols": np.array([r.size for r in rows], dtype=np.uint64), "m": np.zeros(max(1, H), dtype=np.uint64),
              "np": np.zeros(max(1, H), dtype=np.uint64), "buf": np.empty(2 * pairs_cap, dtype=np.int64), "H": H}
        t = C.c_uint64(0)
        self._ck(self._lib.drlz_submit(self._h, st["ptrs"], st["cols"].ctypes.data_as(C.c_void_p), H, int(strip_gaps), None,
                                       st["m"].ctypes.data_as(C.c_void_p), st["buf"].ctypes.data_as(C.c_void_p), pairs_cap,
                                       st["np"].ctypes.data_as(C.c_void_p), C.byref(t)))
        st["ticket"] = t.value
        ctx = self

        class Handle:
            def wait(self_inner):
                ctx._ck(ctx._lib.drlz_wait(ctx._h, st["ticket"]))
                res, o = [], 0
                for h in range(st["H"]):
                    c = int(st["np"][h])
                    res.append(st["buf"][2 * o: 2 * (o + c)].reshape(-1, 2))
                    o += c
                return res, st["m"][: st["H"]].astype(np.int64)
        return Handle()

    def parse(self, row, strip_gaps: bool = True):
        res, m, _ = self.parse_batch([row], strip_gaps)
        return res[0], int(m[0])

    def parse_batch_device(self, rows_dev_ptr: int, row_off, cols, strip_gaps: bool = True):
        """Rows already in HBM.  Returns (pairs device pointer, n_pairs[H], m[H]).  (The argument arrays of the last call
        are kept and reused when the same ones come again: a timed loop then costs one ctypes call per step.)"""
        key = (id(row_off), id(cols))
        st = getattr(self, "_dev_call", None)
        if st is None or st["key"] != key:
            ro = np.ascontiguousarray(row_off, dtype=np.uint64)
            co = np.ascontiguousarray(cols, dtype=np.uint64)
            H = co.size
            st = {"key": key, "keep": (row_off, cols), "ro": ro, "co": co, "H": H, "m": np.zeros(max(1, H), dtype=np.uint64),
                  "np": np.zeros(max(1, H), dtype=np.uint64), "pd": C.c_void_p()}
            st["args"] = (ro.ctypes.data_as(C.c_void_p), co.ctypes.data_as(C.c_void_p), H, st["m"].ctypes.data_as(C.c_void_p),
                          C.byref(st["pd"]), st["np"].ctypes.data_as(C.c_void_p))
            self._dev_call = st
        a = st["args"]
        self._ck(self._lib.drlz_parse_batch_device(self._h, C.c_void_p(rows_dev_ptr), a[0], a[1], a[2], int(strip_gaps), a[3], a[4], a[5]))
        H = st["H"]
        return int(st["pd"].value or 0), st["np"][:H].astype(np.int64), st["m"][:H].astype(np.int64)

    def copy_pairs_to_host(self, dev_ptr: int, total_pairs: int) -> np.ndarray:
        out = np.empty((total_pairs, 2), dtype=np.int64)
        self._ck(self._lib.drlz_copy_to_host(self._h, out.ctypes.data_as(C.c_void_p), C.c_void_p(dev_ptr), out.nbytes))
        return out

    def parse_timings(self):
        a = (C.c_double * 12)()
        self._ck(self._lib.drlz_parse_timings(self._h, a))
        keys = ["memset_ms", "prescan_ms", "host_ms", "pack_write_ms", "spec_ms", "fix_ms", "jump_ms", "emit_ms",
                "device_ms", "first_ms", "fix_rounds", "launches"]
        return dict(zip(keys, list(a)))

    def matching_stats(self, x):
        x = _u8(x)
        ms = np.zeros(x.size, dtype=np.uint32)
        pos = np.zeros(x.size, dtype=np.uint32)
        self._ck(self._lib.drlz_matching_stats(self._h, x.ctypes.data_as(C.c_void_p), x.size,
                                               ms.ctypes.data_as(C.c_void_p), pos.ctypes.data_as(C.c_void_p)))
        return ms, pos
