"""CPU-side checks of the C-ABI library: it builds for sm_100a, loads, and exports every symbol that
include/drlz.h declares.  No compute call is made here (no GPU in this container)."""
import ctypes
import os
import re

from pangenspark_b200 import drlz as D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_header_symbols():
    so = D.build_library()
    assert os.path.exists(so)
    lib = ctypes.CDLL(so)
    hdr = open(os.path.join(ROOT, "include", "drlz.h")).read()
    declared = sorted(set(re.findall(r"\b(drlz_[a-z_0-9]+)\s*\(", hdr)))
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"libdrlz.so does not export {name}"
    assert sorted(D.SYMBOLS) == declared


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        return
    lib = D.load_library()
    h = ctypes.c_void_p()
    rc = lib.drlz_create(ctypes.byref(h), 0)
    assert rc != 0 and not h.value          # no CPU fallback


def test_sass_is_sm100():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", D.build_library()], capture_output=True, text=True).stdout
    assert "sm_100a" in out
