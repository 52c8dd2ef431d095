"""Index build of a configs[1]-sized reference from pinned memory, three times; prints the library's timings.
usage: python tools/index_probe.py [n]   (run under ncu to capture the sort / finish kernels)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402
from pangenspark_b200 import drlz as D  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 48_000_000
d = orc.syn_reference(orc.seed_for(2), n)
pin = D.PinnedBuffer(n)
pin.array[:] = d
ctx = D.Drlz(0)
variants = [int(v) for v in os.environ.get("SORT_VARIANTS", "").split(",") if v] or [None]
sa0 = None
for v in variants:
    if v is not None:
        ctx.set_option("sort_variant", v)
    for it in range(3):
        ctx.build_index(pin.array)
        print("sort_variant", v, it, ctx.index_timings(), ctx.index_info(), flush=True)
    sa = ctx.index_get("sa")
    if sa0 is None:
        sa0 = sa
    print("sort_variant", v, "same SA as the first variant:", bool((sa == sa0).all()), flush=True)
ctx.close()
