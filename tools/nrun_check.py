"""The 4-bit generic-alphabet path on a genome-like reference: chr21-sized, ACGT plus N runs as GRCh37 carries them
(a telomere, a centromere-sized run of `nrun` bases, a few short ones), haplotypes with BASELINE's SNP / indel rates whose
N runs stay N.  Device-resident throughput next to the 2-bit fast path's, and parity of `check` rows with the oracle.

    python tools/nrun_check.py [n] [H] [nrun]        (GPU box)
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from pangenspark_b200 import drlz as D  # noqa: E402
import bench  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 48_000_000
    H = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    nrun = int(sys.argv[3]) if len(sys.argv) > 3 else 3_000_000
    check = int(os.environ.get("NRUN_CHECK_ROWS", 2))
    d, rows, M, stride = bench.gen_nrun_workload(orc, n, H, nrun)
    ctx = D.Drlz(0)
    for kv in filter(None, os.environ.get("NRUN_OPTIONS", "").split(",")):      # library options, e.g. occ_pack=6
        ctx.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    t0 = time.time(); ctx.build_index(d); t_first = time.time() - t0
    ctx.build_index(d)
    info, it = ctx.index_info(), ctx.index_timings()
    dev = torch.from_numpy(rows.reshape(-1)).cuda()
    row_off = np.arange(H, dtype=np.uint64) * np.uint64(stride)
    cols = np.full(H, M, dtype=np.uint64)
    for _ in range(2):
        ptr, npairs, m = ctx.parse_batch_device(dev.data_ptr(), row_off, cols)
    acc, R = {}, 3
    for _ in range(R):
        ptr, npairs, m = ctx.parse_batch_device(dev.data_ptr(), row_off, cols)
        for a, b in ctx.parse_timings().items():
            acc[a] = acc.get(a, 0.0) + b / R
    pairs = ctx.copy_pairs_to_host(ptr, int(npairs.sum()))
    ok = True
    if check:
        sa = orc.sa_build(d)
        ok = bool((ctx.index_get("sa").astype(np.int64) == sa).all())
        o = 0
        for h in range(H):
            c = int(npairs[h])
            if h < check:
                a1, _ = orc.encode_literal(d, sa, orc.strip_gaps(rows[h, :M]))
                ok = ok and orc.lz_bytes(pairs[o:o + c]) == orc.lz_bytes(a1)
            o += c
    out = {"workload": "%.0f Mbp ACGT reference with N runs (longest %.1f Mbp), %d haplotypes, 4-bit symbols" % (n / 1e6, nrun / 1e6, H),
           "bits": info["bits"], "sigma": info["sigma"], "kmer": info["k"], "long_repeats": info.get("long_repeats"), "index_ms": it["total_ms"], "index_first_call_s": t_first,
           "sa_rounds": info["sa_rounds"], "gbp_per_s": float(m.sum()) / acc["device_ms"] / 1e6,
           "kernel_ms": {a: round(b, 3) for a, b in acc.items()}, "phrases": int(npairs.sum()), "parity_rows": check, "ok": ok}
    print(json.dumps(out))
    ctx.close()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
