"""Scale check on the GPU box: chr1-sized reference (config 3 shape, a few haplotypes) against the oracle.
usage: python tools/scale_check.py [n] [H] [cfg]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402
from pangenspark_b200 import drlz as D  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 249_000_000
    H = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    cfg = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    name = "cfg%d" % cfg
    _, _, p_snp, p_ins, p_del = orc.CONFIGS[name]
    seed = orc.seed_for(cfg)
    t = time.time(); d = orc.syn_reference(seed, n, repeats=int(os.environ.get("REPEATS", 0))); print("ref", round(time.time() - t, 1), flush=True)
    t = time.time(); rows = orc.syn_rows(seed, d, 0, H, p_snp, p_ins, p_del); print("rows", rows.shape, round(time.time() - t, 1), flush=True)
    ctx = D.Drlz(0)
    t = time.time(); ctx.build_index(d); print("gpu index s", round(time.time() - t, 2), ctx.index_timings(), ctx.index_info(), flush=True)
    t = time.time(); ctx.build_index(d); print("gpu index (2nd) s", round(time.time() - t, 2), ctx.index_timings(), flush=True)
    t = time.time(); res, m, _ = ctx.parse_batch([r for r in rows]); print("gpu parse s", round(time.time() - t, 3), ctx.parse_timings(), [x.shape[0] for x in res], flush=True)
    t = time.time(); sa = orc.sa_build(d); print("cpu sa s", round(time.time() - t, 1), flush=True)
    ok = bool((ctx.index_get("sa").astype(np.int64) == sa).all())
    print("SA equal:", ok, flush=True)
    cres, cm, wt, sec = orc.encode_rows_mt(d, sa, [r for r in rows], 0)
    print("cpu parse s", round(sec, 2), "would_throw", wt.tolist(), flush=True)
    for h in range(H):
        same = orc.lz_bytes(res[h]) == orc.lz_bytes(cres[h])
        ok = ok and same
        print("row", h, "phrases", res[h].shape[0], "bases", int(m[h]), "equal:", same, flush=True)
    print("SCALE CHECK", "PASS" if ok else "FAIL")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
