"""Shared helpers for the tests: brute-force clean-spec parse and random cases."""
import numpy as np


def naive_sa(d: bytes):
    return sorted(range(len(d)), key=lambda i: d[i:])


def brute_parse(d: bytes, x: bytes):
    """A2 clean spec by brute force: longest match; pos = lexicographically smallest suffix among
    the occurrences of the longest match (== SA[leftmost index of its interval])."""
    out = []
    i, n, m = 0, len(d), len(x)
    while i < m:
        best = 0
        for p in range(n):
            l = 0
            while p + l < n and i + l < m and d[p + l] == x[i + l]:
                l += 1
            best = max(best, l)
        if best == 0:
            out.append((x[i], 0)); i += 1
            continue
        pat = x[i:i + best]
        occ = [p for p in range(n) if d[p:p + best] == pat]
        pos = min(occ, key=lambda p: d[p:])
        out.append((pos, best)); i += best
    return np.array(out, dtype=np.int64).reshape(-1, 2)


def random_case(rng, alphabet=b"ACGT", nmax=40, mmax=40, derive=True):
    n = int(rng.integers(1, nmax + 1))
    d = bytes(rng.choice(list(alphabet), size=n).astype(np.uint8))
    if derive and rng.random() < 0.6:
        # haplotype derived from the reference with edits
        x = bytearray()
        i = 0
        while i < n:
            r = rng.random()
            if r < 0.08:
                x.append(int(rng.choice(list(alphabet))))
                i += 1
            elif r < 0.12:
                i += int(rng.integers(1, 4))
            elif r < 0.16:
                x.extend(rng.choice(list(alphabet), size=int(rng.integers(1, 4))).astype(np.uint8).tobytes())
            else:
                x.append(d[i]); i += 1
        x = bytes(x)
    else:
        m = int(rng.integers(0, mmax + 1))
        x = bytes(rng.choice(list(alphabet), size=m).astype(np.uint8))
    return d, x
