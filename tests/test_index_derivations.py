"""The claims behind the index build's shortcuts (pangenspark_b200/csrc/sa_build.cu), restated in numpy and checked
against the oracle's suffix array and LCP array on the CPU.  The kernels themselves are compared with the oracle in the
`-m gpu` suite (tests/test_gpu_parity.py::_check_index, test_suffix_array_few_rounds, test_kmer_table_variants); this
file pins the arithmetic they rely on:

  * first-round sort key = first S0 symbols (zero padded) | min(remaining, S0): sorting by it puts the suffixes in
    suffix-array order wherever the keys differ, a proper prefix in front of its extensions;
  * k_sa_finish0: with no two equal keys, LCP[j] can be read off the keys of j-1 and j, and the uniqueness array is
    uniq[SA[j]] = max(LCP[j], LCP[j+1]);
  * k_lcp_ties: with equal keys present, the value read off the keys is exactly S0 for the pairs inside a group of equal
    keys (and only for those), and exact for every other pair -- also after the order inside the groups has changed;
  * k_tab_build: the k-mer keys are non-decreasing along the suffix array, so writing j into the entries
    (key[j-1], key[j]] (0 into [0, key[0]], n above key[n-1]) gives tab[K] = number of suffixes with key < K.
"""
import numpy as np
import pytest

from oracle import oracle as orc

S0 = 29            # symbols of the first sort key at 2 bits per symbol (58 >> 1)


def _codes(d: bytes) -> np.ndarray:
    lut = np.zeros(256, dtype=np.int64)
    lut[list(b"ACGT")] = np.arange(4)
    return lut[np.frombuffer(d, dtype=np.uint8)]


def _sort_keys(codes: np.ndarray, s0: int = S0):
    """(58-bit symbol field, length field) of every suffix, as k_sa_init_keys builds them"""
    n = codes.size
    pad = np.concatenate([codes, np.zeros(s0, dtype=np.int64)])
    sym = np.zeros(n, dtype=np.uint64)
    for t in range(s0):
        sym = (sym << np.uint64(2)) | pad[t:t + n].astype(np.uint64)
    length = np.minimum(n - np.arange(n), s0).astype(np.uint64)
    return sym, length


def _lcp_from_keys(sym_a, len_a, sym_b, len_b, s0: int = S0):
    x = sym_a ^ sym_b
    l = np.full(x.shape, s0, dtype=np.int64)
    nz = x != 0
    # leading equal symbols of two 2*s0-bit fields
    bits = np.zeros(x.shape, dtype=np.int64)
    xs = x.copy()
    for _ in range(2 * s0):                       # position of the highest set bit, one bit per round
        bits += (xs != 0)
        xs >>= np.uint64(1)
    l[nz] = (2 * s0 - bits[nz]) // 2
    return np.minimum(l, np.minimum(len_a.astype(np.int64), len_b.astype(np.int64)))


def _cases():
    rng = np.random.default_rng(11)
    out = []
    for n in (1, 2, 30, 31, 500, 5000):
        out.append(bytes(rng.choice(list(b"ACGT"), size=n).astype(np.uint8)))
    d = rng.choice(list(b"ACGT"), size=6000).astype(np.uint8)
    d[4000:4100] = d[1000:1100]                   # a repeat of 100 bases: equal first keys
    d[-40:] = ord("A")                            # and a run of the smallest symbol at the end of the text
    out.append(d.tobytes())
    out.append(b"ACGT" * 50 + b"A" * 70)
    return out


@pytest.mark.parametrize("d", _cases(), ids=lambda d: "n%d" % len(d))
def test_first_keys_order_lcp_and_uniqueness(d):
    n = len(d)
    codes = _codes(d)
    sa = orc.sa_build(d)
    lcp = orc.lcp_build(d, sa)
    sym, length = _sort_keys(codes)
    key = (sym << np.uint64(6)) | length                      # the 64-bit key: symbols above the length
    ks = key[sa]
    # suffix-array order is key order (stable inside equal keys is all the sort promises)
    assert (ks[1:] >= ks[:-1]).all()
    from_keys = np.zeros(n, dtype=np.int64)
    if n > 1:
        from_keys[1:] = _lcp_from_keys(sym[sa[:-1]], length[sa[:-1]], sym[sa[1:]], length[sa[1:]])
    tie = np.zeros(n, dtype=bool)
    tie[1:] = ks[1:] == ks[:-1]
    # exact wherever the two keys differ; exactly S0 for the pairs inside a group of equal keys, and only there
    assert (from_keys[~tie] == lcp[~tie]).all()
    assert (from_keys[tie] == S0).all() and (lcp[tie] >= S0).all()
    assert not ((from_keys == S0) & ~tie & (lcp != S0)).any()
    # what k_lcp_ties does with the rest: compare on from symbol S0 in the FINAL order
    fixed = from_keys.copy()
    for j in np.nonzero(from_keys == S0)[0]:
        a, b, h = int(sa[j - 1]), int(sa[j]), S0
        while a + h < n and b + h < n and d[a + h] == d[b + h]:
            h += 1
        fixed[j] = h
    assert (fixed == lcp).all()
    # uniqueness array scattered through SA == gathered through ISA
    nxt = np.concatenate([lcp[1:], [0]])
    uniq_scatter = np.empty(n, dtype=np.int64)
    uniq_scatter[sa] = np.maximum(lcp, nxt)
    isa = np.empty(n, dtype=np.int64)
    isa[sa] = np.arange(n)
    lcp_pad = np.concatenate([lcp, [0]])
    assert (uniq_scatter == np.maximum(lcp_pad[isa], lcp_pad[isa + 1])).all()


@pytest.mark.parametrize("k", [1, 3, 6])
@pytest.mark.parametrize("d", _cases(), ids=lambda d: "n%d" % len(d))
def test_kmer_table_written_once(d, k):
    n = len(d)
    codes = _codes(d)
    sa = orc.sa_build(d)
    pad = np.concatenate([codes, np.zeros(k, dtype=np.int64)])
    keys = np.zeros(n, dtype=np.int64)
    for t in range(k):
        keys = keys * 4 + pad[sa + t]
    assert (keys[1:] >= keys[:-1]).all()                      # non-decreasing along the suffix array
    cnt = 4 ** k + 1
    tab = np.full(cnt, -1, dtype=np.int64)
    writes = np.zeros(cnt, dtype=np.int64)
    for j in range(n):                                        # the thread of suffix j (k_tab_build / k_sa_finish0)
        start = keys[j - 1] + 1 if j else 0
        tab[start:keys[j] + 1] = j
        writes[start:keys[j] + 1] += 1
    tab[keys[n - 1] + 1:] = n
    writes[keys[n - 1] + 1:] += 1
    assert (writes == 1).all()                                # every entry exactly once
    assert (tab == np.searchsorted(keys, np.arange(cnt), side="left")).all()
    # and the key of the table is a head of the first sort key (k <= S0), which is why k_sa_finish0 can write it
    sym, _ = _sort_keys(codes)
    assert ((sym[sa] >> np.uint64(2 * (S0 - k))).astype(np.int64) == keys).all()
