"""Parity of the CUDA path (through the C ABI) against the CPU oracle.  Run on the B200 box: pytest -m gpu."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import oracle as orc
from pangenspark_b200 import drlz as D
from tests.util import random_case

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module", params=[2, 1, 0])
def ctx(request):
    """every test runs with all strip/pack kernels: 2 = k_pack3 (a CTA streams 8 tiles; the default), 1 = k_pack2 (a
    tile per CTA; also the kernel behind the gapless output), 0 = the round-1 kernel; generic alphabets: 2 and 1 =
    k_packg1 (single pass), 0 = the two-pass kernels"""
    c = D.Drlz(0)
    c.set_option("pack_variant", request.param)
    yield c
    c.close()


def test_scan_primitive():
    L = D.load_library()
    rng = np.random.default_rng(0)
    for n in (1, 5, 4095, 4096, 4097, 100_000, 3_000_001):
        a = rng.integers(0, 1000, size=n).astype(np.uint32)
        out = np.zeros(n, dtype=np.uint64)
        tot = C.c_uint64(0)
        rc = L.drlz_debug_scan_u32(a.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), C.c_uint64(n), C.byref(tot))
        assert rc == 0
        ref = np.concatenate([[0], np.cumsum(a.astype(np.uint64))])
        assert (out == ref[:-1]).all()
        assert tot.value == ref[-1]


@pytest.mark.parametrize("sort_variant", [0, 1, 2])
def test_radix_sort_primitive(ctx, sort_variant):
    L = D.load_library()
    ctx.set_option("sort_variant", sort_variant)           # process-wide: which scatter kernel the sort launches
    try:
        _radix_sort_cases(L)
    finally:
        ctx.set_option("sort_variant", DEFAULT_SORT_VARIANT)


DEFAULT_SORT_VARIANT = 1


def _radix_sort_cases(L):
    rng = np.random.default_rng(1)
    for n, bits in ((1, 64), (2, 64), (777, 64), (4096, 64), (4097, 24), (250_000, 64), (1_000_003, 40)):
        keys = rng.integers(0, 2 ** 63, size=n, dtype=np.uint64) * 2 + rng.integers(0, 2, size=n).astype(np.uint64)
        if bits < 64:
            keys &= np.uint64((1 << bits) - 1)
        keys[: n // 3] = keys[0]                       # many duplicates: stability
        vals = np.arange(n, dtype=np.uint32)
        k2, v2 = keys.copy(), vals.copy()
        rc = L.drlz_debug_sort_pairs(k2.ctypes.data_as(C.c_void_p), v2.ctypes.data_as(C.c_void_p), C.c_uint64(n), 0, bits)
        assert rc == 0
        order = np.argsort(keys, kind="stable")
        assert (k2 == keys[order]).all()
        assert (v2 == vals[order]).all()


def _check_index(ctx, d: bytes, k=None):
    if k is not None:
        ctx.set_option("kmer", k)
    ctx.set_option("lcp", 1)
    ctx.build_index(d)
    ctx.set_option("kmer", 0)
    sa_ref = orc.sa_build(d)
    sa = ctx.index_get("sa")
    assert (sa.astype(np.int64) == sa_ref).all()
    isa = ctx.index_get("isa")
    if len(d):
        assert (isa[sa] == np.arange(len(d))).all()
        lcp = ctx.index_get("lcp")
        lcp_ref = orc.lcp_build(d, sa_ref)
        assert (lcp.astype(np.int64) == lcp_ref).all()
        # uniqueness array: the longest repeat starting at p = max of the LCPs of its suffix with both SA neighbours
        nxt = np.concatenate([lcp_ref[1:], [0]])
        uniq_ref = np.empty(len(d), dtype=np.int64)
        uniq_ref[sa_ref] = np.minimum(65535, np.maximum(lcp_ref, nxt))
        assert (ctx.index_get("uniq").astype(np.int64) == uniq_ref).all()
    return sa_ref


@pytest.mark.parametrize("alphabet", [b"ACGT", b"AC", b"A", b"ACGTN", b"ACGNTacgnt", bytes(range(33, 73)), b"AB"])
def test_suffix_array_small(ctx, alphabet):
    rng = np.random.default_rng(len(alphabet))
    for n in list(range(1, 70)) + [100, 257, 1000, 4097, 20000]:
        d = bytes(rng.choice(list(alphabet), size=n).astype(np.uint8))
        _check_index(ctx, d, k=int(rng.integers(1, 7)))


def test_suffix_array_repeats(ctx):
    rng = np.random.default_rng(9)
    unit = bytes(rng.choice(list(b"ACGT"), size=37).astype(np.uint8))
    d = unit * 300 + b"ACGT" * 500 + b"A" * 3000 + unit * 20 + bytes(rng.choice(list(b"ACGT"), size=50000).astype(np.uint8))
    _check_index(ctx, d)
    d2 = orc.syn_reference(orc.seed_for(3), 1_500_000, repeats=40).tobytes()
    _check_index(ctx, d2)
    assert ctx.index_info()["sa_rounds"] >= 1


def test_suffix_array_few_rounds(ctx):
    """A random-like reference with a few planted repeats of 40..200 bases: one to three doubling rounds, the case in which
    the LCP array is taken from the first sort keys and only the pairs with equal keys are compared on (k_lcp_ties)."""
    rng = np.random.default_rng(21)
    for reps, seg in ((3, 40), (20, 90), (5, 200)):
        d = rng.choice(list(b"ACGT"), size=300_000).astype(np.uint8)
        for _ in range(reps):
            a, b = (int(v) for v in rng.integers(0, d.size - seg, size=2))
            d[b:b + seg] = d[a:a + seg]
        d[-35:] = ord("A")                                   # equal keys among the last suffixes as well
        _check_index(ctx, d.tobytes())
        rounds = ctx.index_info()["sa_rounds"]
        assert 1 <= rounds <= 3, rounds
        ctx.set_option("lcp_ties", 0)                        # the Kasai-order kernel gives the same arrays
        _check_index(ctx, d.tobytes())
        ctx.set_option("lcp_ties", 1)


def test_kmer_table_variants(ctx):
    """The table written in one pass (table_variant 1: every entry once, long ranges of empty keys by the warp or through
    the range list) against fill + mark + scan (0), and against its definition: tab[K] = number of suffixes whose
    zero-padded k-mer key is below K."""
    rng = np.random.default_rng(5)
    cases = []
    cases.append((orc.syn_reference(orc.seed_for(2), 300_000).tobytes(), 0))                       # auto k
    cases.append((orc.syn_reference(orc.seed_for(2), 5_000).tobytes(), 9))                         # sparse: 262 144 keys for 5 000 suffixes
    d = orc.syn_reference(orc.seed_for(3), 400_000)
    d[100_000:180_000] = ord("N"); d[7] = ord("N")
    cases.append((d.tobytes(), 0))                                                                 # ACGNT, radix-5 keys, millions of empty keys around N
    cases.append((d.tobytes(), 10))
    cases.append((bytes(rng.choice(list(bytes(range(40, 90))), size=60_000).astype(np.uint8)), 3))  # 8-bit symbols
    cases.append((b"A" * 5000, 5))
    cases.append((b"T" * 100 + b"A", 6))
    cases.append((b"G", 4))
    try:
        for d, k in cases:
            tabs = []
            for variant in (0, 1):
                ctx.set_option("table_variant", variant)
                ctx.set_option("kmer", k)
                ctx.build_index(d)
                tabs.append(ctx.index_get("tab"))
            assert (tabs[0] == tabs[1]).all(), (len(d), k)
            info = ctx.index_info()
            sigma, kk, n = info["sigma"], info["k"], len(d)
            sa = ctx.index_get("sa").astype(np.int64)
            alphabet = sorted(set(d))
            code_of = np.zeros(256, dtype=np.int64)
            if set(alphabet) <= set(b"ACGT"):
                code_of[list(b"ACGT")] = np.arange(4)
            else:
                code_of[alphabet] = np.arange(len(alphabet))
            codes = code_of[np.frombuffer(d, dtype=np.uint8)]
            pad = np.concatenate([codes, np.zeros(kk, dtype=np.int64)])
            keys = np.zeros(n, dtype=np.int64)
            for t in range(kk):
                keys = keys * sigma + pad[sa + t] * (sa + t < n)
            expect = np.searchsorted(keys, np.arange(sigma ** kk + 1), side="left")
            assert (tabs[1].astype(np.int64) == expect).all(), (len(d), k)
    finally:
        ctx.set_option("table_variant", 1)
        ctx.set_option("kmer", 0)


def test_kmer_table_and_text(ctx):
    d = orc.syn_reference(orc.seed_for(2), 100_000).tobytes()
    ctx.set_option("kmer", 6)
    ctx.build_index(d)
    ctx.set_option("kmer", 0)
    k = ctx.index_info()["k"]
    assert k == 6
    sa = ctx.index_get("sa").astype(np.int64)
    codes = np.frombuffer(d.translate(bytes.maketrans(b"ACGT", b"\x00\x01\x02\x03")), dtype=np.uint8).astype(np.int64)
    pad = np.concatenate([codes, np.zeros(k, dtype=np.int64)])
    keys = np.zeros(len(d), dtype=np.int64)
    for t in range(k):
        keys = keys * 4 + pad[sa + t] * (sa + t < len(d))
    tab = ctx.index_get("tab").astype(np.int64)
    expect = np.searchsorted(keys, np.arange(4 ** k + 1), side="left")
    assert (tab == expect).all()
    text = ctx.index_get("text")
    w0 = 0
    for t in range(32):
        w0 = (w0 << 2) | int(codes[t])
    assert int(text[0]) == w0


def _check_rows(ctx, d: bytes, rows, sa=None, clean=False):
    """clean=False: against A1, the literal restatement of the reference's binary search.  clean=True: against A3
    (longest match, leftmost SA index), for reference texts with bytes <= '1': there the reference's own search
    compares past-the-end as the byte '1' (DistributedRLZ.scala:118-122) and is inconsistent with its SA order."""
    if sa is None:
        sa = orc.sa_build(d)
    lcp = orc.lcp_build(d, sa) if clean else None
    res, m, gl = ctx.parse_batch(rows, strip_gaps=True, want_gapless=True)
    res2, m2, _ = ctx.parse_batch(rows, strip_gaps=True, want_gapless=False)      # the kernel without the gapless bytes
    assert (m == m2).all() and all(orc.lz_bytes(a) == orc.lz_bytes(b) for a, b in zip(res, res2))
    for h, r in enumerate(rows):
        x = orc.strip_gaps(r).tobytes()
        a1 = orc.encode_a3(d, sa, lcp, x) if clean else orc.encode_literal(d, sa, x)[0]
        assert int(m[h]) == len(x)
        assert gl[h].tobytes() == x
        assert orc.lz_bytes(res[h]) == orc.lz_bytes(a1), (h, len(x), res[h][:5], a1[:5])


@pytest.mark.parametrize("seed", range(3))
def test_parse_random_small(ctx, seed):
    rng = np.random.default_rng(100 + seed)
    for it in range(60):
        d, x = random_case(rng, b"ACGT", 200, 300)
        xb = bytearray(x)
        if rng.random() < 0.5:
            for _ in range(int(rng.integers(1, 5))):
                xb.insert(int(rng.integers(0, len(xb) + 1)), int(rng.choice(list(b"N-n-"))))
        ctx.set_option("kmer", int(rng.integers(1, 6)))
        ctx.set_option("chunk", int(rng.choice([1, 2, 3, 7, 16, 64, 2048])))
        ctx.set_option("min_cap", int(rng.choice([1, 3, 64])))
        ctx.build_index(d)
        _check_rows(ctx, d, [bytes(xb), bytes(xb)[::-1], b"", b"----", bytes(xb) + b"\n"])
    ctx.set_option("kmer", 0)
    ctx.set_option("chunk", 0)
    ctx.set_option("min_cap", 0)


@pytest.mark.parametrize("blind", [0, 1, 2, 6])
def test_fix_rounds_blind_or_not(ctx, blind):
    """The fix-up rounds are launched without a read-back in between (option blind_rounds, default 2) and one
    read-back afterwards decides whether more are needed; with tiny chunks many rounds are needed, so every mix of
    blind and checked rounds -- and the redone marking after an insufficient blind guess -- must give the same parse."""
    rng = np.random.default_rng(4242)
    ctx.set_option("blind_rounds", blind)
    try:
        for it in range(25):
            d, x = random_case(rng, b"ACGT", 300, 600)
            ctx.set_option("chunk", int(rng.choice([2, 4, 16, 64])))
            ctx.set_option("kmer", int(rng.integers(1, 6)))
            ctx.build_index(d)
            _check_rows(ctx, d, [x, x[::-1], d, x + x])
            assert ctx.parse_timings()["fix_rounds"] >= 1
    finally:
        ctx.set_option("blind_rounds", 2); ctx.set_option("chunk", 0); ctx.set_option("kmer", 0)


def test_repeated_batches_reuse_buffers(ctx):
    """steady state of a streaming caller: the zero-fill behind a batch serves the next one, records are emitted into
    the previous batch's buffer; batches of different sizes and shapes in a row must not see each other's leftovers"""
    seed = orc.seed_for(1)
    d = orc.syn_reference(seed, 300_000)
    rows = [r.tobytes() for r in orc.syn_rows(seed, d, 0, 6, 2e-3, 2e-4, 2e-4)]
    ctx.build_index(d.tobytes())
    sa = orc.sa_build(d.tobytes())
    for batch in ([rows[1]], rows, [rows[3][:1000], rows[2]], rows[:2], [b"", rows[5]], rows):
        _check_rows(ctx, d.tobytes(), batch, sa=sa)


def test_parse_golden(ctx):
    for f in sorted(os.listdir(GOLD)):
        if not f.endswith(".npz"):
            continue
        z = np.load(os.path.join(GOLD, f))
        d = z["ref"].tobytes()
        for chunk in (16, 2048):
            ctx.set_option("chunk", chunk)
            ctx.build_index(d)
            rows = [z[f"row{h}"].tobytes() for h in range(int(z["nrows"]))]
            res, m, _ = ctx.parse_batch(rows)
            for h in range(len(rows)):
                assert orc.lz_bytes(res[h]) == z[f"lz{h}"].tobytes()
    ctx.set_option("chunk", 0)


def test_long_matches(ctx):
    rng = np.random.default_rng(3)
    d = bytes(rng.choice(list(b"ACGT"), size=50000).astype(np.uint8))
    d = d + d[10000:30000] + bytes(rng.choice(list(b"ACGT"), size=5000).astype(np.uint8))      # exact 20 kb repeat
    rows = [d, d[:25000] + b"T" + d[25000:], d[10000:30000] * 2, d[40000:] + d[:40000], b"N" + d[10:40000] + b"N" + d[40000:]]
    for chunk, min_cap, fcap in ((16, 1, 0), (64, 8, 0), (2048, 64, 0), (2048, 64, 100), (256, 64, 3000)):
        ctx.set_option("chunk", chunk); ctx.set_option("min_cap", min_cap); ctx.set_option("foreign_cap", fcap)
        ctx.build_index(d)
        _check_rows(ctx, d, rows)
    d2 = b"A" * 7000 + b"C"
    ctx.set_option("chunk", 512)
    ctx.build_index(d2)
    _check_rows(ctx, d2, [b"A" * 3000, b"A" * 7000, b"A" * 6999 + b"CA"])
    ctx.set_option("chunk", 0); ctx.set_option("min_cap", 0); ctx.set_option("foreign_cap", 0)


def test_parse_low_complexity_fallback(ctx):
    d = b"A" * 63 + b"C"
    ctx.set_option("chunk", 64)
    ctx.build_index(d)
    _check_rows(ctx, d, [b"A" * 5000, b"C" * 100 + b"A" * 3000])
    ctx.set_option("chunk", 0)


def test_parse_synth_medium(ctx):
    seed = orc.seed_for(1)
    d = orc.syn_reference(seed, 2_000_000, repeats=30)
    rows = orc.syn_rows(seed, d, 0, 6, 1e-3, 1e-4, 1e-4)
    ctx.build_index(d)
    sa = _sa = orc.sa_build(d)
    assert (ctx.index_get("sa").astype(np.int64) == sa).all()
    _check_rows(ctx, d.tobytes(), [r.tobytes() for r in rows], sa)
    t = ctx.parse_timings()
    assert t["launches"] > 0 and t["fix_rounds"] >= 1
    # several groups through the pipelined host path give identical results
    ctx.set_option("group_bytes", 3_000_000)
    _check_rows(ctx, d.tobytes(), [r.tobytes() for r in rows], sa)
    ctx.set_option("group_bytes", 0)


def _gapped_row(rng, d: np.ndarray, events):
    """MSA row over the reference d: events = sorted [(reference position, kind, length)], kind 'g' = a run of gap
    columns in front of that base, 'd' = that many reference bases deleted (written as '-'), 's' = SNP, 'n' = an N."""
    out, i = [], 0
    for pos, kind, ln in events:
        out.append(d[i:pos]); i = pos
        if kind == "g":
            out.append(np.full(ln, ord("-"), dtype=np.uint8))
        elif kind == "d":
            out.append(np.full(ln, ord("-"), dtype=np.uint8)); i = min(len(d), pos + ln)
        elif kind == "s":
            out.append(np.array([b"ACGT"[(b"ACGT".find(bytes([d[pos]])) + 1 + int(rng.integers(0, 3))) % 4]], dtype=np.uint8)); i = pos + 1
        else:
            out.append(np.array([ord("#") if ord("N") in d[:4096] else ord("N")], dtype=np.uint8)); i = pos + 1    # a byte outside the alphabet
    out.append(d[i:])
    return np.concatenate(out)


@pytest.mark.parametrize("alphabet", [b"ACGT", b"ACGTN", bytes(range(60, 100))])
def test_pack_gap_patterns(ctx, alphabet):
    """k_pack* (2-bit) and k_packg1 (4- and 8-bit symbols): gaps, deletions and foreign bytes at and around the borders
    of the 64-byte thread spans, the 2 KB warp spans and the 16 KB tiles, gap runs that empty whole warps and whole
    tiles, rows of different lengths in one batch (tile ids interleave the rows), every alignment of the packed output
    against the words."""
    rng = np.random.default_rng(77)
    d = orc.syn_reference(orc.seed_for(3), 70_000)
    if alphabet != b"ACGT":
        d = rng.choice(list(alphabet), size=70_000).astype(np.uint8)
        if alphabet == b"ACGTN":
            d[30_000:30_900] = ord("N")                            # a run of N across a tile border (tile 2 starts at 32768 - gaps)
    ctx.build_index(d)
    sa = orc.sa_build(d)
    rows = []
    for r in range(10):
        ev, used = [], set()
        marks = [2048 * k + o for k in range(1, 34) for o in (-33, -32, -17, -16, -1, 0, 1, 15, 16, 31, 32)]
        for pos in rng.choice(marks, size=14, replace=False):
            pos = int(pos)
            if pos <= 0 or pos >= len(d) - 9 or any(abs(pos - u) < 12 for u in used):
                continue
            used.add(pos)
            kind = "gdsn"[int(rng.integers(0, 4))]
            ln = int(rng.choice([1, 2, 3, 7, 15, 16, 17, 31, 33])) if kind == "g" else int(rng.integers(1, 9))
            ev.append((pos, kind, ln))
        if r % 3 == 0:       # long gap runs: a whole warp span, more than a tile, and one that starts mid-word
            for pos, ln in ((5000 + r, 2048), (21000 + 3 * r, 16384 + 2048 + 5), (40001, 4099)):
                if not any(abs(pos - u) < 12 for u in used):
                    used.add(pos); ev.append((pos, "g", ln))
        ev.sort()
        row = _gapped_row(rng, d, ev)
        if r % 4 == 1:
            row = row[: int(rng.integers(16384, len(row)))]        # rows of different lengths, ending anywhere in a tile
        if r == 7:
            row = np.concatenate([np.full(16384 + 77, ord("-"), dtype=np.uint8), row])     # leading empty tile
        rows.append(row.tobytes())
    rows.append(b"-" * 40000)                                      # nothing but gaps: an empty haplotype among the others
    _check_rows(ctx, d.tobytes(), rows, sa)
    _check_rows(ctx, d.tobytes(), rows[3:5], sa)                   # the same rows at other positions of the tile order


def test_divergent_haplotype(ctx):
    seed = orc.seed_for(5)
    d = orc.syn_reference(seed, 300_000)
    rows = orc.syn_rows(seed, d, 0, 3, 8e-3, 1e-3, 1e-3)
    rnd = np.random.default_rng(3).choice(list(b"ACGT"), size=50_000).astype(np.uint8)   # unrelated sequence
    ctx.build_index(d)
    _check_rows(ctx, d.tobytes(), [r.tobytes() for r in rows] + [rnd.tobytes()])


def test_pairs_cap_too_small(ctx):
    d = orc.syn_reference(orc.seed_for(1), 50_000)
    rows = orc.syn_rows(orc.seed_for(1), d, 0, 2, 1e-2, 0, 0)
    ctx.build_index(d)
    with pytest.raises(D.DrlzError) as e:
        ctx.parse_batch([r.tobytes() for r in rows], pairs_cap=3)
    assert e.value.code == -4


def test_submit_wait_and_threads(ctx):
    import threading
    seed = orc.seed_for(1)
    d = orc.syn_reference(seed, 300_000)
    rows = orc.syn_rows(seed, d, 0, 6, 1e-3, 1e-4, 1e-4)
    ctx.build_index(d)
    sa = orc.sa_build(d)
    expect = [orc.lz_bytes(orc.encode_literal(d, sa, orc.strip_gaps(r))[0]) for r in rows]
    handles = [ctx.submit([rows[h].tobytes(), rows[(h + 1) % 6].tobytes()], pairs_cap=20000) for h in range(6)]
    for h, hd in reversed(list(enumerate(handles))):
        res, m = hd.wait()
        assert orc.lz_bytes(res[0]) == expect[h] and orc.lz_bytes(res[1]) == expect[(h + 1) % 6]
    # concurrent callers on one context are serialised, results stay correct
    out = {}

    def work(h):
        out[h] = ctx.parse_batch([rows[h].tobytes()])[0][0]
    ths = [threading.Thread(target=work, args=(h,)) for h in range(6)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    for h in range(6):
        assert orc.lz_bytes(out[h]) == expect[h]


def test_errors(ctx):
    with pytest.raises(D.DrlzError) as e:
        ctx.build_index(bytes(range(256)) * 2)          # every byte value in use: no code left for "not in reference"
    assert e.value.code == -5
    c2 = D.Drlz(0)
    with pytest.raises(D.DrlzError) as e:
        c2.parse_batch([b"ACGT"])
    assert e.value.code == -6
    c2.close()


@pytest.mark.parametrize("alphabet", [b"ACGTN", b"ACGNTacgnt", bytes(range(50, 100)), b"AB", bytes(range(40, 90)), b"\x00\x01\xfe\xff!"])
def test_generic_alphabet_parse(ctx, alphabet):
    """references outside {A,C,G,T}: 4-bit and 8-bit symbol layouts against the byte-wise oracle"""
    rng = np.random.default_rng(len(alphabet) + 100)
    for it in range(6):
        n = int(rng.choice([50, 300, 5000, 70000]))
        d = bytes(rng.choice(list(alphabet), size=n).astype(np.uint8))
        if it % 2:
            d = d + d[n // 4: n // 2] + alphabet[:1] * 300 + d[: n // 3]
        rows = []
        for r in range(5):
            x = bytearray(d[int(rng.integers(0, max(1, len(d) // 2))):])
            for _ in range(int(rng.integers(0, 30))):
                p = int(rng.integers(0, len(x) + 1)); op = int(rng.integers(0, 4))
                if op == 0 and p < len(x): x[p] = int(rng.choice(list(alphabet)))
                elif op == 1: x.insert(p, int(rng.choice(list(alphabet + b"-#"))))
                elif op == 2 and p < len(x): del x[p:p + int(rng.integers(1, 9))]
                else: x[p:p] = b"-" * int(rng.integers(1, 40))
            rows.append(bytes(x))
        rows += [d, b"", b"-", d[::-1]]
        ctx.set_option("kmer", int(rng.integers(0, 4)))
        ctx.set_option("chunk", int(rng.choice([64, 1024, 8192])))
        ctx.build_index(d)
        info = ctx.index_info()
        assert info["sigma"] == len(set(d)) and info["bits"] == (4 if len(set(d)) <= 16 else 8)
        _check_rows(ctx, d, rows, clean=min(alphabet) <= ord("1"))
    ctx.set_option("kmer", 0); ctx.set_option("chunk", 0)


def test_generic_alphabet_genome_like(ctx):
    """ACGT reference with N runs and soft-masked (lower-case) stretches, 2 Mbp, haplotypes with SNPs/indels/gaps"""
    seed = orc.seed_for(1)
    d = bytearray(orc.syn_reference(seed, 2_000_000).tobytes())
    rng = np.random.default_rng(5)
    for _ in range(5):
        p = int(rng.integers(0, len(d) - 60000)); L = int(rng.integers(1000, 50000)); d[p:p + L] = b"N" * L
    for _ in range(8):
        p = int(rng.integers(0, len(d) - 60000)); L = int(rng.integers(100, 20000)); d[p:p + L] = bytes(d[p:p + L]).lower()
    d = bytes(d)
    rows = [r.tobytes() for r in orc.syn_rows(seed, np.frombuffer(d, dtype=np.uint8), 0, 6, 1e-3, 1e-4, 1e-4)]
    ctx.build_index(d)
    assert ctx.index_info()["bits"] == 4
    _check_rows(ctx, d, rows)
    ms, pos = ctx.matching_stats(rows[1][:3000].replace(b"-", b""))
    sa = orc.sa_build(d); lcp = orc.lcp_build(d, sa)
    ms_ref, pos_ref = orc.ms_all(d, sa, lcp, rows[1][:3000].replace(b"-", b""))
    assert (ms.astype(np.int64) == ms_ref).all() and (pos.astype(np.int64) == pos_ref).all()


def test_dense_matching_stats(ctx):
    rng = np.random.default_rng(11)
    d = orc.syn_reference(orc.seed_for(2), 30_000, repeats=0).tobytes()
    x = bytearray(d[5000:9000])
    for p in rng.integers(0, len(x), size=40):
        x[p] = ord("ACGT"[int(rng.integers(0, 4))])
    ctx.build_index(d)
    ms, pos = ctx.matching_stats(bytes(x))
    sa = orc.sa_build(d)
    lcp = orc.lcp_build(d, sa)
    ms_ref, pos_ref = orc.ms_all(d, sa, lcp, bytes(x))
    assert (ms.astype(np.int64) == ms_ref).all()
    assert (pos.astype(np.int64) == pos_ref).all()


def test_device_resident_rows(ctx):
    import torch
    seed = orc.seed_for(1)
    d = orc.syn_reference(seed, 400_000)
    rows = orc.syn_rows(seed, d, 0, 5, 1e-3, 1e-4, 1e-4)
    H, M = rows.shape
    stride = (M + 15) // 16 * 16 + 16
    buf = torch.zeros(H * stride, dtype=torch.uint8, device="cuda")
    for h in range(H):
        buf[h * stride: h * stride + M] = torch.from_numpy(rows[h]).cuda()
    torch.cuda.synchronize()
    ctx.build_index(d)
    sa = orc.sa_build(d)
    ptr, npairs, m = ctx.parse_batch_device(buf.data_ptr(), np.arange(H) * stride, np.full(H, M))
    total = int(npairs.sum())
    allp = ctx.copy_pairs_to_host(ptr, total)
    o = 0
    for h in range(H):
        a1, _ = orc.encode_literal(d, sa, orc.strip_gaps(rows[h]))
        assert orc.lz_bytes(allp[o:o + int(npairs[h])]) == orc.lz_bytes(a1)
        o += int(npairs[h])
    res, m2, _ = ctx.parse_batch([r.tobytes() for r in rows])
    assert (m == m2).all()
    assert [r.shape[0] for r in res] == npairs.tolist()
    for h in range(H):
        a1, _ = orc.encode_literal(d, sa, orc.strip_gaps(rows[h]))
        assert orc.lz_bytes(res[h]) == orc.lz_bytes(a1)
