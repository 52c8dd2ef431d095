"""The chunk-parallel parse logic of rlz_core.cuh (run on the host by the test-only harness tests/emu)
must reproduce the oracle's literal restatement of DistributedRLZ.scala bit for bit."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests.emu import emu
from tests.util import random_case


def check(d, rows, k, chunk, min_cap=64, foreign_cap=1 << 20, use_uniq=1, clean=False):
    sa = orc.sa_build(d)
    lcp = orc.lcp_build(d, sa) if clean else None
    rc, res, m, rounds = emu.parse(d, sa, rows, k, chunk, min_cap, foreign_cap, use_uniq)
    assert rc == 0, (rc, rounds)
    for h, r in enumerate(rows):
        x = orc.strip_gaps(r).tobytes()
        a1 = orc.encode_a3(d, sa, lcp, x) if clean else orc.encode_literal(d, sa, x)[0]
        assert m[h] == len(x)
        assert res[h].tolist() == a1.tolist(), (d, r, k, chunk)
    return rounds


@pytest.mark.parametrize("seed", range(4))
def test_random_small(seed):
    rng = np.random.default_rng(seed)
    for it in range(400):
        d, x = random_case(rng, b"ACGT", 60, 80)
        if rng.random() < 0.3:   # haplotype bytes outside the reference alphabet and gaps
            xb = bytearray(x)
            for _ in range(int(rng.integers(1, 4))):
                xb.insert(int(rng.integers(0, len(xb) + 1)), int(rng.choice(list(b"N-n"))))
            x = bytes(xb)
        k = int(rng.integers(1, 6))
        chunk = int(rng.choice([1, 2, 3, 5, 8, 16, 64]))
        min_cap = int(rng.choice([1, 2, 4, 64]))
        check(d, [x, x[::-1], b"", d, d[3:]], k, chunk, min_cap, use_uniq=int(rng.integers(0, 2)))


@pytest.mark.parametrize("alphabet", [b"ACGTN", b"ACGTNacgtn", b"AB", bytes(range(60, 100))])
def test_generic_alphabets(alphabet):
    """references outside {A,C,G,T}: 4-bit (sigma <= 16) and 8-bit symbol paths, order-preserving codes"""
    rng = np.random.default_rng(len(alphabet))
    for it in range(150):
        d, x = random_case(rng, alphabet, 120, 200)
        xb = bytearray(x)
        if rng.random() < 0.5:
            for _ in range(int(rng.integers(1, 4))):
                xb.insert(int(rng.integers(0, len(xb) + 1)), int(rng.choice(list(b"-#z"))))
        k = int(rng.integers(1, 5))
        chunk = int(rng.choice([1, 4, 16, 32, 64, 256]))
        check(d, [bytes(xb), bytes(xb)[::-1], d, b""], k, chunk, int(rng.choice([1, 4, 64])))
    # N runs: long single-symbol stretches in the reference and the haplotype
    d = b"ACGTTGCA" * 50 + b"N" * 700 + b"GATTACA" * 40 + b"N" * 90 + b"ACGT" * 30
    x = d[:300] + b"N" * 650 + d[1100:] + b"NNNNACGTNN"
    check(d, [x, d, x[::-1]], 3, 64)
    check(d, [x, d], 3, 1024)


@pytest.mark.parametrize("alphabet", [bytes(range(40, 60)), b"\x00\x01\xfe\xff!", b"01"])
def test_alphabets_below_the_reference_sentinel(alphabet):
    """Reference bytes <= '1': DistributedRLZ.scala:118-122 compares positions past the end of the text as the
    byte '1', which contradicts the SA order (shorter suffix first) for such bytes, so the reference's result
    depends on the search trajectory there.  The product follows the clean specification (longest match, leftmost
    SA index = A3 = brute force), pinned here; test_oracle pins that A1 == A3 whenever all bytes are > '1'."""
    from tests.util import brute_parse
    rng = np.random.default_rng(len(alphabet) + 7)
    for it in range(120):
        d, x = random_case(rng, alphabet, 100, 160)
        k = int(rng.integers(1, 4))
        chunk = int(rng.choice([1, 8, 64, 256]))
        check(d, [x, x[::-1], d], k, chunk, int(rng.choice([1, 64])), clean=True)
        if it < 25:
            sa = orc.sa_build(d)
            xs = orc.strip_gaps(x).tobytes()
            a3 = orc.encode_a3(d, sa, orc.lcp_build(d, sa), xs)
            assert a3.tolist() == brute_parse(d, xs).tolist()


def test_bracket_start_probe_with_imitating_short_suffixes():
    """lookup(): the leftmost suffix that shares L < k symbols with the pattern is the first entry of the table
    bracket of that L-prefix, unless suffixes shorter than L sit at the start of the bracket because their zero
    padding imitates the prefix (reference ends in A's = code 0).  The probe must then fall back to the search."""
    rng = np.random.default_rng(5)
    for it in range(300):
        n = int(rng.integers(4, 40))
        d = bytes(rng.choice(list(b"AC"), size=n).astype(np.uint8)) + b"A" * int(rng.integers(1, 7))
        rows = []
        for _ in range(4):
            m = int(rng.integers(1, 30))
            x = bytearray(rng.choice(list(b"AC"), size=m).astype(np.uint8))
            x += b"A" * int(rng.integers(0, 6)) + bytes(rng.choice(list(b"ACG"), size=int(rng.integers(0, 4))).astype(np.uint8))
            rows.append(bytes(x))
        check(d, rows, int(rng.integers(3, 9)), int(rng.choice([4, 64])), int(rng.choice([1, 64])), use_uniq=int(rng.integers(0, 2)))


def test_low_complexity_never_merging():
    # parses started at different offsets of a homopolymer never share a phrase start:
    # the fix-up must still converge (one extra record per round) or report overflow, never be wrong
    d = b"A" * 63 + b"C"
    x = b"A" * 1000
    sa = orc.sa_build(d)
    rc, res, m, rounds = emu.parse(d, sa, [x], 3, 256)
    a1, _ = orc.encode_literal(d, sa, x)
    if rc == 0:
        assert res[0].tolist() == a1.tolist()
    else:
        assert rc == 2


def test_long_matches_are_chained():
    # haplotype == reference (one phrase), long shared stretches, an exact long repeat in the reference
    rng = np.random.default_rng(3)
    d = bytes(rng.choice(list(b"ACGT"), size=5000).astype(np.uint8))
    d = d + d[1000:3000] + bytes(rng.choice(list(b"ACGT"), size=500).astype(np.uint8))      # exact 2 kb repeat
    rows = [d, d[:2500] + b"T" + d[2500:], d[1000:3000] * 2, d[4000:] + d[:4000], b"N" + d[10:4000] + b"N" + d[4000:]]
    for chunk, min_cap in ((16, 1), (16, 8), (64, 64), (256, 64), (2048, 64)):
        check(d, rows, 5, chunk, min_cap)
    # homopolymers: every suffix ties, the diagonal is still consistent
    check(b"A" * 700 + b"C", [b"A" * 300, b"A" * 700, b"A" * 699 + b"CA"], 3, 32, 4)


@pytest.mark.parametrize("chunk", [32, 64, 128, 1024])
def test_diag_scan_paths(chunk):
    # chunk sizes >= 32 enable the hinted-diagonal pre-scan; dense edits overflow its 14-entry lists,
    # repeats make the uniqueness test fail, N's cut phrases: every fall-through path must stay exact
    seed = orc.seed_for(2)
    d = orc.syn_reference(seed, 60_000, repeats=12)
    rows = orc.syn_rows(seed, d, 0, 4, 1.5e-2, 2e-3, 2e-3)
    extra = bytearray(rows[1].tobytes())
    for p in range(500, len(extra), 7000):
        extra[p] = ord("N")
    dense = bytearray(rows[2].tobytes())
    rng = np.random.default_rng(chunk)
    for p in rng.integers(3000, 6000, size=400):
        dense[p] = ord("ACGT"[int(rng.integers(0, 4))])
    check(d.tobytes(), [r.tobytes() for r in rows] + [bytes(extra), bytes(dense), d.tobytes()[100:30000]], 7, chunk)


def test_synth_medium():
    seed = orc.seed_for(1)
    d = orc.syn_reference(seed, 200_000, repeats=20)
    rows = orc.syn_rows(seed, d, 0, 4, 2e-3, 1e-4, 1e-4)
    rounds = check(d.tobytes(), [r.tobytes() for r in rows], 8, 512)
    assert rounds <= 4


def test_golden():
    import os
    gold = os.path.join(os.path.dirname(__file__), "golden")
    for f in sorted(os.listdir(gold)):
        if not f.endswith(".npz"):
            continue
        z = np.load(os.path.join(gold, f))
        d = z["ref"].tobytes()
        rows = [z[f"row{h}"].tobytes() for h in range(int(z["nrows"]))]
        sa = z["sa"]
        rc, res, m, rounds = emu.parse(d, sa, rows, 4, 16)
        assert rc == 0
        for h in range(len(rows)):
            assert orc.lz_bytes(res[h]) == z[f"lz{h}"].tobytes()


@pytest.mark.parametrize("seed", range(3))
def test_short_lookups_medium_reference(seed):
    """lookup_short (the five-load form of the lookup on a variant site) answers only where the table brackets are
    small, i.e. with a jump table about as large as the reference -- as in production (k = 13 for 48 Mbp), not in the
    tiny cases above: SNPs, indels, repeats and a suffix of the text that is shorter than the matched prefix"""
    rng = np.random.default_rng(500 + seed)
    n = 30_000
    d = orc.syn_reference(orc.seed_for(1) + seed, n, repeats=0).tobytes()
    if seed == 1:
        d = d[:9000] + d[2000:5000] + d[9000:20000] + d[2500:4000] + d[20000:]
    if seed == 2:
        d = d + d[100:109]                      # the text ends in a copy of an earlier 9-mer: imitators in the k-mer brackets
    sa = orc.sa_build(d)
    rows = []
    for h in range(4):
        x = bytearray(d[int(rng.integers(0, 500)):])
        for _ in range(int(len(x) * (2e-3 if h < 3 else 2e-2))):
            p = int(rng.integers(0, len(x))); x[p] = b"ACGT"[(b"ACGT".index(x[p:p + 1]) + 1 + int(rng.integers(0, 3))) % 4]
        for _ in range(6 if h else 0):
            p = int(rng.integers(0, len(x)))
            if rng.random() < 0.5: del x[p:p + int(rng.integers(1, 9))]
            else: x[p:p] = bytes(rng.choice(list(b"ACGT"), size=int(rng.integers(1, 9))).astype(np.uint8))
        rows.append(bytes(x))
    rows.append(d)
    for k, chunk, min_cap in ((8, 256, 64), (9, 1024, 64), (10, 512, 32), (7, 128, 64)):
        rc, res, m, rounds = emu.parse(d, sa, rows, k, chunk, min_cap, 1 << 20, 1)
        assert rc == 0
        for h, r in enumerate(rows):
            assert res[h].tolist() == orc.encode_literal(d, sa, r)[0].tolist(), (seed, k, chunk, h)


def test_long_repeats_hint_fill_and_extents():
    """Runs far longer than a chunk (N runs, poly-A): chunk starts inside the run have no unique 64-symbol pattern; with
    use_uniq = 2 they inherit the diagonal of the chunk behind them (else in front), the chunk that sees the run's end
    takes its diagonal from a probe behind the run, and the capped-but-not-unique matches take their real
    length from the extents across chunks (ext[]) instead of the uncapped second pass of the lookup.  Same records as the
    oracle either way, for aligned runs, a longer and a shorter run in the haplotype, and a SNP right behind the run."""
    rng = np.random.default_rng(77)
    left = bytes(rng.choice(list(b"ACGT"), size=3000).astype(np.uint8))
    right = bytes(rng.choice(list(b"ACGT"), size=4000).astype(np.uint8))
    for sym in (b"N", b"A"):
        d = left + sym * 5000 + right + sym * 700 + left[:500]
        sa = orc.sa_build(d)
        x0 = bytearray(d)
        for p in (100, 2500, 8100, 8101, 9000, 12650):
            x0[p] = b"ACGT"[(b"ACGT".index(bytes(x0[p:p + 1])) + 1) % 4] if x0[p:p + 1] in (b"A", b"C", b"G", b"T") else x0[p]
        x1 = bytearray(d); x1[8000] = ord("C") if d[8000:8001] != b"C" else ord("G")        # SNP on the first base behind the run
        # variants at the edges of a run that is otherwise the reference's: a SNP ten symbols behind it (the probe at the
        # run's end stands on it), deletions shortly in front of the run and shortly behind it (the neighbouring chunks'
        # diagonals then differ from the one the inside of the run needs), a deletion inside the run
        x2 = bytearray(d); x2[8010] = ord("C") if d[8010:8011] != b"C" else ord("G")
        x3 = bytes(x2[:2900] + x2[2907:8300] + x2[8305:])
        x4 = d[:5000] + d[5009:]
        rows = [bytes(x0), bytes(x1), d, left + sym * 6000 + right, left + sym * 3500 + right, sym * 2000 + right[:100], bytes(x2), x3, x4]
        for k, chunk, min_cap in ((5, 256, 64), (6, 1024, 64), (4, 64, 32), (6, 4096, 64)):
            for mode in (1, 2):
                before = emu.long_hits()
                rc, res, m, _ = emu.parse(d, sa, rows, k, chunk, min_cap, 1 << 20, mode)
                assert rc == 0, (sym, k, chunk, mode, rc)
                hits = emu.long_hits() - before                  # phrases taken from the extents across chunks
                assert hits == 0 if mode == 1 else (hits > 0 or chunk >= 4096), (sym, k, chunk, mode, hits)
                for h, r in enumerate(rows):
                    assert res[h].tolist() == orc.encode_literal(d, sa, r)[0].tolist(), (sym, k, chunk, mode, h)
