"""Parity at the shapes BASELINE.json names beyond configs[1], and of the multi-GPU index hand-over.  Run on the B200 box.

  * index built in one context, moved into a second one through drlz_index_reserve / buffers / commit (what the NCCL
    broadcast does on ranks >= 1, DistributedRLZ.scala:58,63), parsed there, compared with the oracle
  * configs[2] shape: chr1-sized 249 Mbp reference, rows compared with the oracle (CPU suffix array ~20 s)
  * configs[4] shape: positions beyond 2^31 (the reference's Int limit) at 1 % divergence; no CPU suffix
    array at that size, so the size-independent properties of tools/big_check.py: the suffix array is a permutation
    in suffix order (sampled), every phrase stream decodes to its haplotype, no phrase can be extended
  * references with N runs of more than 1 Mbp (real GRCh37 chromosomes): the 4-bit generic-alphabet path
  * the exception-list overflow retry
"""
import numpy as np
import pytest

from oracle import oracle as orc
from pangenspark_b200 import drlz as D

pytestmark = pytest.mark.gpu


class _Raw:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def _check_rows(ctx, d, sa, rows):
    res, m, _ = ctx.parse_batch(rows)
    for h, r in enumerate(rows):
        x = orc.strip_gaps(r)
        a1, _ = orc.encode_literal(d, sa, x)
        assert int(m[h]) == x.size
        assert orc.lz_bytes(res[h]) == orc.lz_bytes(a1), (h, res[h][:4], a1[:4])


@pytest.mark.parametrize("alphabet", ["acgt", "acgtn"])
def test_index_moves_between_contexts(alphabet):
    import torch
    seed = orc.seed_for(1)
    d = orc.syn_reference(seed, 1_500_000, repeats=8)
    if alphabet == "acgtn":
        d[400_000:430_000] = ord("N")
    rows = [r for r in orc.syn_rows(seed, d, 0, 4, 1e-3, 1e-4, 1e-4)]
    a, b = D.Drlz(0), D.Drlz(0)
    try:
        a.build_index(d)
        info = a.index_info()
        b.index_reserve(info["n"], info["kinfo"])
        src, dst = a.index_buffers(), b.index_buffers()
        assert [nb for _, nb in src] == [nb for _, nb in dst]
        for (ps, nb), (pd, _) in zip(src, dst):          # device-to-device copy: the stand-in for dist.broadcast
            assert ps != pd
            torch.as_tensor(_Raw(pd, nb), device="cuda").copy_(torch.as_tensor(_Raw(ps, nb), device="cuda"))
        torch.cuda.synchronize()
        b.index_commit()
        assert b.index_info()["kinfo"] == info["kinfo"] and b.index_info()["n"] == info["n"]
        a.close()                                          # the receiving context must not depend on the builder's memory
        sa = orc.sa_build(d)
        assert (b.index_get("sa").astype(np.int64) == sa).all()
        _check_rows(b, d, sa, rows)
    finally:
        a.close(); b.close()


def test_config3_shape_against_oracle():
    n, cfg = 249_000_000, 3
    seed = orc.seed_for(cfg)
    d = orc.syn_reference(seed, n)
    rows = orc.syn_rows(seed, d, 0, 3, 1e-3, 1e-5, 1e-5)
    ctx = D.Drlz(0)
    try:
        ctx.build_index(d)
        res, m, _ = ctx.parse_batch([r for r in rows])
        sa = orc.sa_build(d)
        assert (ctx.index_get("sa").astype(np.int64) == sa).all()
        cres, cm, wt, _ = orc.encode_rows_mt(d, sa, [r for r in rows], 0)
        for h in range(rows.shape[0]):
            assert int(m[h]) == int(cm[h])
            assert orc.lz_bytes(res[h]) == orc.lz_bytes(cres[h]), h
        assert res[0].shape[0] == 1 and int(res[0][0, 1]) == n          # row 0 is the reference: one phrase
    finally:
        ctx.close()


def test_positions_beyond_2_31():
    n = (1 << 31) + (1 << 26)
    _, _, p_snp, p_ins, p_del = orc.CONFIGS["cfg5"]
    seed = orc.seed_for(5)
    d = orc.syn_reference(seed, n)
    rows = orc.syn_rows(seed, d, 1, 1, p_snp, p_ins, p_del)
    ctx = D.Drlz(0)
    try:
        ctx.build_index(d)
        sa = ctx.index_get("sa")
        chk = np.zeros(n, dtype=np.uint8)
        chk[sa] = 1
        assert bool(chk.all()), "suffix array is not a permutation"
        del chk
        rng = np.random.default_rng(1)
        for j in rng.integers(1, n, size=20_000):
            a, b = int(sa[j - 1]), int(sa[j])
            L = 64
            x, y = d[a:a + L].tobytes(), d[b:b + L].tobytes()
            while x == y and a + L < n and b + L < n:
                a += L; b += L
                x, y = d[a:a + L].tobytes(), d[b:b + L].tobytes()
            assert x < y, (j, a, b)
        assert int(sa.max()) == n - 1 and int(sa.max()) > (1 << 31)
        del sa
        res, m, _ = ctx.parse_batch([rows[0]])
        x = orc.strip_gaps(rows[0])
        assert int(m[0]) == x.size
        assert orc.decode(d, res[0]) == x.tobytes()
        assert int(res[0][:, 0].max()) > (1 << 31)
        starts = np.concatenate([[0], np.cumsum(np.maximum(res[0][:, 1], 1))[:-1]])
        for t in rng.integers(0, res[0].shape[0], size=20_000):
            pos, ln, i = int(res[0][t, 0]), int(res[0][t, 1]), int(starts[t])
            if ln and i + ln < x.size and pos + ln < n:
                assert d[pos + ln] != x[i + ln], "phrase %d can be extended at its own position" % t
    finally:
        ctx.close()


def test_reference_with_long_n_runs():
    """GRCh37 chromosomes carry N runs of millions of bases: the reference is then not pure ACGT (4-bit symbols, radix-5
    table keys), uniq[] saturates inside the runs and the capped lookups there are never unique."""
    seed = orc.seed_for(2)
    d = orc.syn_reference(seed, 6_000_000)
    d[:300_000] = ord("N")                       # telomere
    d[2_000_000:3_200_000] = ord("N")            # centromere-sized run: 1.2 Mbp
    d[5_000_000:5_050_000] = ord("N")
    rows = [r for r in orc.syn_rows(seed, d, 0, 4, 1e-3, 1e-5, 1e-5)]
    x0 = orc.strip_gaps(rows[0])
    rows.append(np.concatenate([x0[:2_100_000], np.full(1_500_000, ord("N"), np.uint8), x0[3_100_000:]]))   # a longer run than any in the reference
    rows.append(np.concatenate([x0[:2_500_000], x0[2_700_000:]]))                                       # a shorter one
    # the runs as a panel called against the reference carries them (untouched), with variants right at their edges: a SNP
    # ten bases behind a run (the probe at the run's end stands on it), deletions in the chunk in front of a run and in
    # the chunk behind it (the neighbours' diagonals then differ from the one the run's inside needs)
    y = d.copy()
    for p in (3_200_010, 5_050_005, 300_003, 1_234_567):
        y[p] = ord("ACGT"["ACGT".index(chr(y[p])) ^ 1])
    keep = np.ones(y.size, dtype=bool)
    for a, b in ((1_999_800, 1_999_805), (3_200_300, 3_200_310), (4_999_000, 4_999_001), (299_000, 299_017)):
        keep[a:b] = False
    rows.append(y[keep])
    ctx = D.Drlz(0)
    try:
        ctx.build_index(d)
        info = ctx.index_info()
        assert info["bits"] == 4 and info["sigma"] == 5
        sa = orc.sa_build(d)
        assert (ctx.index_get("sa").astype(np.int64) == sa).all()
        _check_rows(ctx, d, sa, rows)
        # rows whose runs are the reference's must not fall into the uncapped search (one such lookup streams the run in
        # every step of its binary search: milliseconds each, 80 ms in the 50-row run of profiles/README.md item 42)
        _check_rows(ctx, d, sa, [rows[-1]])
        assert ctx.parse_timings()["spec_ms"] < 5.0
    finally:
        ctx.close()


def test_exception_list_overflow_retry():
    """more foreign bytes in a row than the exception list holds at first (cols / 256 + 256): the group is redone
    with the list sized from the exact counts (api.cu: flags & 1)"""
    seed = orc.seed_for(1)
    d = orc.syn_reference(seed, 400_000)
    rows = orc.syn_rows(seed, d, 0, 3, 1e-3, 1e-4, 1e-4)
    rng = np.random.default_rng(5)
    noisy = rows[1].copy()
    hit = rng.random(noisy.size) < 0.05                    # 20 000 foreign bytes against a list of ~1 800
    noisy[hit & (noisy != ord("-"))] = ord("N")
    ctx = D.Drlz(0)
    try:
        ctx.build_index(d)
        sa = orc.sa_build(d)
        _check_rows(ctx, d, sa, [rows[0], noisy, rows[2]])
        _check_rows(ctx, d, sa, [noisy])                   # and once more with the grown list
    finally:
        ctx.close()


def test_gap_table():
    """gap-position side table against the host restatement of ScoreMatrix.scala:91-93"""
    seed = orc.seed_for(1)
    d = orc.syn_reference(seed, 300_000)
    rows = [r.copy() for r in orc.syn_rows(seed, d, 0, 5, 1e-3, 3e-3, 3e-3)]
    rows[1][:40] = ord("-")                     # run at the row start
    rows[1][-3:] = ord("-")                     # and at its end
    rows[2][16384 - 5:16384 + 9] = ord("-")     # across a 16 KB tile border
    rows[2][2048 * 3 - 1:2048 * 3 + 1] = ord("-")
    rows[3][15:17] = ord("-")                   # across a 16-byte vector border
    rows[3][31] = ord("-"); rows[3][32] = ord("-"); rows[3][33] = ord("A"); rows[3][34] = ord("-")
    rows.append(np.full(5000, ord("-"), np.uint8))
    rows.append(np.frombuffer(b"ACGT", np.uint8))
    rows.append(np.zeros(0, np.uint8))
    ctx = D.Drlz(0)
    try:
        ctx.build_index(d)
        sa = orc.sa_build(d)
        res, m, gaps = ctx.parse_batch_gaps(rows)
        for h, r in enumerate(rows):
            want = orc.gap_runs(r)
            assert gaps[h].shape == want.shape and (gaps[h] == want).all(), (h, gaps[h][:4], want[:4])
            assert int(m[h]) + int(want[:, 1].sum()) == r.size
            a1, _ = orc.encode_literal(d, sa, orc.strip_gaps(r))
            assert orc.lz_bytes(res[h]) == orc.lz_bytes(a1)
        # too small a gaps_cap: required counts come back, second call succeeds (handled inside the binding)
        res2, m2, gaps2 = ctx.parse_batch_gaps(rows, gaps_cap=3)
        assert all((a == b).all() for a, b in zip(gaps, gaps2))
    finally:
        ctx.close()
