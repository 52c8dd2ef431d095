"""Host-side logic of the drop-in stage (file naming, ordering, sharding, record layout) on CPU.
The parser is replaced by the oracle here -- these tests check plumbing, not the CUDA path."""
import os
import subprocess
import sys

import numpy as np

from oracle import oracle as orc
from pangenspark_b200 import distributed_rlz as drz

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def oracle_parser(ref, rows):
    sa = orc.sa_build(ref)
    pairs, gapless = [], []
    for r in rows:
        x = orc.strip_gaps(r).tobytes()
        pairs.append(orc.encode_literal(ref, sa, x)[0])
        gapless.append(x)
    return pairs, gapless


def make_dataset(tmp, H=7, n=3000):
    seed = orc.seed_for(1)
    d = orc.syn_reference(seed, n)
    rows = orc.syn_rows(seed, d, 0, H, 1e-2, 2e-3, 2e-3)
    os.makedirs(os.path.join(tmp, "msa"))
    with open(os.path.join(tmp, "ref.txt"), "wb") as f:
        f.write(d.tobytes())
    for h in range(H):
        with open(os.path.join(tmp, "msa", "H%05d.01" % h), "wb") as f:
            f.write(rows[h].tobytes() + b"\n")
    return d, rows


def check_outputs(tmp, d, rows, out="drlz", gap="gapless"):
    sa = orc.sa_build(d)
    parts = sorted(f for f in os.listdir(os.path.join(tmp, gap)) if f.startswith("part-"))
    assert len(parts) == rows.shape[0] and os.path.exists(os.path.join(tmp, gap, "_SUCCESS"))
    fasta = b"".join(open(os.path.join(tmp, gap, p), "rb").read() for p in parts)
    expect = b""
    for h in range(rows.shape[0]):
        x = orc.strip_gaps(rows[h]).tobytes()
        lz = open(os.path.join(tmp, out, "H%05d.01.lz" % h), "rb").read()
        assert lz == orc.lz_bytes(orc.encode_literal(d, sa, x)[0])
        assert len(lz) % 16 == 0
        expect += b">H%05d.01\n" % h + x + b"\n"
    assert fasta == expect            # getmerge order == sorted basenames (index_chr.sh:16,20)


def test_single_rank_layout(tmp_path):
    tmp = str(tmp_path)
    d, rows = make_dataset(tmp)
    rc = drz.main([tmp + "/radix", tmp + "/ref.txt", tmp + "/msa/*.01", "file:///", tmp + "/drlz", tmp + "/gapless"],
                  parser=oracle_parser, rank=0, world=1, batch_bytes=5000)
    assert rc == 0
    check_outputs(tmp, d, rows)


def test_merged_outputs(tmp_path, monkeypatch):
    tmp = str(tmp_path)
    d, rows = make_dataset(tmp, H=4)
    monkeypatch.setenv("DRLZ_MERGED_LZ", tmp + "/chr01.lz")
    monkeypatch.setenv("DRLZ_MERGED_FA", tmp + "/chr01.fa")
    rc = drz.main([tmp + "/radix", tmp + "/ref.txt", tmp + "/msa/*.01", "", tmp + "/drlz", tmp + "/gapless"],
                  parser=oracle_parser, rank=0, world=1)
    assert rc == 0
    lz = b"".join(open(tmp + "/drlz/H%05d.01.lz" % h, "rb").read() for h in range(4))
    assert open(tmp + "/chr01.lz", "rb").read() == lz
    fa = open(tmp + "/chr01.fa", "rb").read()
    assert fa.count(b">") == 4 and fa.startswith(b">H00000.01\n")


def test_gzip_rows(tmp_path):
    """legacy inputs are gzip files (create.sh:27 globs *.NN.gz); the output keeps the input's full basename + .lz
    (DistributedRLZ.scala:257)"""
    import gzip
    tmp = str(tmp_path)
    d, rows = make_dataset(tmp, H=3)
    for h in range(3):
        src = os.path.join(tmp, "msa", "H%05d.01" % h)
        with open(src, "rb") as f, gzip.open(src + ".gz", "wb") as g:
            g.write(f.read())
        os.remove(src)
    rc = drz.main([tmp + "/radix", tmp + "/ref.txt", tmp + "/msa/*.01.gz", "", tmp + "/drlz", tmp + "/gapless"],
                  parser=oracle_parser, rank=0, world=1)
    assert rc == 0
    sa = orc.sa_build(d)
    for h in range(3):
        x = orc.strip_gaps(rows[h]).tobytes()
        lz = open(os.path.join(tmp, "drlz", "H%05d.01.gz.lz" % h), "rb").read()
        assert lz == orc.lz_bytes(orc.encode_literal(d, sa, x)[0])
        assert open(os.path.join(tmp, "gapless", "part-%05d" % h), "rb").read() == b">H%05d.01.gz\n" % h + x + b"\n"


def test_shards_cover_everything():
    for n in range(0, 40):
        for w in (1, 2, 3, 8):
            got = sum((drz.shard(list(range(n)), r, w) for r in range(w)), [])
            assert got == list(range(n))


def test_multiline_row_rejected(tmp_path):
    p = tmp_path / "x.01"
    p.write_bytes(b"ACGT\nACGT\n")
    try:
        drz.read_row(str(p))
        assert False
    except ValueError:
        pass
    assert drz.read_reference(str(p)) == b"ACGTACGT"


WORKER = r'''
import os, sys
sys.path.insert(0, {root!r})
import torch.distributed as dist
from pangenspark_b200 import distributed_rlz as drz
from tests.test_host_logic import oracle_parser
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
tmp = {tmp!r}
rc = drz.main([tmp + "/radix", tmp + "/ref.txt", tmp + "/msa/*.01", "", tmp + "/drlz", tmp + "/gapless"],
              parser=oracle_parser, rank=dist.get_rank(), world=dist.get_world_size())
dist.barrier()
dist.destroy_process_group()
sys.exit(rc)
'''


def test_world_size_2_gloo(tmp_path):
    """Two ranks (gloo, CPU) shard the haplotypes; the union of their outputs equals the single-rank result."""
    tmp = str(tmp_path)
    d, rows = make_dataset(tmp, H=9)
    script = os.path.join(tmp, "worker.py")
    import random
    port = 29500 + random.randint(0, 2000)
    with open(script, "w") as f:
        f.write(WORKER.format(root=ROOT, port=port, tmp=tmp))
    procs = [subprocess.Popen([sys.executable, script, str(r)]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    check_outputs(tmp, d, rows)


def test_pack_classification_arithmetic():
    """The SIMD-in-register ACGT test and code extraction of k_pack (pack.cu::codes4, top_bytes), restated in numpy
    with the constants read from the source, over every byte value in every position of a 32-bit word: a byte's lane
    of `bad` is zero exactly for A, C, G, T, valid bytes get the codes 0..3 in alphabet order, and the multiply gathers
    the four codes into the top byte, first base first."""
    import re
    src = open(os.path.join(ROOT, "pangenspark_b200", "csrc", "pack.cu")).read()
    body = src[src.index("uint32_t codes4(uint32_t w, uint32_t* bad)"):]
    body = body[:body.index("}")]
    for const in ("0x01010101u", "0xF9F9F9F9u", "15u", "0x41414141u", "0x03030303u"):
        assert const in body, const
    assert re.search(r"c\d \* 0x40100401u", src)
    u = np.uint32
    rng = np.random.default_rng(9)
    code = {65: 0, 67: 1, 71: 2, 84: 3}
    words = []
    for pos in range(4):
        for b in range(256):
            for _ in range(8):
                bs = [int(x) for x in (rng.choice([65, 67, 71, 84], 4) if rng.random() < 0.7 else rng.integers(0, 256, 4))]
                bs[pos] = b
                words.append(bs)
    W = np.array(words, dtype=np.uint64)
    w = (W[:, 0] | (W[:, 1] << 8) | (W[:, 2] << 16) | (W[:, 3] << 24)).astype(u)
    s1, s2 = w >> u(1), w >> u(2)
    t = s2 & ~s1 & u(0x01010101)
    bad = (w & u(0xF9F9F9F9)) ^ (t * u(15) + u(0x41414141))
    c = (s1 ^ (s2 & u(0x01010101))) & u(0x03030303)
    top = ((c.astype(np.uint64) * 0x40100401) & 0xFFFFFFFF) >> 24
    for row, bd, cc, tp in zip(words, bad.tolist(), c.tolist(), top.tolist()):
        for i, v in enumerate(row):
            assert (((bd >> (8 * i)) & 255) == 0) == (v in code), (row, i)
        if all(v in code for v in row):
            assert [(cc >> (8 * i)) & 3 for i in range(4)] == [code[v] for v in row]
            assert tp == (code[row[0]] << 6) | (code[row[1]] << 4) | (code[row[2]] << 2) | code[row[3]]
