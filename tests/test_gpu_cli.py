"""End-to-end drop-in check on the GPU box: the native `drlz` CLI (same six arguments as the Spark job) and the
Python mirror must write byte-identical .lz files and gapless FASTA parts, equal to the oracle's."""
import os
import subprocess
import sys

import pytest

from tests.test_host_logic import check_outputs, make_dataset

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_native_cli(tmp_path):
    tmp = str(tmp_path)
    d, rows = make_dataset(tmp, H=6, n=200_000)
    exe = os.path.join(ROOT, "pangenspark_b200", "bin", "drlz")
    env = dict(os.environ, DRLZ_EMIT_RADIX="1", DRLZ_BATCH_BYTES="500000", DRLZ_MERGED_LZ=tmp + "/chr01.lz",
               DRLZ_MERGED_FA=tmp + "/chr01.fa")
    r = subprocess.run([exe, tmp + "/radix", tmp + "/ref.txt", tmp + "/msa/*.01", "file:///", tmp + "/drlz", tmp + "/gapless", "2"],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "started encoding" in r.stdout
    check_outputs(tmp, d, rows)
    lz = b"".join(open(tmp + "/drlz/H%05d.01.lz" % h, "rb").read() for h in range(6))
    assert open(tmp + "/chr01.lz", "rb").read() == lz
    assert open(tmp + "/chr01.fa", "rb").read().count(b">") == 6
    from oracle import oracle as orc
    sa = [int(x) for x in open(tmp + "/radix").read().split()]
    assert sa == orc.sa_build(d).tolist()


def test_python_mirror(tmp_path):
    tmp = str(tmp_path)
    d, rows = make_dataset(tmp, H=5, n=50_000)
    r = subprocess.run([sys.executable, "-m", "pangenspark_b200.distributed_rlz", tmp + "/radix", tmp + "/ref.txt",
                        tmp + "/msa/*.01", "", tmp + "/drlz", tmp + "/gapless"], capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert r.returncode == 0, r.stderr
    check_outputs(tmp, d, rows)


def test_native_cli_gzip_rows(tmp_path):
    """gzip rows through the native CLI (zlib): same records as the plain files"""
    import gzip
    from oracle import oracle as orc
    tmp = str(tmp_path)
    d, rows = make_dataset(tmp, H=3, n=100_000)
    for h in range(3):
        src = os.path.join(tmp, "msa", "H%05d.01" % h)
        with open(src, "rb") as f, gzip.open(src + ".gz", "wb") as g:
            g.write(f.read())
        os.remove(src)
    exe = os.path.join(ROOT, "pangenspark_b200", "bin", "drlz")
    r = subprocess.run([exe, tmp + "/radix", tmp + "/ref.txt", tmp + "/msa/*.01.gz", "", tmp + "/drlz", tmp + "/gapless"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    sa = orc.sa_build(d)
    for h in range(3):
        x = orc.strip_gaps(rows[h]).tobytes()
        lz = open(os.path.join(tmp, "drlz", "H%05d.01.gz.lz" % h), "rb").read()
        assert lz == orc.lz_bytes(orc.encode_literal(d, sa, x)[0])


def test_native_cli_streaming_pipeline(tmp_path):
    """many small batches through few slots and few I/O threads: every batch boundary, slot reuse and the FIFO order of
    the asynchronous submissions are exercised; plus the gap table and the offset map of the merged FASTA"""
    import json
    import numpy as np
    from oracle import oracle as orc
    from pangenspark_b200 import distributed_rlz as drz
    tmp = str(tmp_path)
    d, rows = make_dataset(tmp, H=23, n=60_000)
    exe = os.path.join(ROOT, "pangenspark_b200", "bin", "drlz")
    for slots, threads, batch in (("2", "1", "70000"), ("3", "5", "200000"), ("4", "16", "1")):
        out, gap = tmp + "/drlz" + slots, tmp + "/gapless" + slots
        env = dict(os.environ, DRLZ_BATCH_BYTES=batch, DRLZ_SLOTS=slots, DRLZ_IO_THREADS=threads, DRLZ_GAPTABLE="1",
                   DRLZ_MERGED_FA=tmp + "/chr01.fa", DRLZ_MERGED_LZ=tmp + "/chr01.lz", DRLZ_STATS=tmp + "/stats.json")
        r = subprocess.run([exe, tmp + "/radix", tmp + "/ref.txt", tmp + "/msa/*.01", "file:///", out, gap],
                           capture_output=True, text=True, env=env, timeout=300)
        assert r.returncode == 0, r.stderr
        check_outputs(tmp, d, rows, out="drlz" + slots, gap="gapless" + slots)
        st = json.load(open(tmp + "/stats.json"))
        assert st["haplotypes"] == 23 and st["bases"] == sum(int(orc.strip_gaps(r_).size) for r_ in rows)
        assert st["batches"] >= (23 if batch == "1" else 2)
        for h in range(23):
            g = np.frombuffer(open(out + "/H%05d.01.gaps" % h, "rb").read(), dtype="<i8").reshape(-1, 2)
            assert (g == orc.gap_runs(rows[h])).all()
        offs = drz.read_offsets(tmp + "/chr01.fa.offsets")
        fa = open(tmp + "/chr01.fa", "rb").read()
        assert len(offs) == 23 and fa[offs[0][2]:offs[0][2] + offs[0][3]] == d.tobytes()
    # without the gapless FASTA (a caller that only wants the phrases)
    env = dict(os.environ, DRLZ_GAPLESS="0")
    r = subprocess.run([exe, tmp + "/radix", tmp + "/ref.txt", tmp + "/msa/*.01", "", tmp + "/only_lz", tmp + "/none"],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr
    assert not os.path.exists(tmp + "/none") and len(os.listdir(tmp + "/only_lz")) == 23


def test_native_cli_rejects_multiline_rows(tmp_path):
    tmp = str(tmp_path)
    d, rows = make_dataset(tmp, H=2, n=5000)
    with open(tmp + "/msa/H00001.01", "ab") as f:
        f.write(b"ACGT\n")
    exe = os.path.join(ROOT, "pangenspark_b200", "bin", "drlz")
    r = subprocess.run([exe, tmp + "/radix", tmp + "/ref.txt", tmp + "/msa/*.01", "", tmp + "/drlz", tmp + "/gapless"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 1 and "more than one line" in r.stderr


def test_native_cli_reference_with_line_terminators(tmp_path):
    """the reference file is the sequence bytes; a trailing newline (or CR LF, or a wrapped file) is dropped
    with a warning -- `getLines.mkString` in DistributedRLZ.scala:61 does the same to what the parse sees"""
    from oracle import oracle as orc
    tmp = str(tmp_path)
    d, rows = make_dataset(tmp, H=2, n=50_000)
    raw = bytes(d)
    sa = orc.sa_build(raw)
    exe = os.path.join(ROOT, "pangenspark_b200", "bin", "drlz")
    for k, text in enumerate((raw + b"\n", raw + b"\r\n", raw[:20_000] + b"\n" + raw[20_000:33_333] + b"\r\n" + raw[33_333:] + b"\n")):
        with open(tmp + "/ref.txt", "wb") as f:
            f.write(text)
        out = tmp + "/drlz%d" % k
        r = subprocess.run([exe, tmp + "/radix", tmp + "/ref.txt", tmp + "/msa/*.01", "", out, tmp + "/gapless%d" % k],
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        warned = "line terminators" in r.stderr
        assert not (k == 0 and warned) and not (k == 2 and not warned)      # one trailing newline is the normal case
        for h in range(2):
            x = orc.strip_gaps(rows[h]).tobytes()
            lz = open(os.path.join(out, "H%05d.01.lz" % h), "rb").read()
            assert lz == orc.lz_bytes(orc.encode_literal(raw, sa, x)[0])


def test_native_cli_hdfs_mode(tmp_path):
    from tests.test_stage_scripts import test_cli_hdfs_inputs_through_the_hdfs_command
    test_cli_hdfs_inputs_through_the_hdfs_command(tmp_path)


def test_two_contexts_in_one_process_count_their_own_launches():
    """DRLZ_GPUS > 1 runs several contexts from several threads of one process: the launch counter is per calling
    thread, and a ticket can be collected once"""
    import threading
    from oracle import oracle as orc
    from pangenspark_b200 import drlz as D
    seed = orc.seed_for(1)
    d = orc.syn_reference(seed, 200_000)
    rows = [r for r in orc.syn_rows(seed, d, 0, 3, 1e-3, 1e-4, 1e-4)]
    ctxs = [D.Drlz(0), D.Drlz(0)]
    for c in ctxs:
        c.build_index(d)
    counts = [None, None]

    def work(i):
        for _ in range(5):
            ctxs[i].parse_batch(rows)
        counts[i] = ctxs[i].parse_timings()["launches"]
    th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    [t.start() for t in th]; [t.join() for t in th]
    assert counts[0] == counts[1] and 10 <= counts[0] <= 40
    h = ctxs[0].submit(rows, pairs_cap=100_000)
    h.wait()
    with pytest.raises(D.DrlzError):
        h.wait()                               # collected already: E_ARG instead of blocking forever
    for c in ctxs:
        c.close()
