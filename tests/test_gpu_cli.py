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
