"""CPU checks of the stage scripts and of what the native CLI does before it needs a GPU: argument handling, the
empty-glob error, the hdfs:// input path (against a stand-in `hdfs` command), the batch script over chromosomes,
the CHIC offset map and the gap-table oracle."""
import os
import stat
import subprocess
import sys

import numpy as np

from oracle import oracle as orc
from pangenspark_b200 import distributed_rlz as drz
from tests.test_host_logic import make_dataset, oracle_parser

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "pangenspark_b200", "bin", "drlz")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def _script(path, body):
    with open(path, "w") as f:
        f.write(body)
    os.chmod(path, os.stat(path).st_mode | stat.S_IEXEC)


def test_cli_usage_and_empty_glob(tmp_path):
    from pangenspark_b200 import drlz as D
    D.build_library()
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "pangenspark_b200"), "-s", "bin/drlz"])
    tmp = str(tmp_path)
    assert subprocess.run([EXE], capture_output=True).returncode == 2
    (tmp_path / "ref.txt").write_bytes(b"ACGTACGT")
    r = subprocess.run([EXE, tmp + "/radix", tmp + "/ref.txt", tmp + "/msa/*.01", "file:///", tmp + "/out", tmp + "/gap"],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "no input matches" in r.stderr          # Spark: "Path does not exist"
    assert not os.path.exists(tmp + "/gap/_SUCCESS")
    # the Python mirror behaves the same
    assert drz.main([tmp + "/radix", tmp + "/ref.txt", tmp + "/msa/*.01", "", tmp + "/out", tmp + "/gap"], parser=oracle_parser,
                    rank=0, world=1) == 1


HDFS_STUB = r'''#!/usr/bin/env bash
# stand-in for the `hdfs` CLI: "hdfs://x/<path>" is the local directory $FAKE_HDFS/<path>
echo "$@" >> "$FAKE_HDFS/calls.log"
loc() { echo "$FAKE_HDFS/${1#hdfs://x/}"; }
[ "$1" = dfs ] || exit 2
shift
case "$1" in
  -ls) shift; [ "$1" = -C ] && shift
       n=0; for f in $(loc "$1"); do [ -e "$f" ] && { echo "hdfs://x/${f#$FAKE_HDFS/}"; n=1; }; done; [ $n = 1 ] ;;
  -get) shift; [ "$1" = -f ] && shift; cp "$(loc "$1")" "$2" ;;
  -mkdir) shift; [ "$1" = -p ] && shift; for d in "$@"; do mkdir -p "$(loc "$d")"; done ;;
  -put) shift; [ "$1" = -f ] && shift; args=("$@"); dst="${args[-1]}"; unset 'args[-1]'; cp "${args[@]}" "$(loc "$dst")" ;;
  *) exit 2 ;;
esac
'''


def test_cli_hdfs_inputs_through_the_hdfs_command(tmp_path):
    """With an hdfs:// fsURI the rows live in HDFS (drlz.sh:15 hands Spark `$INPATH/*.NN`): the CLI lists and fetches
    them with `hdfs dfs` and uploads its outputs the same way.  Without a GPU the run stops at drlz_create -- after
    the inputs have been fetched; on the GPU box it goes all the way and the outputs are compared with the oracle."""
    tmp = str(tmp_path)
    fake = tmp + "/fakehdfs"
    os.makedirs(fake + "/pg")
    d, rows = make_dataset(fake + "/pgsrc", H=3, n=3000)
    for h in range(3):
        os.rename(fake + "/pgsrc/msa/H%05d.01" % h, fake + "/pg/H%05d.01" % h)
    os.makedirs(tmp + "/bin")
    _script(tmp + "/bin/hdfs", HDFS_STUB)
    env = dict(os.environ, PATH=tmp + "/bin:" + os.environ["PATH"], FAKE_HDFS=fake, TMPDIR=tmp)
    r = subprocess.run([EXE, tmp + "/radix", fake + "/pgsrc/ref.txt", "hdfs://x/pg/*.01", "hdfs://x", "hdfs://x/pgdrlz", "hdfs://x/pggapless"],
                       capture_output=True, text=True, env=env, timeout=300)
    calls = open(fake + "/calls.log").read()
    assert "dfs -ls -C hdfs://x/pg/*.01" in calls
    assert calls.count("dfs -get -f hdfs://x/pg/H0000") == 3
    assert not [f for f in os.listdir(tmp) if f.startswith("drlz_stage_")], "staging directory left behind"
    if not _have_gpu():
        assert r.returncode == 1 and "no CUDA device" in r.stderr
        return
    assert r.returncode == 0, r.stderr
    sa = orc.sa_build(d)
    for h in range(3):
        x = orc.strip_gaps(rows[h]).tobytes()
        assert open(fake + "/pgdrlz/H%05d.01.lz" % h, "rb").read() == orc.lz_bytes(orc.encode_literal(d, sa, x)[0])
        assert open(fake + "/pggapless/part-%05d" % h, "rb").read() == b">H%05d.01\n" % h + x + b"\n"
    assert os.path.exists(fake + "/pggapless/_SUCCESS")


def test_drlz_all_spreads_chromosomes_over_gpus(tmp_path):
    """launch.sh:41 for one box: every chromosome runs exactly once, never two at a time on one GPU"""
    tmp = str(tmp_path)
    _script(tmp + "/fake_drlz.sh", r'''#!/usr/bin/env bash
lock="$LOGDIR/gpu$DRLZ_DEVICE.busy"
if ! mkdir "$lock" 2>/dev/null; then echo "overlap on gpu $DRLZ_DEVICE" >> "$LOGDIR/errors"; fi
echo "$1 $DRLZ_DEVICE $DRLZ_GPUS $2 $3" >> "$LOGDIR/runs"
sleep 0.0$(( RANDOM % 9 ))
rmdir "$lock"
[ "$1" = "$FAIL_CHR" ] && exit 1
exit 0
''')
    for fail, want_rc in (("", 0), ("7", 1)):
        logdir = tmp + "/log" + fail
        os.makedirs(logdir)
        env = dict(os.environ, DRLZ_GPUS="3", DRLZ_SH=tmp + "/fake_drlz.sh", LOGDIR=logdir, FAIL_CHR=fail, TMPDIR=tmp)
        r = subprocess.run(["bash", os.path.join(ROOT, "scripts", "drlz_all.sh"), "pg", "hdfs://nn:8020", "1", "22"], cwd=tmp, env=env,
                           capture_output=True, text=True, timeout=120)
        assert r.returncode == want_rc, r.stderr
        runs = [ln.split() for ln in open(logdir + "/runs").read().splitlines()]
        assert sorted(int(x[0]) for x in runs) == list(range(1, 23))
        assert {x[1] for x in runs} == {"0", "1", "2"} and all(x[2] == "1" and x[3] == "pg" and x[4] == "hdfs://nn:8020" for x in runs)
        assert not os.path.exists(logdir + "/errors")
        if fail:
            assert "failed chromosomes: 7" in r.stderr
    assert "DRLZ:" in open(tmp + "/runtime.log").read()


def test_offsets_map_and_rebased_phrases(tmp_path, monkeypatch):
    """phrase positions are reference coordinates, CHIC indexes the merged multi-FASTA (headers and
    newlines included), whose first record is the reference (README.md:90).  The offset map written next to the merged
    FASTA gives the shift: every rebased phrase stream must decode against the text of chrNN.fa itself."""
    tmp = str(tmp_path)
    d, rows = make_dataset(tmp, H=5)
    monkeypatch.setenv("DRLZ_MERGED_LZ", tmp + "/chr01.lz")
    monkeypatch.setenv("DRLZ_MERGED_FA", tmp + "/chr01.fa")
    assert drz.main([tmp + "/radix", tmp + "/ref.txt", tmp + "/msa/*.01", "", tmp + "/drlz", tmp + "/gapless"],
                    parser=oracle_parser, rank=0, world=1) == 0
    fa = open(tmp + "/chr01.fa", "rb").read()
    offs = drz.read_offsets(tmp + "/chr01.fa.offsets")
    assert [o[0] for o in offs] == ["H%05d.01" % h for h in range(5)]
    for name, hdr, first, nb in offs:
        assert fa[hdr:hdr + 1] == b">" and fa[hdr + 1:first - 1].decode() == name and fa[first - 1:first] == b"\n"
        assert fa[first + nb:first + nb + 1] == b"\n"
    # record 0 is the reference itself (row 0 of the MSA without its gaps)
    ref_first, ref_len = offs[0][2], offs[0][3]
    assert fa[ref_first:ref_first + ref_len] == d.tobytes()
    text = np.frombuffer(fa, dtype=np.uint8)
    for h, (name, hdr, first, nb) in enumerate(offs):
        pairs = np.frombuffer(open(tmp + "/drlz/%s.lz" % name, "rb").read(), dtype="<i8").reshape(-1, 2)
        reb = drz.rebase_pairs(pairs, ref_first)
        assert orc.decode(text, reb) == fa[first:first + nb]            # decodes against the FASTA text, not the reference
        assert orc.decode(d, pairs) == fa[first:first + nb]


def test_gap_runs_oracle_matches_scoreMatrix_definition():
    """orc.gap_runs against a literal restatement of ScoreMatrix.scala:91-93 (gap column list) and :63-79 (consecutive)"""
    rng = np.random.default_rng(3)
    for _ in range(200):
        n = int(rng.integers(0, 60))
        row = bytes(rng.choice(list(b"ACGT---"), size=n).astype(np.uint8))
        gap = [i for i in range(len(row)) if row[i:i + 1] == b"-"]

        def consecutive(index):
            ln, i = 1, index
            while i + 1 < len(gap) and gap[i + 1] == gap[i] + 1:
                ln += 1; i += 1
            return ln
        want, i = [], 0
        while i < len(gap):
            c = consecutive(i)
            want.append((gap[i], c)); i += c
        got = orc.gap_runs(row)
        assert got.tolist() == [list(w) for w in want]
