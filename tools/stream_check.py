"""BASELINE.json configs[3] through the drop-in itself: a chr1-sized reference and many haplotype files streamed from the
host through bin/drlz (pinned slot ring, asynchronous copies, DRLZ_GPUS devices of one box).

    python tools/stream_check.py [n] [distinct rows] [copies per row] [gpus]

`distinct` rows are generated (oracle/synth.c, configs[3] rates) into /dev/shm and every one of them is given `copies`
names (hard links: the haplotype files of a 1000-sample panel take 250 GB; the links make the CLI read, copy and parse
that many bytes while the box only holds the distinct ones).  The gapless FASTA is switched off (it would write the 250 GB
back).  Checked: every .lz of the first copy against the oracle for `check` rows, every other copy byte-identical to the
first.  Prints one JSON line: Gbp/s of the encode phase and of the whole process, host->device GB/s.
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 249_000_000
    distinct = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    copies = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    gpus = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    check = int(os.environ.get("STREAM_CHECK_ROWS", 2))
    _, _, p_snp, p_ins, p_del = orc.CONFIGS["cfg4"]
    seed = orc.seed_for(4)
    tmp = tempfile.mkdtemp(prefix="drlz_stream_", dir="/dev/shm")
    try:
        t0 = time.time()
        d = orc.syn_reference(seed, n)
        d.tofile(os.path.join(tmp, "ref.txt"))
        os.makedirs(os.path.join(tmp, "msa"))
        M = orc.syn_columns(seed, n, p_ins)
        keep = {}
        for a in range(0, distinct, 8):                    # eight rows at a time (memory)
            rows = orc.syn_rows(seed, d, a, min(8, distinct - a), p_snp, p_ins, p_del)
            for r in range(rows.shape[0]):
                h = a + r
                first = os.path.join(tmp, "msa", "S%03dH%05d.01" % (0, h))
                rows[r].tofile(first)
                for c in range(1, copies):
                    os.link(first, os.path.join(tmp, "msa", "S%03dH%05d.01" % (c, h)))
                if h < check:
                    keep[h] = rows[r].copy()
        gen_s = time.time() - t0
        env = dict(os.environ, DRLZ_GPUS=str(gpus), DRLZ_GAPLESS="0", DRLZ_STATS=os.path.join(tmp, "stats.json"))
        env.setdefault("DRLZ_BATCH_BYTES", str(256 << 20))       # one chr1-sized row per batch: small pinned slots, short setup
        exe = os.path.join(ROOT, "pangenspark_b200", "bin", "drlz")
        t0 = time.time()
        p = subprocess.run([exe, os.path.join(tmp, "radix"), os.path.join(tmp, "ref.txt"), os.path.join(tmp, "msa", "*.01"), "file:///",
                            os.path.join(tmp, "out"), os.path.join(tmp, "gapless")], env=env, capture_output=True, text=True)
        wall = time.time() - t0
        if p.returncode != 0:
            print(p.stdout[-2000:], p.stderr[-2000:], file=sys.stderr)
            return 1
        st = json.load(open(os.path.join(tmp, "stats.json")))
        ok = True
        sa = orc.sa_build(d) if check else None
        for h in range(min(check, distinct)):
            want = orc.lz_bytes(orc.encode_literal(d, sa, orc.strip_gaps(keep[h]))[0])
            got = open(os.path.join(tmp, "out", "S%03dH%05d.01.lz" % (0, h)), "rb").read()
            ok = ok and got == want
        sums = {}
        for f in sorted(os.listdir(os.path.join(tmp, "out"))):
            hid = f[4:]
            dig = hashlib.sha256(open(os.path.join(tmp, "out", f), "rb").read()).hexdigest()
            ok = ok and sums.setdefault(hid, dig) == dig
        nfiles = len(os.listdir(os.path.join(tmp, "out")))
        ok = ok and nfiles == distinct * copies
        out = {"workload": "configs[3]: %.0f Mbp reference, %d haplotype files (%d distinct x %d names), %d GPU(s), streamed through bin/drlz" % (
                   n / 1e6, distinct * copies, distinct, copies, gpus),
               "bases": st["bases"], "input_bytes": st["input_bytes"], "encode_s": st["encode_s"], "setup_s": st["index_s"], "wall_s": wall,
               "encode_gbp_per_s": st["bases"] / st["encode_s"] / 1e9, "job_gbp_per_s": st["bases"] / wall / 1e9,
               "h2d_gb_per_s": st["input_bytes"] / st["encode_s"] / 1e9, "batches": st["batches"], "io_threads": st["io_threads"],
               "read_wait_s": st["read_wait_s"], "gpu_wait_s": st["gpu_wait_s"], "write_wait_s": st["write_wait_s"],
               "lz_files": nfiles, "oracle_checked_rows": min(check, distinct), "copies_identical": bool(ok), "ok": bool(ok),
               "generate_s": gen_s, "columns": M}
        print(json.dumps(out))
        return 0 if ok else 1
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    sys.exit(main())
