"""bin/drlz on BASELINE.json configs[1] as files on tmpfs, over the knobs of its host pipeline (GPU box).

    python tools/cli_sweep.py [n] [H]          env CLI_SWEEP="threads:batchMiB:slots,..." (default: a small grid)

One line per run: encode Gbp/s, the waits of the pipeline (read / GPU / write), and the DRLZ_IO_PROBE figure (same
pipeline, device left out) for the same knobs.
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 48_000_000
    H = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    cpus = len(os.sched_getaffinity(0))
    grid = os.environ.get("CLI_SWEEP", "%d:128:4,%d:128:4,%d:64:6,%d:256:3,%d:128:4:probe" % (cpus, max(4, cpus - 4), cpus, cpus, cpus))
    seed = orc.seed_for(2)
    d = orc.syn_reference(seed, n)
    tmp = tempfile.mkdtemp(prefix="drlz_sweep_", dir="/dev/shm")
    try:
        d.tofile(os.path.join(tmp, "ref.txt"))
        os.makedirs(os.path.join(tmp, "msa"))
        bases = 0
        for a in range(0, H, 10):
            rows = orc.syn_rows(seed, d, a, min(10, H - a), 1e-3, 1e-5, 1e-5)
            for r in range(rows.shape[0]):
                rows[r].tofile(os.path.join(tmp, "msa", "H%05d.21" % (a + r)))
        exe = os.path.join(ROOT, "pangenspark_b200", "bin", "drlz")
        print(json.dumps({"cpus": cpus, "n": n, "H": H}))
        for item in grid.split(","):
            f = item.split(":")
            env = dict(os.environ, DRLZ_IO_THREADS=f[0], DRLZ_BATCH_BYTES=str(int(f[1]) << 20), DRLZ_SLOTS=f[2],
                       DRLZ_STATS=os.path.join(tmp, "stats.json"))
            if len(f) > 3 and f[3] == "probe":
                env["DRLZ_IO_PROBE"] = "1"
            if len(f) > 3 and f[3] == "nogapless":
                env["DRLZ_GAPLESS"] = "0"
            if len(f) > 4:
                env["DRLZ_OPTIONS"] = f[4].replace(";", ",")
            for rep in range(int(os.environ.get("CLI_SWEEP_REPS", 2))):
                shutil.rmtree(os.path.join(tmp, "out"), ignore_errors=True)
                shutil.rmtree(os.path.join(tmp, "gapless"), ignore_errors=True)
                p = subprocess.run([exe, os.path.join(tmp, "radix"), os.path.join(tmp, "ref.txt"), os.path.join(tmp, "msa", "*.21"), "file:///",
                                    os.path.join(tmp, "out"), os.path.join(tmp, "gapless")], env=env, capture_output=True, text=True)
                if p.returncode != 0:
                    print(item, "FAILED", p.stderr[-500:])
                    continue
                st = json.load(open(os.path.join(tmp, "stats.json")))
                print(json.dumps({"knobs": item, "gbp_per_s": round(st["bases"] / st["encode_s"] / 1e9, 3), "encode_s": st["encode_s"],
                                  "read_wait_s": st["read_wait_s"], "gpu_wait_s": st["gpu_wait_s"], "write_wait_s": st["write_wait_s"],
                                  "index_s": st["index_s"], "batches": st["batches"]}))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
