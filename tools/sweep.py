"""Parameter sweep of the device-resident parse on the bench workload (development tool, GPU box only).
usage: python tools/sweep.py "chunk=2048,kmer=12,machine=0" "chunk=4096,kmer=13,machine=0" ..."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from pangenspark_b200 import drlz as D  # noqa: E402
import bench  # noqa: E402


def main():
    wl = bench.workload(0)
    for key in ("p_snp", "p_ins", "p_del"):          # e.g. SWEEP_P_SNP=8e-3 SWEEP_P_INS=1e-3 SWEEP_P_DEL=1e-3: config-5-like divergence
        if os.environ.get("SWEEP_" + key.upper()):
            wl[key] = float(os.environ["SWEEP_" + key.upper()])
    d = orc.syn_reference(orc.seed_for(wl["cfg"]), wl["n"])
    rows_arr, M, stride = bench.gen_rows(orc, wl, d)
    H = wl["H"]
    dev_rows = torch.from_numpy(rows_arr.reshape(-1)).cuda()
    row_off = np.arange(H, dtype=np.uint64) * np.uint64(stride)
    cols = np.full(H, M, dtype=np.uint64)
    ctx = D.Drlz(0)
    ref_pairs = None
    last_k = None
    for spec in sys.argv[1:]:
        opts = dict(kv.split("=") for kv in spec.split(","))
        k = int(opts.get("kmer", 0))
        if k != last_k:
            ctx.set_option("kmer", k)
            ctx.build_index(d)
            last_k = k
        for key, v in opts.items():
            if key != "kmer":
                ctx.set_option(key, int(v))
        for _ in range(3):
            ptr, npairs, m = ctx.parse_batch_device(dev_rows.data_ptr(), row_off, cols)
        acc = {}
        R = 5
        for _ in range(R):
            ptr, npairs, m = ctx.parse_batch_device(dev_rows.data_ptr(), row_off, cols)
            for a, b in ctx.parse_timings().items():
                acc[a] = acc.get(a, 0.0) + b / R
        pairs = ctx.copy_pairs_to_host(ptr, int(npairs.sum()))
        if ref_pairs is None:
            ref_pairs = pairs
        same = pairs.shape == ref_pairs.shape and bool((pairs == ref_pairs).all())
        print(spec, "same=%s" % same, "Gbp/s=%.1f" % (m.sum() / acc["device_ms"] / 1e6),
              {a: round(b, 3) for a, b in acc.items()}, flush=True)


if __name__ == "__main__":
    main()
