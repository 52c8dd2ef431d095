"""Regenerates tests/golden/*.npz from the CPU oracle (oracle/drlz_oracle.c).

The reference (NGSeq/PanGenSpark) has no golden vectors and cannot run in this image, so these
fixtures are the oracle's own outputs: they pin the oracle (and the CUDA path) against drift, they do
not pin the oracle against the reference ("parity unpinned").

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def emit(name, d, rows):
    d = np.frombuffer(bytes(d), dtype=np.uint8)
    sa = orc.sa_build(d)
    out = {"ref": d, "sa": sa.astype(np.int64), "nrows": np.int64(len(rows))}
    for h, r in enumerate(rows):
        r = np.frombuffer(bytes(r), dtype=np.uint8)
        x = orc.strip_gaps(r)
        a1, _ = orc.encode_literal(d, sa, x)
        out[f"row{h}"] = r
        out[f"lz{h}"] = np.frombuffer(orc.lz_bytes(a1), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, d.size, len(rows))


def main():
    # 1. tiny hand-written MSA with gaps, literals, ties
    emit("tiny", b"ACGTACGAACGTTTGACCA",
         [b"ACGTACGAACG--TTTGACCA", b"ACGTAC-AACGNNTTTGACCA", b"--------------------", b"TTTTTTTTTT", b"ACG"])
    # 2. synthetic, cfg-1-like at 6 kbp, higher rates so that every edit type appears
    seed = orc.seed_for(1)
    d = orc.syn_reference(seed, 6000)
    rows = orc.syn_rows(seed, d, 0, 5, 1e-2, 2e-3, 2e-3)
    emit("synth6k", d.tobytes(), [r.tobytes() for r in rows])
    # 3. repetitive reference (tandem repeats + homopolymer) with a divergent haplotype
    rng = np.random.default_rng(5)
    unit = bytes(rng.choice(list(b"ACGT"), size=23).astype(np.uint8))
    d3 = unit * 40 + b"A" * 200 + unit[:11] * 30 + b"ACGT" * 50
    x3 = bytearray(d3)
    for p in rng.integers(0, len(d3), size=25):
        x3[p] = ord("ACGT"[int(rng.integers(0, 4))])
    emit("repeats", d3, [bytes(x3), bytes(x3[100:900]), b"A" * 300 + b"C" + b"A" * 150])


if __name__ == "__main__":
    main()
