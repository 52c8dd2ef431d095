"""Python mirror of the reference's stage entry point `org.ngseq.pangenspark.DistributedRLZ.main`
(DistributedRLZ.scala:32-293), same six positional arguments and same outputs:

    python -m pangenspark_b200.distributed_rlz <localradix> <localref> <dataGlob> <fsURI> <outDir> <gaplessOut>

It exists for tests and for multi-GPU runs under torchrun (one process per GPU; haplotypes are sharded by
sorted basename, every rank writes its own `.lz` and `part-NNNNN` files exactly like a Spark executor does,
DistributedRLZ.scala:247-290).  The native equivalent is `pangenspark_b200/bin/drlz` (cli/drlz_main.cpp).
The parse itself always goes through libdrlz.so; there is no CPU path here.
"""
from __future__ import annotations

import glob
import os
import sys

import numpy as np


def strip_scheme(p: str) -> str:
    return p[7:] if p.startswith("file://") else p


def list_inputs(data_glob: str):
    """Input files sorted by basename: the order `hdfs dfs -getmerge <outDir>/*.NN.lz` later concatenates them
    (components/index/index_chr.sh:16)."""
    files = glob.glob(strip_scheme(data_glob))
    return sorted(files, key=lambda f: os.path.basename(f))


def shard(items, rank: int, world: int):
    """Contiguous slices of the sorted list; sizes differ by at most one (rows are equal length in an MSA)."""
    n = len(items)
    lo = (n * rank) // world
    hi = (n * (rank + 1)) // world
    return list(range(lo, hi))


def read_reference(path: str) -> bytes:
    """DistributedRLZ.scala:61: all lines concatenated."""
    raw = read_bytes(path)
    return raw.replace(b"\n", b"").replace(b"\r", b"")


def read_bytes(path: str) -> bytes:
    """gzip files (the legacy flow globs *.NN.gz, create.sh:27; Spark's text reader inflates them transparently)
    are inflated, anything else is read as it is."""
    with open(path, "rb") as f:
        raw = f.read()
    if raw[:2] == b"\x1f\x8b":
        import gzip
        raw = gzip.decompress(raw)
    return raw


def read_row(path: str) -> bytes:
    raw = read_bytes(path)
    body = raw.rstrip(b"\r\n")
    if b"\n" in body:
        raise ValueError(f"{path}: more than one line; every line would overwrite the same .lz in the reference "
                         "(DistributedRLZ.scala:255-257), so multi-line rows are unsupported")
    return body


def write_atomic(path: str, chunks):
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        for c in chunks:
            f.write(c)
    os.replace(tmp, path)


def write_outputs(out_dir: str, gapless_dir: str, name: str, index: int, pairs: np.ndarray, gapless: bytes):
    """<outDir>/<basename>.lz = int64 LE (pos, len) records (DistributedRLZ.scala:257,263-284);
    <gaplessOut>/part-NNNNN = ">basename\\n<gapless>\\n" (DistributedRLZ.scala:75-80, one record per text line)."""
    write_atomic(os.path.join(out_dir, name + ".lz"), [np.ascontiguousarray(pairs, dtype="<i8").tobytes()])
    write_atomic(os.path.join(gapless_dir, "part-%05d" % index), [b">" + name.encode() + b"\n", gapless, b"\n"])


def main(argv, parser=None, rank: int | None = None, world: int | None = None, batch_bytes: int = 1 << 30) -> int:
    """parser(ref_bytes, rows) -> (list of pairs arrays, list of gapless bytes).  Default: the CUDA library.
    Tests inject a stand-in to exercise the host logic (naming, ordering, sharding) without a GPU."""
    if len(argv) < 6:
        print("usage: distributed_rlz <localradix> <localref> <dataGlob> <fsURI> <outDir> <gaplessOut>", file=sys.stderr)
        return 2
    localradix, localref, data_glob, fs_uri, out_dir, gapless_dir = argv[:6]
    if fs_uri.startswith("hdfs://"):
        print("hdfs:// output needs the `hdfs` CLI; use the native drlz binary or a file:// URI", file=sys.stderr)
        return 2
    rank = int(os.environ.get("RANK", 0)) if rank is None else rank
    world = int(os.environ.get("WORLD_SIZE", 1)) if world is None else world
    out_dir, gapless_dir = strip_scheme(out_dir), strip_scheme(gapless_dir)
    os.makedirs(out_dir, exist_ok=True)
    os.makedirs(gapless_dir, exist_ok=True)
    ref = read_reference(localref)
    files = list_inputs(data_glob)
    if not files:
        print("no input matches %s (Path does not exist)" % data_glob, file=sys.stderr)      # Spark throws here
        return 1
    mine = shard(files, rank, world)
    ctx = None
    if parser is None:
        from . import drlz as D
        ctx = D.Drlz(int(os.environ.get("LOCAL_RANK", 0)))      # raises without a GPU: no CPU fallback
        print("Create suffix")
        ctx.build_index(ref)
        print("Load suffix\nbroadcasting")
        if os.environ.get("DRLZ_EMIT_RADIX") == "1" and rank == 0:
            np.savetxt(localradix, ctx.index_get("sa"), fmt="%d")

        def parser(_ref, rows):
            res, m, gl = ctx.parse_batch(rows, strip_gaps=True, want_gapless=True)
            return res, [g.tobytes() for g in gl]
    print("started encoding")
    i = 0
    while i < len(mine):
        batch, acc = [], 0
        while i < len(mine) and (not batch or acc < batch_bytes):
            acc += os.path.getsize(files[mine[i]])
            batch.append(mine[i]); i += 1
        rows = [read_row(files[j]) for j in batch]
        pairs, gapless = parser(ref, rows)
        for j, p, g in zip(batch, pairs, gapless):
            write_outputs(out_dir, gapless_dir, os.path.basename(files[j]), j, p, g)
    if ctx is not None:
        ctx.close()
    if rank == 0:
        write_atomic(os.path.join(gapless_dir, "_SUCCESS"), [])
    if world == 1:
        merge_outputs(out_dir, gapless_dir, [os.path.basename(f) for f in files])
    return 0


def merge_outputs(out_dir: str, gapless_dir: str, names):
    """what `hdfs dfs -getmerge <outDir>/*.NN.lz chrNN.lz` and `-getmerge <gapless>/*NN* chrNN.fa`
    do in components/index/index_chr.sh:16,20 -- concatenation in sorted-basename order -- written directly when
    DRLZ_MERGED_LZ / DRLZ_MERGED_FA name the target files (call after every rank has finished)."""
    lz, fa = os.environ.get("DRLZ_MERGED_LZ"), os.environ.get("DRLZ_MERGED_FA")
    if lz:
        with open(lz + ".tmp", "wb") as out:
            for n in sorted(names):
                with open(os.path.join(out_dir, n + ".lz"), "rb") as f:
                    out.write(f.read())
        os.replace(lz + ".tmp", lz)
    if fa:
        parts = sorted(p for p in os.listdir(gapless_dir) if p.startswith("part-"))
        at = 0
        with open(fa + ".tmp", "wb") as out, open(fa + ".offsets", "w") as omap:
            for p in parts:
                with open(os.path.join(gapless_dir, p), "rb") as f:
                    rec = f.read()
                out.write(rec)
                name = rec[1:rec.index(b"\n")].decode()
                hl = len(name) + 2
                omap.write("%s\t%d\t%d\t%d\n" % (name, at, at + hl, max(0, len(rec) - hl - 1)))
                at += len(rec)
        os.replace(fa + ".tmp", fa)


def read_offsets(path: str):
    """<chrNN.fa>.offsets: name, offset of '>', offset of the first base, number of bases -- one
    line per haplotype in file order."""
    out = []
    with open(path) as f:
        for ln in f:
            name, a, b, c = ln.rstrip("\n").split("\t")
            out.append((name, int(a), int(b), int(c)))
    return out


def rebase_pairs(pairs: np.ndarray, first_base_of_reference: int) -> np.ndarray:
    """Phrase positions are coordinates of the reference sequence; CHIC indexes the merged multi-FASTA, whose first
    record is the reference (README.md:90).  A match (pos, len) becomes (first_base_of_reference + pos, len); literals
    (len == 0, pos = byte value) are left alone."""
    p = np.array(pairs, dtype=np.int64).reshape(-1, 2).copy()
    p[p[:, 1] > 0, 0] += first_base_of_reference
    return p


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
