"""Where a kernel's instructions and stall samples go, by source line and by inlined call site (development tool,
runs in the build container, no GPU needed).

ncu's source page of an imported report lists metrics per SASS instruction; `nvdisasm -gi` lists, for the same
instructions in the same order, the source line and the chain of inlined call sites.  Joining the two by instruction
offset attributes every warp instruction / stall sample of a kernel that is one big inlined function (k_spec is
4300 instructions of rlz_core.cuh) to the line that caused it AND to the place it was called from -- which is what
items 14, 24 and 25 of profiles/README.md are based on.

usage:  python tools/ncu_lines.py <report.ncu-rep> <kernel regex> <cubin name inside libdrlz.so: parse|pack|sa_build|api>
                                  [--metric inst|long_sb|samples] [--frames N] [--top N]

The library must be the build the report was taken from: the tool picks the instantiation of the kernel whose
instruction offsets and opcodes agree with the report's and stops if there is none.
"""
import argparse
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pangenspark_b200", "lib", "libdrlz.so")


def sass_with_inline_chains(cubin_name: str, tmp: str):
    subprocess.check_call(["cuobjdump", "-xelf", cubin_name, LIB], cwd=tmp, stdout=subprocess.DEVNULL)
    cubins = [f for f in os.listdir(tmp) if f.startswith(cubin_name) and f.endswith(".cubin")]
    if not cubins:
        sys.exit("no cubin named %s* in %s" % (cubin_name, LIB))
    return subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubins[0])], check=True, capture_output=True,
                          text=True).stdout.split("\n")


def kernel_instructions(lines, kernel_re):
    """{section name: [(offset, ((file, line), ...innermost first), text)]} of every .text section whose name matches
    (a templated kernel has one section per instantiation)"""
    out = {}
    starts = [i for i, l in enumerate(lines) if l.startswith(".text.") and l.rstrip().endswith(":") and re.search(kernel_re, l)]
    if not starts:
        sys.exit("kernel %s not found in the disassembly" % kernel_re)
    for start in starts:
        ins, chain, pending = [], (), []
        for l in lines[start + 1:]:
            if l.startswith("//---------------------"):
                break
            m = re.search(r'//## File "([^"]+)", line (\d+)', l)
            if m:
                pending.append((os.path.basename(m.group(1)), int(m.group(2))))
                continue
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
            if m:
                if pending:
                    chain, pending = tuple(pending), []
                ins.append((int(m.group(1), 16), chain, m.group(2)))
        out[lines[start].strip()] = ins
    return out


def opcode(text):
    t = text.split()
    if t and t[0].startswith("@"):
        t = t[1:]
    return t[0] if t else ""


def source_page(report, kernel_re):
    out = subprocess.run(["ncu", "-i", report, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel_re,
                          "--print-source", "sass"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr, data = rows[hdr_i], []
    for r in rows[hdr_i + 1:]:
        if not r or not r[0].startswith("0x"):
            break                                        # the next launch of the same kernel starts a new block
        data.append(r)
    return hdr, data


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report"); ap.add_argument("kernel"); ap.add_argument("cubin")
    ap.add_argument("--metric", default="inst", choices=["inst", "long_sb", "samples"])
    ap.add_argument("--frames", type=int, default=3, help="outermost call-site frames to keep in the key")
    ap.add_argument("--top", type=int, default=30)
    a = ap.parse_args()
    with tempfile.TemporaryDirectory() as tmp:
        sections = kernel_instructions(sass_with_inline_chains(a.cubin, tmp), a.kernel)
    hdr, data = source_page(a.report, a.kernel)
    col = {"inst": "Instructions Executed", "long_sb": "stall_long_sb", "samples": "# Samples"}[a.metric]
    ci, ti, ai, si = hdr.index(col), hdr.index("Thread Instructions Executed"), hdr.index("Address"), hdr.index("Source")
    ei = hdr.index("Instructions Executed")
    base = int(data[0][ai], 16)
    ins = None
    for name, cand in sections.items():          # the instantiation the report holds: same length, same opcodes
        if abs(len(cand) - len(data)) <= 64 and all(
                int(data[k][ai], 16) - base == cand[k][0] and opcode(data[k][si]) == opcode(cand[k][2])
                for k in range(min(len(cand), len(data)))):
            ins = cand
            print("section", name)
            break
    if ins is None:
        sys.exit("no instantiation of %s in %s has the SASS of the report: not the same build" % (a.kernel, LIB))
    n = min(len(data), len(ins))
    by_line, by_site = collections.Counter(), collections.Counter()
    lanes_n, lanes_d = collections.Counter(), collections.Counter()
    total = 0
    for k in range(n):
        v = int(data[k][ci] or 0)
        chain = ins[k][1]
        leaf = chain[0] if chain else ("?", 0)
        site = tuple(chain[1:][-a.frames:]) if len(chain) > 1 else ()
        by_line[leaf] += v
        by_site[(leaf, site)] += v
        lanes_n[(leaf, site)] += int(data[k][ti] or 0); lanes_d[(leaf, site)] += int(data[k][ei] or 0)
        total += v
    print("%s of %s: %d in %d instructions" % (col, a.kernel, total, n))
    print("\n-- by source line")
    for key, v in by_line.most_common(a.top):
        print("%6.2f %%  %s:%d" % (100.0 * v / max(total, 1), key[0], key[1]))
    print("\n-- by source line and call site (outermost %d frames), with active lanes per instruction" % a.frames)
    for key, v in by_site.most_common(a.top):
        leaf, site = key
        print("%6.2f %%  lanes %4.1f  %s:%d  <-  %s" % (100.0 * v / max(total, 1), lanes_n[key] / max(lanes_d[key], 1), leaf[0], leaf[1],
                                                      " <- ".join("%s:%d" % f for f in site) or "-"))


if __name__ == "__main__":
    main()
