"""Regenerate the committed golden vectors from the UNMODIFIED reference (run in the
build container, where /root/reference exists):

    make -C oracle            # builds oracle/_ref/gc_refdump from /root/reference
    python tests/golden/make_golden.py

c1   = the reference's own smoke fixture test/graph.gfa + test/read.fa (BASELINE config 1)
tiny = synthetic 20 kbp bubble graph + 8 simulated 1.5 kb reads at 15 % error, one of
       them with a novel 400-bp insertion (graphchainer_b200.synth, fixed seeds)
tiny_nocc = the reference's GAM and stdout for the tiny case with --no-colinear-chaining
tiny_vg = the tiny graph written as a .vg stream (sparse node ids, two gzip members) and the reference's GAM for the same reads
Each case stores the reference's index arrays (.gcidx) and its per-stage records
(.stages) -- every K1 extension with its trace, anchors, chain, path, edlib result and
the final alignments -- plus the GAM written by the unmodified whole program.
"""
import gzip
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from graphchainer_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
REFDUMP = os.path.join(ROOT, "oracle", "_ref", "gc_refdump")
REFBIN = os.path.join(ROOT, "oracle", "_ref", "GraphChainer_ref")
TMP = "/tmp/gc_golden"


def gz(src, dst):
    with open(src, "rb") as fi, gzip.GzipFile(dst, "wb", mtime=0) as fo:
        shutil.copyfileobj(fi, fo)


def _varint(v):
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _field(number, wire, payload):
    return _varint((number << 3) | wire) + (_varint(len(payload)) + payload if wire == 2 else payload)


def gfa_to_vg(gfa, vg, first_id=10, id_step=3, messages=2):
    """The GFA's segments and links as a .vg file: `messages` vg::Graph messages in two gzip members of the stream.hpp framing
    (varint count, {varint size, message}*).  Node ids first_id, first_id + id_step, ... in segment order (vg ids are arbitrary
    positive numbers, unlike the dense ids a GFA gets), names = the segment names."""
    ids, nodes, edges = {}, [], []
    for line in open(gfa):
        f = line.split()
        if f and f[0] == "S":
            ids[f[1]] = first_id + id_step * len(ids)
            nodes.append(_field(1, 2, _field(1, 2, f[2].encode()) + _field(2, 2, f[1].encode()) + _field(3, 0, _varint(ids[f[1]]))))
    for line in open(gfa):
        f = line.split()
        if f and f[0] == "L":
            e = _field(1, 0, _varint(ids[f[1]])) + _field(2, 0, _varint(ids[f[3]]))
            if f[2] == "-":
                e += _field(3, 0, _varint(1))  # from_start
            if f[4] == "-":
                e += _field(4, 0, _varint(1))  # to_end
            edges.append(_field(2, 2, e))
    graphs = []
    for m in range(messages):
        graphs.append(b"".join(nodes[m::messages]) + b"".join(edges[m::messages]))
    with open(vg, "wb") as out:
        # one group of one message, then a group with the rest, each its own gzip member
        for group in ([graphs[0]], graphs[1:]):
            if group:
                out.write(gzip.compress(_varint(len(group)) + b"".join(_varint(len(g)) + g for g in group), mtime=0))


def run_vg_case(name, gfa, fa):
    """the same reads against the graph given as .vg: the reference's GAM carries the vg ids and names"""
    vg, gam = f"{TMP}/{name}.vg", f"{TMP}/{name}.gam"
    gfa_to_vg(gfa, vg)
    subprocess.run([REFBIN, "-t", "1", "-g", vg, "-f", fa, "-a", gam], check=True, stdout=subprocess.DEVNULL)
    shutil.copy(vg, f"{OUT}/{name}.vg")
    shutil.copy(gam, f"{OUT}/{name}.gam")


def run_case(name, gfa, fa):
    idx, st, gam = f"{TMP}/{name}.gcidx", f"{TMP}/{name}.stages", f"{TMP}/{name}.gam"
    subprocess.run([REFDUMP, "-t", "1", "-g", gfa, "-f", fa, "--gc-index", idx, "--gc-stages", st], check=True, stdout=subprocess.DEVNULL)
    subprocess.run([REFBIN, "-t", "1", "-g", gfa, "-f", fa, "-a", gam], check=True, stdout=subprocess.DEVNULL)
    gz(idx, f"{OUT}/{name}.gcidx.gz")
    gz(st, f"{OUT}/{name}.stages.gz")
    shutil.copy(gam, f"{OUT}/{name}.gam")
    shutil.copy(gfa, f"{OUT}/{name}.gfa")
    shutil.copy(fa, f"{OUT}/{name}.fa")


if __name__ == "__main__":
    os.makedirs(TMP, exist_ok=True)
    run_case("c1", "/root/reference/test/graph.gfa", "/root/reference/test/read.fa")
    g = synth.SynthGraph(20_000, seed=11)
    with open(f"{TMP}/tiny_in.gfa", "w") as f:
        f.write(g.gfa())
    synth.write_fasta(f"{TMP}/tiny_in.fa", synth.simulate_reads(g, 8, 1500, 0.15, seed=12, novel_insertion_frac=0.2))
    run_case("tiny", f"{TMP}/tiny_in.gfa", f"{TMP}/tiny_in.fa")
    run_vg_case("tiny_vg", f"{TMP}/tiny_in.gfa", f"{TMP}/tiny_in.fa")
    # edge.fa / edge.fq.gz (committed; made once from tiny.fa: lower case, reads of 10 / 34 / 35 / 36 / 64 / 65 / 100 bases, random and
    # poly-A reads, a run of N, a name used twice, a name with blanks) against the tiny graph: GAM and the summary lines
    for fixture, tag in (("edge.fa", "edge"), ("edge.fq.gz", "edge_fq")):
        with open(f"{TMP}/{tag}.log", "w") as log:
            subprocess.run([REFBIN, "-t", "1", "-g", f"{TMP}/tiny_in.gfa", "-f", f"{OUT}/{fixture}", "-a", f"{OUT}/{tag}.gam"], check=True, stdout=log)
        with open(f"{OUT}/{tag}.txt", "w") as f:
            f.write("".join(open(f"{TMP}/{tag}.log").readlines()[-8:]))
    # the reference "as in GraphAligner": GAM and the text it prints (banner, summary)
    with open(f"{OUT}/tiny_nocc.txt", "w") as log:
        subprocess.run([REFBIN, "-t", "1", "-g", f"{TMP}/tiny_in.gfa", "-f", f"{TMP}/tiny_in.fa", "-a", f"{OUT}/tiny_nocc.gam", "--no-colinear-chaining"], check=True, stdout=log)
    print("golden vectors written to", OUT)
