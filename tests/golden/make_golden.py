"""Regenerate the committed golden vectors from the UNMODIFIED reference (run in the
build container, where /root/reference exists):

    make -C oracle            # builds oracle/_ref/gc_refdump from /root/reference
    python tests/golden/make_golden.py

c1   = the reference's own smoke fixture test/graph.gfa + test/read.fa (BASELINE config 1)
tiny = synthetic 20 kbp bubble graph + 8 simulated 1.5 kb reads at 15 % error, one of
       them with a novel 400-bp insertion (graphchainer_b200.synth, fixed seeds)
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
    print("golden vectors written to", OUT)
