"""S0 parity.  (1) The device form of the reference's k-mer walk + minimizer index probes (gcgpu_seed, gc_seed.cuh) and the
host seed ordering (gc_seeder.h) against the seed vectors of the UNMODIFIED reference (gc_refdump's SEEDS_ORDERED /
SEEDS_BYPOS records: golden files, and a live run on the GPU box), every field in both orders.  (2) The k-mer emission
rule against the sequential restatement of MinimizerSeeder::iterateKmers (tests/hostsim/seed_ref.h), match by match, on
reads that exercise the restart rule (N, U, lower case), homopolymer re-emission and reads shorter than k.  `not gpu`: through the C-ABI test
double; `gpu`: through libgcgpu.so on the device."""
import os
import random
import subprocess

import pytest

from conftest import ROOT

HOSTSIM = os.path.join(ROOT, "tests", "hostsim")


def _reads(path):
    rng = random.Random(7)
    src = open(os.path.join(ROOT, "tests", "golden", "tiny.fa")).read().split("\n")
    base = [l for l in src if l and not l.startswith(">")]
    reads = list(base)
    for s in base[:8]:
        t = list(s)
        for _ in range(12):
            p = rng.randrange(len(t))
            t[p] = rng.choice("NnUuRY")
        reads.append("".join(t))
    for s in base[:6]:
        p = rng.randrange(len(s) - 200)
        reads.append(s[:p] + rng.choice("ACGT") * rng.choice([15, 16, 17, 21, 22, 40, 300]) + s[p:])
    reads += ["A" * 500, "ACGT" * 50, "acgtacgtacgtacgtacgt", "ACGTACGTACGTAC", "", "T" * 15, "T" * 16, "G" * 21 + "N" + "G" * 30 + "U" + "G" * 14]
    with open(path, "w") as f:
        for i, s in enumerate(reads):
            f.write(f">r{i}\n{s}\n")
    return len(reads)


def _build(tmp_path, link):
    exe = str(tmp_path / "seed_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fopenmp", "-Wno-sign-compare", "-o", exe, os.path.join(HOSTSIM, "seed_check.cpp"), *link, "-lz"], check=True)
    return exe


def _run(tmp_path, golden_files, link):
    idx, stages = golden_files["tiny"]
    fa = str(tmp_path / "seed_reads.fa")
    n = _reads(fa)
    exe = _build(tmp_path, link)
    # third argument: the reference's own seed records (SEEDS_ORDERED / SEEDS_BYPOS of gc_refdump) for the golden reads
    out = subprocess.run([exe, idx, fa, stages], check=True, capture_output=True, text=True).stdout.split()
    res = dict(zip(out[0::2], map(int, out[1::2])))
    assert res["reads"] == n and res["mismatches"] == 0
    assert res["matches"] > 100 and res["seeds"] > 100 and res["emitted"] > 20000
    assert res["ref_reads"] >= 8 and res["ref_seeds"] >= 500   # every seed of every golden read: position, node, offset, goodness, cluster size, both orders


_GPU_LINK = ["-L" + os.path.join(ROOT, "graphchainer_b200"), "-lgcgpu", "-Wl,-rpath," + os.path.join(ROOT, "graphchainer_b200")]


def test_seed_lookups_match_reference_restatement_sim(tmp_path, golden_files):
    _run(tmp_path, golden_files, [os.path.join(HOSTSIM, "gcgpu_sim.cpp")])


@pytest.mark.gpu
def test_seed_lookups_match_reference_restatement_gpu(tmp_path, golden_files):
    assert os.path.exists(os.path.join(ROOT, "graphchainer_b200", "libgcgpu.so")), "libgcgpu.so not built (run __graft_entry__.build())"
    _run(tmp_path, golden_files, _GPU_LINK)


@pytest.mark.gpu
def test_seeds_match_live_reference_records_gpu(tmp_path):
    """S0 on the device + the host seed ordering against the seed vectors the unmodified reference builds for 150 fresh
    reads (gc_refdump run live): SEEDS_ORDERED and SEEDS_BYPOS, every field, both tie orders."""
    from conftest import REFDUMP
    if not os.path.exists(REFDUMP):
        pytest.skip("oracle/_ref/gc_refdump not built")
    from graphchainer_b200 import synth
    g = synth.SynthGraph(600_000, seed=81)
    gfa, fa = str(tmp_path / "g.gfa"), str(tmp_path / "r.fa")
    with open(gfa, "w") as f:
        f.write(g.gfa())
    synth.write_fasta(fa, synth.simulate_reads(g, 150, 7000, 0.15, seed=82))
    idx, st = str(tmp_path / "x.gcidx"), str(tmp_path / "x.stages")
    subprocess.run([REFDUMP, "-t", "1", "-g", gfa, "-f", fa, "--gc-index", idx, "--gc-stages", st], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    exe = _build(tmp_path, _GPU_LINK)
    out = subprocess.run([exe, idx, fa, st], check=True, capture_output=True, text=True).stdout.split()
    res = dict(zip(out[0::2], map(int, out[1::2])))
    assert res["mismatches"] == 0 and res["ref_reads"] == 150 and res["ref_seeds"] > 30000
