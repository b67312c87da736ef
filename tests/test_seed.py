"""S0 parity: the device form of the reference's k-mer walk + minimizer index probes (gcgpu_seed,
gc_seed.cuh) against the sequential restatement of MinimizerSeeder::getSeeds/iterateKmers
(tests/hostsim/seed_ref.h), match by match, on reads that exercise the restart rule (N, U, lower
case), homopolymer re-emission and reads shorter than k.  `not gpu`: through the C-ABI test
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


def _run(tmp_path, golden_files, link):
    idx, _ = golden_files["tiny"]
    fa = str(tmp_path / "seed_reads.fa")
    n = _reads(fa)
    exe = str(tmp_path / "seed_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fopenmp", "-Wno-sign-compare", "-o", exe, os.path.join(HOSTSIM, "seed_check.cpp"), *link, "-lz"], check=True)
    out = subprocess.run([exe, idx, fa], check=True, capture_output=True, text=True).stdout.split()
    res = dict(zip(out[0::2], map(int, out[1::2])))
    assert res["reads"] == n and res["mismatches"] == 0
    assert res["matches"] > 100 and res["seeds"] > 100 and res["emitted"] > 20000


def test_seed_lookups_match_reference_restatement_sim(tmp_path, golden_files):
    _run(tmp_path, golden_files, [os.path.join(HOSTSIM, "gcgpu_sim.cpp")])


@pytest.mark.gpu
def test_seed_lookups_match_reference_restatement_gpu(tmp_path, golden_files):
    libdir = os.path.join(ROOT, "graphchainer_b200")
    assert os.path.exists(os.path.join(libdir, "libgcgpu.so")), "libgcgpu.so not built (run __graft_entry__.build())"
    _run(tmp_path, golden_files, ["-L" + libdir, "-lgcgpu", "-Wl,-rpath," + libdir])
