"""End-to-end parity of the per-read pipeline: decoded GAM messages (every field, mapping and
edit) must equal the output of the UNMODIFIED reference program for the same inputs.
Golden GAMs: tests/golden/*.gam written by oracle/_ref/GraphChainer_ref (make_golden.py)."""
import os
import subprocess

import pytest

from conftest import GOLDEN, REFBIN, REFDUMP, ROOT
from graphchainer_b200 import gam

DRIVER = os.path.join(ROOT, "graphchainer_b200", "GraphChainerB200")


@pytest.fixture(scope="session")
def driver_sim(tmp_path_factory):
    """Host driver linked against the C-ABI test double (tests/hostsim/gcgpu_sim.cpp): checks the
    HOST logic on the GPU-less box.  The shipped driver links libgcgpu.so and has no such path."""
    out = str(tmp_path_factory.mktemp("drv") / "driver_sim")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fopenmp", "-Wno-sign-compare", "-o", out, os.path.join(ROOT, "graphchainer_b200", "csrc", "gc_driver.cpp"),
                    os.path.join(ROOT, "tests", "hostsim", "gcgpu_sim.cpp"), "-lz"], check=True)
    return out


@pytest.mark.parametrize("name", ["c1", "tiny"])
def test_host_pipeline_matches_reference_gam(driver_sim, golden_files, tmp_path, name):
    idx, _ = golden_files[name]
    out = str(tmp_path / "out.gam")
    subprocess.run([driver_sim, "--gc-index", idx, "-f", os.path.join(GOLDEN, name + ".fa"), "-a", out, "-t", "4", "--gc-quiet"], check=True, stdout=subprocess.DEVNULL)
    diffs = gam.diff_gam(gam.read_gam(out), gam.read_gam(os.path.join(GOLDEN, name + ".gam")))
    assert not diffs, diffs


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c1", "tiny"])
def test_gpu_pipeline_matches_reference_gam(golden_files, tmp_path, name):
    assert os.path.exists(DRIVER), "GraphChainerB200 not built (run __graft_entry__.build())"
    idx, _ = golden_files[name]
    out = str(tmp_path / "out.gam")
    subprocess.run([DRIVER, "--gc-index", idx, "-f", os.path.join(GOLDEN, name + ".fa"), "-a", out, "-t", "4"], check=True, stdout=subprocess.DEVNULL)
    diffs = gam.diff_gam(gam.read_gam(out), gam.read_gam(os.path.join(GOLDEN, name + ".gam")))
    assert not diffs, diffs


@pytest.mark.gpu
def test_gpu_pipeline_matches_reference_on_fresh_synthetic(tmp_path):
    """300 reads x 8 kb at 15 % error (5 % with a novel insertion) on a 1 Mbp bubble graph: the
    unmodified reference runs live on the box's CPU, the GPU pipeline must give the same GAM."""
    if not (os.path.exists(REFBIN) and os.path.exists(REFDUMP)):
        pytest.skip("oracle/_ref not built")
    from graphchainer_b200 import synth
    g = synth.SynthGraph(1_000_000, seed=51)
    gfa, fa = str(tmp_path / "g.gfa"), str(tmp_path / "r.fa")
    with open(gfa, "w") as f:
        f.write(g.gfa())
    synth.write_fasta(fa, synth.simulate_reads(g, 300, 8000, 0.15, seed=52))
    idx, ref_gam, out = str(tmp_path / "x.gcidx"), str(tmp_path / "ref.gam"), str(tmp_path / "out.gam")
    subprocess.run([REFDUMP, "-t", "1", "-g", gfa, "--gc-index", idx], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    subprocess.run([REFBIN, "-t", str(os.cpu_count() or 8), "-g", gfa, "-f", fa, "-a", ref_gam], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    subprocess.run([DRIVER, "--gc-index", idx, "-f", fa, "-a", out, "-t", str(min(32, os.cpu_count() or 8))], check=True, stdout=subprocess.DEVNULL)
    a, b = gam.read_gam(out), gam.read_gam(ref_gam)
    assert len(b) == 300
    diffs = gam.diff_gam(a, b)
    assert not diffs, diffs


@pytest.mark.gpu
def test_gpu_pipeline_matches_reference_on_ultralong_reads(tmp_path):
    """BASELINE config-4 shape: 50-100 kb reads at 12 % error (K1 items of ~1500 slices, NW bands of
    ~8-16 k diagonals, i.e. the widest K3 lane groups, many gap fills); the unmodified reference runs
    live on the box's CPU, every field of every decoded GAM record must be equal."""
    if not (os.path.exists(REFBIN) and os.path.exists(REFDUMP)):
        pytest.skip("oracle/_ref not built")
    from graphchainer_b200 import synth
    g = synth.SynthGraph(1_500_000, seed=61)
    gfa, fa = str(tmp_path / "g.gfa"), str(tmp_path / "r.fa")
    with open(gfa, "w") as f:
        f.write(g.gfa())
    synth.write_fasta(fa, synth.simulate_reads(g, 10, (50_000, 100_000), 0.12, seed=62, novel_insertion_frac=0.3))
    idx, ref_gam, out = str(tmp_path / "x.gcidx"), str(tmp_path / "ref.gam"), str(tmp_path / "out.gam")
    subprocess.run([REFDUMP, "-t", "1", "-g", gfa, "--gc-index", idx], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    subprocess.run([REFBIN, "-t", str(os.cpu_count() or 8), "-g", gfa, "-f", fa, "-a", ref_gam], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    subprocess.run([DRIVER, "--gc-index", idx, "-f", fa, "-a", out, "-t", str(min(32, os.cpu_count() or 8))], check=True, stdout=subprocess.DEVNULL)
    a, b = gam.read_gam(out), gam.read_gam(ref_gam)
    assert len(b) == 10
    diffs = gam.diff_gam(a, b)
    assert not diffs, diffs


@pytest.mark.gpu
def test_gpu_pipeline_matches_reference_on_high_width_hifi_reads(tmp_path):
    """BASELINE config-5 shape: extra haplotype alleles (MPC width 4), 20 kb reads at 1 % error, fragments every 18 bp
    (what --sampling-step 0.5 sets: overlapping fragments, almost all anchored, the I-type chaining term matters);
    unmodified reference live on the CPU vs the GPU pipeline, decoded GAM records compared field by field."""
    if not (os.path.exists(REFBIN) and os.path.exists(REFDUMP)):
        pytest.skip("oracle/_ref not built")
    from graphchainer_b200 import synth
    g = synth.SynthGraph(400_000, seed=71, extra_alleles=2, mean_spacing=25)
    gfa, fa = str(tmp_path / "g.gfa"), str(tmp_path / "r.fa")
    with open(gfa, "w") as f:
        f.write(g.gfa())
    synth.write_fasta(fa, synth.simulate_reads(g, 24, 20_000, 0.01, seed=72, novel_insertion_frac=0.25))
    idx, ref_gam, out = str(tmp_path / "x.gcidx"), str(tmp_path / "ref.gam"), str(tmp_path / "out.gam")
    subprocess.run([REFDUMP, "-t", "1", "-g", gfa, "--gc-index", idx], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    subprocess.run([REFBIN, "-t", str(os.cpu_count() or 8), "-g", gfa, "-f", fa, "-a", ref_gam, "--colinear-split-gap", "18"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    subprocess.run([DRIVER, "--gc-index", idx, "-f", fa, "-a", out, "-t", str(min(32, os.cpu_count() or 8)), "--sampling-step", "0.5"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    a, b = gam.read_gam(out), gam.read_gam(ref_gam)
    assert len(b) == 24
    diffs = gam.diff_gam(a, b)
    assert not diffs, diffs
