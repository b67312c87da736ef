"""End-to-end parity of the per-read pipeline: decoded GAM messages (every field, mapping and
edit) must equal the output of the UNMODIFIED reference program for the same inputs.
Golden GAMs: tests/golden/*.gam written by oracle/_ref/GraphChainer_ref (make_golden.py)."""
import os
import re
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


def _edge_case(driver, golden_files, tmp_path, reads, tag):
    idx, _ = golden_files["tiny"]
    out = str(tmp_path / "out.gam")
    run = subprocess.run([driver, "--gc-index", idx, "-f", os.path.join(GOLDEN, reads), "-a", out, "-t", "4", "--gc-quiet"], check=True, capture_output=True, text=True)
    ours, ref = gam.read_gam(out), gam.read_gam(os.path.join(GOLDEN, tag + ".gam"))
    assert sorted(ours) == sorted(ref)          # the same reads get a record (none for the 10-base, random, poly-A ... reads)
    diffs = gam.diff_gam(ours, ref)
    assert not diffs, diffs
    want = [l for l in open(os.path.join(GOLDEN, tag + ".txt")).read().splitlines() if l]
    got = [l for l in run.stdout.splitlines() if l]
    assert got[got.index("Alignment finished"):] == want[want.index("Alignment finished"):]


@pytest.mark.parametrize("reads,tag", [("edge.fa", "edge"), ("edge.fq.gz", "edge_fq")])
def test_ragged_reads_match_reference(driver_sim, golden_files, tmp_path, reads, tag):
    """tests/golden/edge.fa: lower-case reads, reads of 10 / 34 / 35 / 36 / 64 / 65 / 100 bases (around the fragment length and the
    64-row slice), a random and a poly-A read (no alignment), a run of N inside a read, one name used twice, a name with blanks;
    edge.fq.gz: the first eight as gzipped FASTQ.  Records and the summary lines of the unmodified reference."""
    _edge_case(driver_sim, golden_files, tmp_path, reads, tag)


@pytest.mark.parametrize("content", ["", ">only_header\n\n>r2\nACGT\n"])
def test_empty_read_files_match_reference(driver_sim, golden_files, tmp_path, content):
    """no reads at all / a header without bases and a 4-base read: the (empty) GAM file and the summary of the reference, run live"""
    if not os.path.exists(REFBIN):
        pytest.skip("oracle/_ref not built")
    idx, _ = golden_files["tiny"]
    fa, out, ref_out = str(tmp_path / "r.fa"), str(tmp_path / "out.gam"), str(tmp_path / "ref.gam")
    with open(fa, "w") as f:
        f.write(content)
    ref = subprocess.run([REFBIN, "-t", "1", "-g", os.path.join(GOLDEN, "tiny.gfa"), "-f", fa, "-a", ref_out], check=True, capture_output=True, text=True)
    run = subprocess.run([driver_sim, "--gc-index", idx, "-f", fa, "-a", out, "-t", "2", "--gc-quiet"], check=True, capture_output=True, text=True)
    assert gam.read_gam(out) == gam.read_gam(ref_out) == {}
    want, got = [l for l in ref.stdout.splitlines() if l], [l for l in run.stdout.splitlines() if l]
    assert got[got.index("Alignment finished"):] == want[want.index("Alignment finished"):]


@pytest.mark.gpu
@pytest.mark.parametrize("reads,tag", [("edge.fa", "edge"), ("edge.fq.gz", "edge_fq")])
def test_gpu_ragged_reads_match_reference(golden_files, tmp_path, reads, tag):
    assert os.path.exists(DRIVER), "GraphChainerB200 not built (run __graft_entry__.build())"
    _edge_case(DRIVER, golden_files, tmp_path, reads, tag)


def test_no_colinear_chaining_mode_matches_reference(driver_sim, golden_files, tmp_path):
    """--no-colinear-chaining ("align as in GraphAligner", AlignerMain.cpp:108,198): GAM and the summary text of the unmodified
    reference in that mode (tests/golden/tiny_nocc.gam / .txt, make_golden.py)."""
    idx, _ = golden_files["tiny"]
    out = str(tmp_path / "out.gam")
    run = subprocess.run([driver_sim, "--gc-index", idx, "-f", os.path.join(GOLDEN, "tiny.fa"), "-a", out, "-t", "4", "--gc-quiet", "--no-colinear-chaining"], check=True, capture_output=True, text=True)
    ref = gam.read_gam(os.path.join(GOLDEN, "tiny_nocc.gam"))
    diffs = gam.diff_gam(gam.read_gam(out), ref)
    assert not diffs, diffs
    assert gam.diff_gam(ref, gam.read_gam(os.path.join(GOLDEN, "tiny.gam"))), "the mode must change something on this input"
    want = [l for l in open(os.path.join(GOLDEN, "tiny_nocc.txt")).read().splitlines() if l]
    got = [l for l in run.stdout.splitlines() if l]
    assert "Co-linear chaining off" in got
    assert got[got.index("Alignment finished"):] == want[want.index("Alignment finished"):]


def test_vg_graph_input_matches_reference_gam(driver_sim, tmp_path):
    """The graph given as a .vg stream (tests/golden/tiny_vg.vg: the tiny graph with sparse node ids in two gzip members, made by
    make_golden.gfa_to_vg): index built by gc_buildindex, GAM identical to the unmodified reference's on the same .vg
    (StreamVGGraphFromFile, BigraphToDigraph.cpp:134-179) -- vg node ids and names in every Position."""
    builder = str(tmp_path / "gc_buildindex")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fopenmp", "-Wno-sign-compare", "-o", builder, os.path.join(ROOT, "graphchainer_b200", "csrc", "gc_buildindex.cpp"), "-lz"], check=True)
    idx, out = str(tmp_path / "tiny_vg.gcidx"), str(tmp_path / "out.gam")
    subprocess.run([builder, os.path.join(GOLDEN, "tiny_vg.vg"), idx, "--quiet"], check=True, stdout=subprocess.DEVNULL)
    subprocess.run([driver_sim, "--gc-index", idx, "-f", os.path.join(GOLDEN, "tiny.fa"), "-a", out, "-t", "4", "--gc-quiet"], check=True, stdout=subprocess.DEVNULL)
    ours, ref = gam.read_gam(out), gam.read_gam(os.path.join(GOLDEN, "tiny_vg.gam"))
    assert len(ref) == 8
    diffs = gam.diff_gam(ours, ref)
    assert not diffs, diffs
    # the ids are the .vg's (10, 13, 16, ...), not the dense ids the GFA path gives the same segments
    node_ids = {int(m.get("node_id", 0)) for alns in ref.values() for a in alns for m in a["mappings"]}
    assert node_ids and all(i >= 10 and (i - 10) % 3 == 0 for i in node_ids)


_PROGRESS = re.compile(r"\d+ (\S+) len=(\d+) : chained (\d+) / (\d+) anchors, actual (\d+) bps, time \S+ \S+ \S+  score=(\d+) long_edit_distance=(\d+) one_node_overlaps=(\d+) / (\d+)")


def _progress_lines(stderr: str):
    """{read: (len, chained, anchors, path bp, score, long_edit_distance, overlaps now, overlaps all)} from --short-verbose output (Aligner.cpp:909-915)."""
    out = {}
    for line in stderr.splitlines():
        m = _PROGRESS.search(line)
        if m:
            out[m.group(1)] = tuple(int(x) for x in m.groups()[1:])
    return out


def _json_and_progress_parity(exe, golden_files, tmp_path, name):
    """JSON output (writeJSONToQueue, Aligner.cpp:283-298) line for line and the --short-verbose progress fields (anchors, chained,
    path bp, score, long_edit_distance, one-node overlaps: BASELINE.md 3.5) against the unmodified reference run on the same input."""
    idx, _ = golden_files[name]
    gfa, fa = os.path.join(GOLDEN, name + ".gfa"), os.path.join(GOLDEN, name + ".fa")
    ref_json, out_json = str(tmp_path / "ref.json"), str(tmp_path / "out.json")
    ref = subprocess.run([REFBIN, "-t", "1", "-g", gfa, "-f", fa, "-a", ref_json, "--short-verbose"], check=True, capture_output=True, text=True)
    ours = subprocess.run([exe, "--gc-index", idx, "-f", fa, "-a", out_json, "-t", "1", "--short-verbose"], check=True, capture_output=True, text=True)
    a, b = sorted(open(out_json).read().splitlines()), sorted(open(ref_json).read().splitlines())
    assert len(b) > 0 and a == b, f"{name}: JSON lines differ ({len(a)} vs {len(b)})"
    pa, pb = _progress_lines(ours.stderr), _progress_lines(ref.stderr)
    assert pb and set(pa) == set(pb), (sorted(pa), sorted(pb))
    has_long = set(re.findall(r"Aligned long read(\S+) with long_edit_distance", ref.stderr))  # the reference prints an uninitialised value otherwise
    for read, want in pb.items():
        got = pa[read]
        if read not in has_long:
            got, want = got[:5] + got[6:], want[:5] + want[6:]
        assert got == want, f"{name} {read}: progress fields {got} vs the reference's {want}"


@pytest.mark.parametrize("name", ["c1", "tiny"])
def test_host_pipeline_json_and_progress_lines_match_reference(driver_sim, golden_files, tmp_path, name):
    if not os.path.exists(REFBIN):
        pytest.skip("oracle/_ref/GraphChainer_ref not built")
    _json_and_progress_parity(driver_sim, golden_files, tmp_path, name)


def _gaf_parity(exe, golden_files, tmp_path, name):
    """GAF output (GraphAlignerGAFAlignment::traceToAlignment, src/GraphAlignerGAFAlignment.h:37-205): every column of every line
    (path, path length / start / end, matches, block length, NM, dv, id, CIGAR) against the unmodified reference."""
    idx, _ = golden_files[name]
    gfa, fa = os.path.join(GOLDEN, name + ".gfa"), os.path.join(GOLDEN, name + ".fa")
    ref_gaf, out_gaf = str(tmp_path / "ref.gaf"), str(tmp_path / "out.gaf")
    subprocess.run([REFBIN, "-t", "1", "-g", gfa, "-f", fa, "-a", ref_gaf], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    subprocess.run([exe, "--gc-index", idx, "-f", fa, "-a", out_gaf, "-t", "2"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    a, b = sorted(open(out_gaf).read().splitlines()), sorted(open(ref_gaf).read().splitlines())
    assert len(b) > 0 and a == b, f"{name}: GAF lines differ ({len(a)} vs {len(b)})"


@pytest.mark.parametrize("name", ["c1", "tiny"])
def test_host_pipeline_gaf_matches_reference(driver_sim, golden_files, tmp_path, name):
    if not os.path.exists(REFBIN):
        pytest.skip("oracle/_ref/GraphChainer_ref not built")
    _gaf_parity(driver_sim, golden_files, tmp_path, name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c1", "tiny"])
def test_gpu_pipeline_gaf_matches_reference(golden_files, tmp_path, name):
    if not os.path.exists(REFBIN):
        pytest.skip("oracle/_ref/GraphChainer_ref not built")
    _gaf_parity(DRIVER, golden_files, tmp_path, name)


def test_read_that_hits_an_assertion_class_state_is_dropped(driver_sim, golden_files, tmp_path):
    """A work item that reaches a state the reference guards with assert() (GCGPU_ITEM_INTERNAL) in the whole-read pass drops that
    read only, as the reference's catch block does (Aligner.cpp:585-592: alignments cleared, cont = true, every fragment skipped):
    no record for it, every other read unchanged.  In the fragment pass (Aligner.cpp:695-703) the read keeps the anchors
    collected before the failing fragment and is still written.  The state is injected through the C-ABI test double."""
    idx, _ = golden_files["tiny"]
    fa = os.path.join(GOLDEN, "tiny.fa")
    golden = gam.read_gam(os.path.join(GOLDEN, "tiny.gam"))
    victim = "read_2"
    assert victim in golden
    out = str(tmp_path / "out.gam")
    run = subprocess.run([driver_sim, "--gc-index", idx, "-f", fa, "-a", out, "-t", "2"], check=True, capture_output=True, text=True, env=dict(os.environ, GCGPU_SIM_FAIL="s1:2"))
    got = gam.read_gam(out)
    assert victim not in got and "alignment failed (assertion!)" in run.stderr and "Alignment broke with some reads" in run.stdout
    rest = {k: v for k, v in golden.items() if k != victim}
    assert not gam.diff_gam(got, rest)
    run = subprocess.run([driver_sim, "--gc-index", idx, "-f", fa, "-a", out, "-t", "2"], check=True, capture_output=True, text=True, env=dict(os.environ, GCGPU_SIM_FAIL="s2:2:700"))
    got = gam.read_gam(out)
    assert victim in got and "Alignment broke with some reads" in run.stdout
    assert not gam.diff_gam({k: v for k, v in got.items() if k != victim}, rest)
    # a character outside the IUPAC alphabet: the reference's Complement() asserts (CommonUtils.cpp:131-133; the unmodified program
    # aborts as a whole there) -- here that read alone is dropped
    lines = open(fa).read().split("\n")
    k = lines.index(">" + victim) + 1
    lines[k] = lines[k][:300] + "X" + lines[k][301:]
    bad_fa = str(tmp_path / "bad.fa")
    open(bad_fa, "w").write("\n".join(lines))
    run = subprocess.run([driver_sim, "--gc-index", idx, "-f", bad_fa, "-a", out, "-t", "2"], check=True, capture_output=True, text=True)
    got = gam.read_gam(out)
    assert victim not in got and not gam.diff_gam(got, rest)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c1", "tiny"])
def test_gpu_pipeline_json_and_progress_lines_match_reference(golden_files, tmp_path, name):
    if not os.path.exists(REFBIN):
        pytest.skip("oracle/_ref/GraphChainer_ref not built")
    assert os.path.exists(DRIVER), "GraphChainerB200 not built (run __graft_entry__.build())"
    _json_and_progress_parity(DRIVER, golden_files, tmp_path, name)


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
    unmodified reference runs live on the box's CPU (one thread: its multi-threaded runs are not run-to-run deterministic,
    profiles/r04h_reference_nondeterminism.txt), the GPU pipeline must give the same GAM."""
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
    subprocess.run([REFBIN, "-t", "1", "-g", gfa, "-f", fa, "-a", ref_gam], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    subprocess.run([DRIVER, "--gc-index", idx, "-f", fa, "-a", out, "-t", str(min(32, os.cpu_count() or 8))], check=True, stdout=subprocess.DEVNULL)
    a, b = gam.read_gam(out), gam.read_gam(ref_gam)
    assert len(b) == 300
    diffs = gam.diff_gam(a, b)
    assert not diffs, diffs
    # the same reads "as in GraphAligner" (--no-colinear-chaining): the whole-read pass and its GreedyLength selection only
    subprocess.run([REFBIN, "-t", "1", "-g", gfa, "-f", fa, "-a", ref_gam, "--no-colinear-chaining"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    subprocess.run([DRIVER, "--gc-index", idx, "-f", fa, "-a", out, "-t", str(min(32, os.cpu_count() or 8)), "--no-colinear-chaining"], check=True, stdout=subprocess.DEVNULL)
    a2, b2 = gam.read_gam(out), gam.read_gam(ref_gam)
    assert len(b2) >= 290
    diffs = gam.diff_gam(a2, b2)
    assert not diffs, diffs
    assert gam.diff_gam(b2, b), "the mode changes nothing on this input: the comparison above proves nothing"


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
    subprocess.run([REFBIN, "-t", "1", "-g", gfa, "-f", fa, "-a", ref_gam], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
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
    subprocess.run([REFBIN, "-t", "1", "-g", gfa, "-f", fa, "-a", ref_gam, "--colinear-split-gap", "18"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    subprocess.run([DRIVER, "--gc-index", idx, "-f", fa, "-a", out, "-t", str(min(32, os.cpu_count() or 8)), "--sampling-step", "0.5"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    a, b = gam.read_gam(out), gam.read_gam(ref_gam)
    assert len(b) == 24
    diffs = gam.diff_gam(a, b)
    assert not diffs, diffs
