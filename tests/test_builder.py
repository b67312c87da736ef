"""Index build parity: every array the per-read path consumes bit for bit (split-node numbering, neighbour order, component
numbers, chain positions, minimizer position lists and maxCount) must equal a dump of the reference's own structures
(tests/golden/*.gcidx.gz, written by oracle/_ref/gc_refdump from the unmodified reference).  The path-cover index is an own
construction: chaining reads only reachability from it, so it is checked for exactly that (index-reachability == graph
reachability), for covering every node and for having the reference's (minimum) width."""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from graphchainer_b200 import lib


@pytest.fixture(scope="session")
def buildindex(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("bld") / "gc_buildindex")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fopenmp", "-Wno-sign-compare", "-o", out, os.path.join(ROOT, "graphchainer_b200", "csrc", "gc_buildindex.cpp"), "-lz"], check=True)
    return out


def _partition(a):
    seen = {}
    return [seen.setdefault(x, len(seen)) for x in a.tolist()]


def _minimizers(idx):
    st, ps = idx["mzKmerStart"], idx["mzPositions"]
    return {int(k): ps[st[i]:st[i + 1]].tolist() for i, k in enumerate(idx["mzKmers"])}


MPC_ARRAYS = ("compMap", "compIdx", "compStart", "compIds", "topoIds", "mpcWidth", "mpcPathStart", "mpcNodes", "pathsStart", "pathsK", "backStart", "backNode", "backK")


def _index_reaches(idx, e, s):
    """gc_k2_reaches (graphchainer_b200/csrc/gc_k2.cuh) on the index arrays: does node e reach node s according to the path cover?"""
    if e == s:
        return True
    c = int(idx["compMap"][s])
    if int(idx["compMap"][e]) != c:
        return False
    base = int(idx["compStart"][c])
    ge, gs = base + int(idx["compIdx"][e]), base + int(idx["compIdx"][s])
    topo_e = int(idx["topoIds"][ge])
    paths_e = set(idx["pathsK"][idx["pathsStart"][ge]:idx["pathsStart"][ge + 1]].tolist())
    for b in range(int(idx["backStart"][gs]), int(idx["backStart"][gs + 1])):
        if topo_e <= int(idx["topoIds"][base + int(idx["backNode"][b])]) and int(idx["backK"][b]) in paths_e:
            return True
    return False


def _true_reach_set(idx, e):
    """nodes reachable from e in the split-node graph (depth-first over outNbr)"""
    seen, stack = {e}, [e]
    while stack:
        v = stack.pop()
        for w in idx["outNbr"][idx["outStart"][v]:idx["outStart"][v + 1]].tolist():
            if w not in seen:
                seen.add(w)
                stack.append(w)
    return seen


def check_path_cover(idx, sources, targets):
    """The path-cover index is correct iff index-reachability equals graph reachability (chaining reads nothing else from it,
    so ANY cover of all nodes gives the reference's chains); also every node lies on a path and the width is what is reported."""
    n = len(idx["compMap"])
    assert all(idx["pathsStart"][g + 1] > idx["pathsStart"][g] for g in range(n)), "a node lies on no path"
    bad = 0
    for e in sources:
        truth = _true_reach_set(idx, e)
        for s in targets:
            bad += _index_reaches(idx, e, s) != (s in truth)
    return bad


def compare_index(mine: dict, ref: dict):
    bad = []
    for k in ref:
        if k in ("chainNumber", "mzKmers", "mzKmerStart", "mzPositions", "mzBucketStart") or k in MPC_ARRAYS:
            continue
        if k not in mine or mine[k].shape != ref[k].shape or not (mine[k] == ref[k]).all():
            bad.append(k)
    # chain ids are union-find representatives: only the partition is observable (seed clustering)
    if _partition(mine["chainNumber"]) != _partition(ref["chainNumber"]):
        bad.append("chainNumber")
    # the reference stores k-mers in minimal-perfect-hash order; the map k-mer -> position list is what matters
    if _minimizers(mine) != _minimizers(ref):
        bad.append("minimizers")
    return bad


@pytest.mark.parametrize("name", ["c1", "tiny"])
def test_builder_matches_reference_index(buildindex, golden_files, tmp_path, name):
    out = str(tmp_path / "my.gcidx")
    subprocess.run([buildindex, os.path.join(GOLDEN, name + ".gfa"), out, "--quiet"], check=True)
    mine, ref = lib.read_gcidx(out), lib.read_gcidx(golden_files[name][0])
    bad = compare_index(mine, ref)
    assert not bad, f"arrays differing from the reference: {bad}"
    # path cover: own construction (any minimum cover serves), same width as the reference's, and exactly the graph's reachability
    assert mine["mpcWidth"].tolist() == ref["mpcWidth"].tolist()
    rng = np.random.default_rng(5)
    n = len(mine["compMap"])
    sources = rng.choice(n, size=min(n, 40), replace=False).tolist()
    targets = list(range(n)) if n <= 400 else rng.choice(n, size=400, replace=False).tolist()
    assert check_path_cover(mine, sources, targets) == 0
    assert check_path_cover(ref, sources[:10], targets) == 0   # the checker itself, on the reference's index


def test_builder_path_cover_on_a_wide_graph(buildindex, tmp_path):
    """16 parallel alleles per site: minimum width 16 per strand, reachability through the index exact."""
    from graphchainer_b200 import synth
    g = synth.SynthGraph(30_000, seed=17, extra_alleles=14, mean_spacing=30)
    gfa, out = str(tmp_path / "w.gfa"), str(tmp_path / "w.gcidx")
    with open(gfa, "w") as f:
        f.write(g.gfa())
    subprocess.run([buildindex, gfa, out, "--quiet"], check=True)
    idx = lib.read_gcidx(out)
    assert idx["mpcWidth"].tolist() == [16, 16]
    rng = np.random.default_rng(6)
    n = len(idx["compMap"])
    assert check_path_cover(idx, rng.choice(n, size=30, replace=False).tolist(), rng.choice(n, size=500, replace=False).tolist()) == 0


def test_builder_rejects_cycles(buildindex, tmp_path):
    gfa = str(tmp_path / "cyc.gfa")
    with open(gfa, "w") as f:
        f.write("S\t1\tACGTACGTACGTACGTACGTAC\nS\t2\tTTGACCATGACAGTACCATGGA\nL\t1\t+\t2\t+\t0M\nL\t2\t+\t1\t+\t0M\n")
    r = subprocess.run([buildindex, gfa, str(tmp_path / "x.gcidx"), "--quiet"], capture_output=True, text=True)
    assert r.returncode != 0 and "directed cycle" in r.stderr
