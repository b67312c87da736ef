"""Index build parity: every array the per-read path consumes (split-node numbering, neighbour
order, component numbers, chain positions, MPC paths / backward links / topological ids,
minimizer position lists and maxCount) must equal a dump of the reference's own structures
(tests/golden/*.gcidx.gz, written by oracle/_ref/gc_refdump from the unmodified reference)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from graphchainer_b200 import lib


@pytest.fixture(scope="session")
def buildindex(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("bld") / "gc_buildindex")
    subprocess.run(["g++", "-O2", "-std=c++17", "-Wno-sign-compare", "-o", out, os.path.join(ROOT, "graphchainer_b200", "csrc", "gc_buildindex.cpp")], check=True)
    return out


def _partition(a):
    seen = {}
    return [seen.setdefault(x, len(seen)) for x in a.tolist()]


def _minimizers(idx):
    st, ps = idx["mzKmerStart"], idx["mzPositions"]
    return {int(k): ps[st[i]:st[i + 1]].tolist() for i, k in enumerate(idx["mzKmers"])}


def compare_index(mine: dict, ref: dict):
    bad = []
    for k in ref:
        if k in ("chainNumber", "mzKmers", "mzKmerStart", "mzPositions", "mzBucketStart"):
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
    bad = compare_index(lib.read_gcidx(out), lib.read_gcidx(golden_files[name][0]))
    assert not bad, f"arrays differing from the reference: {bad}"


def test_builder_rejects_cycles(buildindex, tmp_path):
    gfa = str(tmp_path / "cyc.gfa")
    with open(gfa, "w") as f:
        f.write("S\t1\tACGTACGTACGTACGTACGTAC\nS\t2\tTTGACCATGACAGTACCATGGA\nL\t1\t+\t2\t+\t0M\nL\t2\t+\t1\t+\t0M\n")
    r = subprocess.run([buildindex, gfa, str(tmp_path / "x.gcidx"), "--quiet"], capture_output=True, text=True)
    assert r.returncode != 0 and "directed cycle" in r.stderr
