"""K2 (co-linear chaining) and K3 (edlib-style NW distance / path) parity against the records
of the unmodified reference: identical chain index lists, identical edit distances and
identical edit-operation strings."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import REFDUMP
import stages


@pytest.mark.parametrize("mode", ["k2", "k3", "k3w"])
def test_logic_on_cpu_matches_golden(hostsim, golden_files, mode):
    for name, (idx, st) in golden_files.items():
        out = subprocess.run([hostsim, mode, idx, st], capture_output=True, text=True)
        assert out.returncode == 0, f"{name}: {out.stdout} {out.stderr}"
        rep = json.loads(out.stdout.strip().splitlines()[-1])
        assert rep["mismatches"] == 0 and rep["items"] > 0


def replay_k2_k3_on_gpu(idx_path, st_path):
    from graphchainer_b200 import lib
    index = lib.read_gcidx(idx_path)
    ctx = lib.Context(index)
    reads = stages.parse(st_path)
    # ---- K2
    anchors, offs, want_chain = [], [0], []
    for r in reads:
        if "chain" not in r:
            continue
        for a in r["anchors"]:
            anchors.append((a["path"][0], a["path"][-1], a["x"], a["y"]))
        offs.append(len(anchors))
        want_chain.append(r["chain"])
    chain, clen, cscore = ctx.chain(np.array(anchors, dtype=lib.ANCHOR), np.array(offs, dtype=np.uint64))
    bad_chain = 0
    for k, want in enumerate(want_chain):
        got = chain[offs[k]:offs[k] + clen[k]].tolist()
        bad_chain += got != want
    # ---- K3
    buf = bytearray()
    items, want_nw = [], []
    for r in reads:
        t_off = len(buf)
        buf += r["seq"].encode()
        if "ga_pathseq" in r:
            q_off = len(buf)
            buf += r["ga_pathseq"].encode()
            items.append((q_off, t_off, len(r["ga_pathseq"]), len(r["seq"]), 0, 0))
            want_nw.append((r["long_edit_distance"], None))
        if r.get("edlib"):
            q_off = len(buf)
            buf += r["pathseq"].encode()
            items.append((q_off, t_off, len(r["pathseq"]), len(r["seq"]), 0, 1))
            want_nw.append((r["edlib"]["distance"], r["edlib"]["ops"]))
    res, ops = ctx.nw(bytes(buf), np.array(items, dtype=lib.NW_ITEM))
    bad_nw = 0
    for k, (d, o) in enumerate(want_nw):
        ok = res[k]["status"] == 0 and res[k]["distance"] == d
        if ok and o is not None:
            got = "".join(chr(48 + int(c)) for c in ops[res[k]["ops_offset"]:res[k]["ops_offset"] + res[k]["ops_len"]])
            ok = got == o
        bad_nw += not ok
    ctx.close()
    return len(want_chain), bad_chain, len(want_nw), bad_nw


@pytest.mark.gpu
def test_k2_k3_gpu_match_golden(golden_files):
    for name, (idx, st) in golden_files.items():
        nc, bc, nn, bn = replay_k2_k3_on_gpu(idx, st)
        assert nc > 0 and bc == 0, f"{name}: {bc}/{nc} chains differ from the reference"
        assert nn > 0 and bn == 0, f"{name}: {bn}/{nn} NW alignments differ from the reference"


@pytest.mark.gpu
def test_k2_k3_gpu_match_reference_on_high_width_overlapping_fragments(tmp_path):
    """BASELINE config-5 shape: extra alleles (MPC width 4) + --colinear-split-gap 18 (overlapping
    fragments exercise the I-type chaining term); reference run live through gc_refdump."""
    if not os.path.exists(REFDUMP):
        pytest.skip("oracle/_ref/gc_refdump not built")
    from graphchainer_b200 import synth
    g = synth.SynthGraph(200_000, seed=31, extra_alleles=2, mean_spacing=25)
    gfa, fa = str(tmp_path / "g.gfa"), str(tmp_path / "r.fa")
    with open(gfa, "w") as f:
        f.write(g.gfa())
    synth.write_fasta(fa, synth.simulate_reads(g, 30, 12000, 0.02, seed=32, novel_insertion_frac=0.2))
    idx, st = str(tmp_path / "x.gcidx"), str(tmp_path / "x.stages")
    subprocess.run([REFDUMP, "-t", "1", "-g", gfa, "-f", fa, "--colinear-split-gap", "18", "--gc-index", idx, "--gc-stages", st], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    nc, bc, nn, bn = replay_k2_k3_on_gpu(idx, st)
    assert nc == 30 and bc == 0
    assert nn >= 30 and bn == 0


@pytest.mark.gpu
@pytest.mark.parametrize("force", ["thread", "dfs"])
def test_k3_fallback_kernels_match_golden(golden_files, force):
    """The K3 fallback forms -- thread-per-alignment distance and edit-path kernels (bands beyond the warp kernels' register
    budget) and the depth-first warp edit-path kernel (alignments the level-parallel form gives up on) -- forced for every
    alignment through GCGPU_K3_FORCE: same distances and operation strings as the reference."""
    os.environ["GCGPU_K3_FORCE"] = force
    try:
        for name, (idx, st) in golden_files.items():
            nc, bc, nn, bn = replay_k2_k3_on_gpu(idx, st)
            assert nn > 0 and bn == 0, f"{name} ({force}): {bn}/{nn} NW alignments differ from the reference"
    finally:
        del os.environ["GCGPU_K3_FORCE"]


@pytest.mark.gpu
def test_k2_pairwise_form_matches_golden(golden_files):
    """The pairwise form of K2 (the fallback for reads whose anchors have different lengths) forced for every read."""
    os.environ["GCGPU_K2_FORCE"] = "pairwise"
    try:
        for name, (idx, st) in golden_files.items():
            nc, bc, nn, bn = replay_k2_k3_on_gpu(idx, st)
            assert nc > 0 and bc == 0, f"{name}: {bc}/{nc} chains differ from the reference"
    finally:
        del os.environ["GCGPU_K2_FORCE"]


def _live_k2_k3(tmp_path, graph, reads, extra=()):
    from graphchainer_b200 import synth
    gfa, fa = str(tmp_path / "g.gfa"), str(tmp_path / "r.fa")
    with open(gfa, "w") as f:
        f.write(graph.gfa())
    synth.write_fasta(fa, reads)
    idx, st = str(tmp_path / "x.gcidx"), str(tmp_path / "x.stages")
    subprocess.run([REFDUMP, "-t", "1", "-g", gfa, "-f", fa, *extra, "--gc-index", idx, "--gc-stages", st], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return replay_k2_k3_on_gpu(idx, st)


@pytest.mark.gpu
def test_k2_gpu_matches_reference_at_mpc_width_16(tmp_path):
    """16 parallel alleles per variant site (MPC width 16 per strand): every anchor start has 16 backward links, every shared node
    lies on 16 paths -- the per-path trees of the sweep form carry the chaining; overlapping fragments (--colinear-split-gap 18)."""
    if not os.path.exists(REFDUMP):
        pytest.skip("oracle/_ref/gc_refdump not built")
    from graphchainer_b200 import synth
    g = synth.SynthGraph(120_000, seed=91, extra_alleles=14, mean_spacing=30)
    nc, bc, nn, bn = _live_k2_k3(tmp_path, g, synth.simulate_reads(g, 20, 10_000, 0.02, seed=92, novel_insertion_frac=0.2), ("--colinear-split-gap", "18"))
    assert nc == 20 and bc == 0, f"{bc}/{nc} chains differ from the reference"
    assert bn == 0


@pytest.mark.gpu
def test_k2_gpu_matches_reference_on_ultralong_anchor_sets(tmp_path):
    """BASELINE config-4 shape for K2: 70-90 kb reads, ~2 k anchors per read."""
    if not os.path.exists(REFDUMP):
        pytest.skip("oracle/_ref/gc_refdump not built")
    from graphchainer_b200 import synth
    g = synth.SynthGraph(1_200_000, seed=93)
    nc, bc, nn, bn = _live_k2_k3(tmp_path, g, synth.simulate_reads(g, 6, (70_000, 90_000), 0.10, seed=94, novel_insertion_frac=0.3))
    assert nc == 6 and bc == 0, f"{bc}/{nc} chains differ from the reference"
    assert bn == 0
