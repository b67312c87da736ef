"""K1 (graph extension) parity: every extension the reference ran on the golden inputs
-- both directions of every seed of the whole-read pass (S1) and of the 35-bp fragment
pass (S2) -- must give the identical score and the identical trace (node, offset,
seqPos, nodeSwitch per cell), bit for bit.  Reference records: tests/golden/*.stages.gz,
produced by tests/golden/make_golden.py from the unmodified reference."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import REFDUMP, ROOT
import stages


@pytest.mark.parametrize("mode", ["k1", "k1s", "k1q"])
def test_k1_logic_on_cpu_matches_golden(hostsim, golden_files, mode):
    """The device functions, compiled for the host, replay the golden K1 records (k1: the thread/warp-per-item form of
    gc_k1.cuh; k1s: the lane-per-item form of gc_k1s.cuh as a warp of one lane, Eq masks from bit planes for every other item;
    k1q: all records of a file through ONE lane one after the other, the way a lane of the GPU kernels takes item after item)."""
    assert golden_files, "golden fixtures missing"
    for name, (idx, st) in golden_files.items():
        out = subprocess.run([hostsim, mode, idx, st], capture_output=True, text=True)
        assert out.returncode == 0, f"{name}: {out.stdout} {out.stderr}"
        rep = json.loads(out.stdout.strip().splitlines()[-1])
        assert rep["mismatches"] == 0 and rep["items"] > 0


def _replay_on_gpu(idx_path, st_path, max_reads=None):
    from graphchainer_b200 import lib
    index = lib.read_gcidx(idx_path)
    ctx = lib.Context(index)
    reads = stages.parse(st_path)
    if max_reads:
        reads = reads[:max_reads]
    seqs, items, expect = [], [], []
    pos = 0
    for r in reads:
        for e in r["ext"]:
            codes = lib.encode(e["seq"])
            node, off = ctx.unitig_node(e["node"], e["offset"])
            items.append((pos, len(codes), node, off, 0))
            seqs.append(codes)
            pos += len(codes)
            expect.append(e)
    items = np.array(items, dtype=lib.EXT_ITEM)
    res, traces = ctx.extend(np.concatenate(seqs) if seqs else np.zeros(0, np.uint8), items)
    bad = 0
    for k, e in enumerate(expect):
        r = res[k]
        if e["failed"]:
            ok = r["status"] == 1
        else:
            t = traces[r["trace_offset"]:r["trace_offset"] + r["trace_len"]]
            want = lib.pack_trace(*e["trace"])
            ok = r["status"] == 0 and r["score"] == e["score"] and len(t) == len(want) and bool((t == want).all())
        bad += not ok
    ctx.close()
    return len(expect), bad, int(res["columns"].sum())


@pytest.mark.gpu
@pytest.mark.parametrize("form", ["auto", "lane", "lockstep", "slots", "queue"])
def test_k1_gpu_matches_golden(golden_files, form):
    """form: which kernels take the whole-read extensions -- lane-per-item (gc_k1s_*), warp-per-item in lock-step (gc_k1_long_*),
    or the library's per-launch choice (GCGPU_K1_FORM); slots: the fragment items go through a pool of 256 slabs chunk after chunk
    (GCGPU_K1_SHORT_SLOTS; the default pool of 2 M slabs is only exceeded by HiFi-sized batches); queue: see below."""
    if form == "slots":
        os.environ["GCGPU_K1_SHORT_SLOTS"] = "256"
    elif form == "queue":
        # lane-per-item kernels on two blocks: 128 lanes take the launch's items one after the other from the counter (what
        # happens on the full GPU once a launch has more items than resident lanes -- the whole read set as one batch)
        os.environ["GCGPU_K1_FORM"] = "lane"
        os.environ["GCGPU_K1_BLOCKS"] = "2"
    elif form != "auto":
        os.environ["GCGPU_K1_FORM"] = form
    try:
        for name, (idx, st) in golden_files.items():
            n, bad, cols = _replay_on_gpu(idx, st)
            assert n > 0 and bad == 0, f"{name} ({form}): {bad}/{n} extensions differ from the reference"
    finally:
        os.environ.pop("GCGPU_K1_FORM", None)
        os.environ.pop("GCGPU_K1_SHORT_SLOTS", None)
        os.environ.pop("GCGPU_K1_BLOCKS", None)


@pytest.mark.gpu
def test_k1_gpu_matches_reference_on_fresh_synthetic(tmp_path):
    """Bigger seeded case generated on the box: reference run live through oracle/_ref/gc_refdump."""
    if not os.path.exists(REFDUMP):
        pytest.skip("oracle/_ref/gc_refdump not built")
    from graphchainer_b200 import synth
    g = synth.SynthGraph(300_000, seed=21)
    gfa, fa = str(tmp_path / "g.gfa"), str(tmp_path / "r.fa")
    with open(gfa, "w") as f:
        f.write(g.gfa())
    synth.write_fasta(fa, synth.simulate_reads(g, 60, 6000, 0.15, seed=22, novel_insertion_frac=0.1))
    idx, st = str(tmp_path / "x.gcidx"), str(tmp_path / "x.stages")
    subprocess.run([REFDUMP, "-t", "1", "-g", gfa, "-f", fa, "--gc-index", idx, "--gc-stages", st], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    for form in ("lane", "lockstep", "queue"):
        os.environ["GCGPU_K1_FORM"] = "lane" if form == "queue" else form
        if form == "queue":
            os.environ["GCGPU_K1_BLOCKS"] = "3"   # > 5000 items through 192 lanes
        try:
            n, bad, cols = _replay_on_gpu(idx, st)
        finally:
            os.environ.pop("GCGPU_K1_FORM", None)
            os.environ.pop("GCGPU_K1_BLOCKS", None)
        assert n > 5000 and bad == 0, f"{form}: {bad}/{n} extensions differ from the reference"
