"""N > 1 path on the CPU box: two gloo ranks, each aligning its shard (host logic through the C-ABI test
double), records gathered on rank 0 == the reference's golden GAM.  Also the partitioner's invariants."""
import os
import subprocess
import sys

import numpy as np

from conftest import GOLDEN, ROOT
from graphchainer_b200 import shard


def test_length_balanced_shards_partition_and_balance():
    rng = np.random.default_rng(7)
    lengths = rng.integers(50, 100_000, size=1001)
    for world in (1, 2, 4, 8):
        parts = shard.length_balanced_shards(lengths, world)
        allidx = np.concatenate(parts)
        assert sorted(allidx.tolist()) == list(range(len(lengths)))  # a partition
        sums = [int(lengths[p].sum()) for p in parts]
        assert max(sums) - min(sums) <= int(lengths.max())           # within one read of each other
    assert [p.tolist() for p in shard.length_balanced_shards([], 2)] == [[], []]


def test_two_gloo_ranks_reproduce_golden_gam(golden_files, tmp_path):
    lib = str(tmp_path / "libgcalign_sim.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fopenmp", "-fPIC", "-shared", "-Wno-sign-compare", "-o", lib,
                    os.path.join(ROOT, "graphchainer_b200", "csrc", "gc_capi.cpp"), os.path.join(ROOT, "tests", "hostsim", "gcgpu_sim.cpp"), "-lz"], check=True)
    idx, _ = golden_files["tiny"]
    flag = str(tmp_path / "result.txt")
    env = dict(os.environ, OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29613",
           os.path.join(ROOT, "tests", "multirank_worker.py"), idx, os.path.join(GOLDEN, "tiny.fa"), os.path.join(GOLDEN, "tiny.gam"), flag, lib]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert open(flag).read() == "OK", open(flag).read()
