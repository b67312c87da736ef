import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
REFDUMP = os.path.join(ROOT, "oracle", "_ref", "gc_refdump")
REFBIN = os.path.join(ROOT, "oracle", "_ref", "GraphChainer_ref")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def hostsim(tmp_path_factory):
    """tests/hostsim/host_sim compiled with g++: the device algorithms on the CPU (logic check only)."""
    out = str(tmp_path_factory.mktemp("hostsim") / "host_sim")
    src = os.path.join(ROOT, "tests", "hostsim", "host_sim.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", out, src], check=True)
    return out


@pytest.fixture(scope="session")
def golden_files(tmp_path_factory):
    """Decompress the committed golden stage dumps; returns {name: (gcidx_path, stages_path)}."""
    import gzip
    import shutil
    d = tmp_path_factory.mktemp("golden")
    out = {}
    for name in ("c1", "tiny"):
        idx = os.path.join(GOLDEN, f"{name}.gcidx.gz")
        st = os.path.join(GOLDEN, f"{name}.stages.gz")
        if not (os.path.exists(idx) and os.path.exists(st)):
            continue
        paths = []
        for src, suffix in ((idx, ".gcidx"), (st, ".stages")):
            dst = str(d / (name + suffix))
            with gzip.open(src, "rb") as fi, open(dst, "wb") as fo:
                shutil.copyfileobj(fi, fo)
            paths.append(dst)
        out[name] = tuple(paths)
    return out
