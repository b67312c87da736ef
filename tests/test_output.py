"""Output path: the driver's own gzip member encoder (graphchainer_b200/csrc/gc_deflate.h, used at
--gc-gzip-level 1) must produce members that zlib inflates back to the exact record bytes -- checked
on the reference's golden GAM records (tests/golden/*.gam, written by the unmodified reference) and
on edge cases (empty/tiny records, long runs, incompressible data, alphabets that need the 15-bit
code length limit).  The decoded-GAM equality of the whole pipeline is in test_pipeline.py."""
import os
import subprocess

from conftest import GOLDEN, ROOT


def test_fast_deflate_round_trips_through_zlib(tmp_path):
    exe = str(tmp_path / "deflate_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "hostsim", "deflate_check.cpp"), "-lz"], check=True)
    out = subprocess.run([exe, os.path.join(GOLDEN, "tiny.gam"), os.path.join(GOLDEN, "c1.gam")], check=True, capture_output=True, text=True).stdout.split()
    res = dict(zip(out[0::2], map(int, out[1::2])))
    assert res["from_files"] >= 2 and res["raw"] > 100000 and res["bad"] == 0 and res["fallback"] == 0
    assert res["out"] < res["raw"]
