"""The gzip members the device makes of the GAM records (gc_gam.cuh: 32 chunks per record, one Huffman code, the chunks' codes
written at their bit offsets, chunk CRCs joined) must be valid gzip: zlib inflates them to the bytes they were made from.
The stages are run lane after lane on the host (tests/hostsim/gz_check.cpp) -- the kernel runs the same functions on a warp."""
import json
import os
import subprocess

from conftest import ROOT


def test_device_gzip_members_inflate_with_zlib(tmp_path):
    exe = str(tmp_path / "gz_check")
    subprocess.run(["g++", "-O1", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "hostsim", "gz_check.cpp"), "-lz"], check=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    rep = json.loads(out.stdout.strip().splitlines()[-1])
    assert rep["bad"] == 0 and rep["cases"] >= 70
