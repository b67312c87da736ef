"""The C-ABI boundary: both shared libraries load on the GPU-less box and export every function
include/*.h declares; the structs mirrored with ctypes/numpy have the C sizes; without a CUDA
device the product fails loudly instead of falling back (no compute is attempted here)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from conftest import ROOT

INCLUDE = os.path.join(ROOT, "include")
LIBS = {"gcgpu.h": os.path.join(ROOT, "graphchainer_b200", "libgcgpu.so"), "gcalign.h": os.path.join(ROOT, "graphchainer_b200", "libgcalign.so")}


def declared_functions(header):
    text = open(os.path.join(INCLUDE, header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gc(?:gpu|align)_[a-z_0-9]+)\s*\(", text)))


@pytest.mark.parametrize("header", sorted(LIBS))
def test_library_exports_every_declared_symbol(header):
    lib_path = LIBS[header]
    assert os.path.exists(lib_path), f"{lib_path} missing: run __graft_entry__.build()"
    names = declared_functions(header)
    assert len(names) >= 5
    lib = C.CDLL(lib_path)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_struct_sizes_match_the_python_mirrors(tmp_path):
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "gcgpu.h"\n#include "gcalign.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(gcgpu_ext_item), sizeof(gcgpu_ext_result), sizeof(gcgpu_nw_item), sizeof(gcgpu_nw_result), sizeof(gcgpu_anchor), sizeof(gcalign_options), sizeof(gcalign_read_summary), sizeof(gcalign_stats), sizeof(gcgpu_graph));return 0;}\n')
    exe = str(tmp_path / "sizes")
    subprocess.run(["gcc", "-I", INCLUDE, "-o", exe, str(src)], check=True)   # the headers are plain C
    sizes = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    from graphchainer_b200 import align, lib
    assert sizes[0] == lib.EXT_ITEM.itemsize and sizes[1] == lib.EXT_RESULT.itemsize
    assert sizes[2] == lib.NW_ITEM.itemsize and sizes[3] == lib.NW_RESULT.itemsize and sizes[4] == lib.ANCHOR.itemsize
    assert sizes[8] == C.sizeof(lib.GraphStruct)
    assert sizes[5] == C.sizeof(align.Options) and sizes[6] == align.SUMMARY.itemsize and sizes[7] == C.sizeof(align.Stats)


def test_no_cpu_fallback_without_a_device(golden_files):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from graphchainer_b200 import align
    idx, _ = golden_files["c1"]
    with pytest.raises(RuntimeError) as e:
        align.Aligner(idx)
    assert "CUDA" in str(e.value) or "device" in str(e.value)
