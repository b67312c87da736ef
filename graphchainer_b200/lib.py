"""ctypes binding of libgcgpu (include/gcgpu.h) -- the thin Python mirror used by the
tests and bench.py.  The product is the C-ABI library; this module only marshals
numpy arrays into it.  There is no fallback: if the CUDA library is missing or no
GPU is present every entry point raises."""
from __future__ import annotations

import ctypes as C
import os
import struct

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgcgpu.so")

_DT = {1: np.uint8, 4: np.uint32, 5: np.int32, 8: np.uint64}

BASE_CODE = np.zeros(256, dtype=np.uint8)
for _chars, _mask in (("Aa", 1), ("Cc", 2), ("Gg", 4), ("TtUu", 8), ("Rr", 5), ("Yy", 10), ("Kk", 12), ("Mm", 3), ("Ss", 6),
                      ("Ww", 9), ("Bb", 14), ("Dd", 13), ("Hh", 11), ("Vv", 7), ("Nn", 15)):
    for _c in _chars:
        BASE_CODE[ord(_c)] = _mask
_COMP = {"A": "T", "C": "G", "G": "C", "T": "A", "a": "t", "c": "g", "g": "c", "t": "a", "N": "N", "n": "n"}


def encode(seq: str) -> np.ndarray:
    """Read characters -> IUPAC bit masks (bit0 A, bit1 C, bit2 G, bit3 T)."""
    return BASE_CODE[np.frombuffer(seq.encode(), dtype=np.uint8)]


def revcomp(seq: str) -> str:
    return "".join(_COMP.get(c, "N") for c in reversed(seq))


def read_gcidx(path: str) -> dict:
    """Parse a .gcidx file (graphchainer_b200/csrc/gc_index.h) into {name: ndarray}."""
    out = {}
    with open(path, "rb") as f:
        data = f.read()
    if data[:8] != b"GCIDX001":
        raise ValueError(f"{path}: not a gcidx file")
    pos = 8
    while pos < len(data):
        (nl,) = struct.unpack_from("<I", data, pos)
        pos += 4
        name = data[pos:pos + nl].decode()
        pos += nl
        dtype = data[pos]
        pos += 1
        (count,) = struct.unpack_from("<Q", data, pos)
        pos += 8
        dt = np.dtype(_DT[dtype])
        out[name] = np.frombuffer(data, dtype=dt, count=count, offset=pos).copy()
        pos += count * dt.itemsize
    return out


class GraphStruct(C.Structure):
    _fields_ = [("num_nodes", C.c_uint32),
                ("node_length", C.c_void_p), ("node_seq", C.c_void_p),
                ("in_start", C.c_void_p), ("in_nbr", C.c_void_p), ("out_start", C.c_void_p), ("out_nbr", C.c_void_p),
                ("component_number", C.c_void_p), ("linearizable", C.c_void_p),
                ("num_components", C.c_uint32),
                ("comp_map", C.c_void_p), ("comp_idx", C.c_void_p), ("comp_start", C.c_void_p), ("topo_ids", C.c_void_p),
                ("paths_start", C.c_void_p), ("paths_k", C.c_void_p), ("back_start", C.c_void_p), ("back_node", C.c_void_p), ("back_k", C.c_void_p),
                ("node_ids", C.c_void_p), ("node_offset", C.c_void_p), ("num_orig", C.c_uint32),
                ("orig_ids", C.c_void_p), ("orig_start", C.c_void_p), ("orig_nodes", C.c_void_p), ("orig_size", C.c_void_p)]


class ParamsStruct(C.Structure):
    _fields_ = [("initial_bandwidth", C.c_int32)]


EXT_ITEM = np.dtype([("seq_offset", "<u8"), ("seq_len", "<i4"), ("node", "<u4"), ("offset", "<u4"), ("reserved", "<u4")])
EXT_RESULT = np.dtype([("status", "<i4"), ("score", "<i4"), ("trace_len", "<u4"), ("reserved", "<u4"), ("trace_offset", "<u8"), ("columns", "<u8")])

NW_ITEM = np.dtype([("query_offset", "<u8"), ("target_offset", "<u8"), ("query_len", "<i4"), ("target_len", "<i4"), ("k_hint", "<i4"), ("want_path", "<i4")])
NW_RESULT = np.dtype([("status", "<i4"), ("distance", "<i4"), ("ops_len", "<u4"), ("reserved", "<u4"), ("ops_offset", "<u8"), ("blocks", "<u8")])
ANCHOR = np.dtype([("start_node", "<u4"), ("end_node", "<u4"), ("x", "<i4"), ("y", "<i4")])

_lib = None


def load() -> C.CDLL:
    """Load libgcgpu.so; fails loudly when it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (libgcgpu has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    lib.gcgpu_version.restype = C.c_int
    lib.gcgpu_last_error.restype = C.c_char_p
    lib.gcgpu_create.argtypes = [C.c_int, C.POINTER(GraphStruct), C.POINTER(ParamsStruct), C.POINTER(C.c_void_p)]
    lib.gcgpu_create.restype = C.c_int
    lib.gcgpu_destroy.argtypes = [C.c_void_p]
    lib.gcgpu_destroy.restype = None
    lib.gcgpu_extend.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
    lib.gcgpu_extend.restype = C.c_int
    lib.gcgpu_nw.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
    lib.gcgpu_nw.restype = C.c_int
    lib.gcgpu_chain.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.gcgpu_chain.restype = C.c_int
    lib.gcgpu_last_kernel_ms.argtypes = [C.c_void_p]
    lib.gcgpu_last_kernel_ms.restype = C.c_float
    lib.gcgpu_launch_count.argtypes = [C.c_void_p]
    lib.gcgpu_launch_count.restype = C.c_uint64
    _lib = lib
    return lib


class GcgpuError(RuntimeError):
    pass


def _ptr(a: np.ndarray) -> int:
    return a.ctypes.data


class Context:
    """One libgcgpu context (= one device) holding a replica of the graph index."""

    def __init__(self, index: dict, device: int = 0, bandwidth: int = 10):
        self.lib = load()
        self.index = index
        g = GraphStruct()
        self._keep = []

        def arr(name, dtype):
            a = np.ascontiguousarray(index[name], dtype=dtype)
            self._keep.append(a)
            return _ptr(a)

        n = len(index["nodeLength"])
        g.num_nodes = n
        g.node_length = arr("nodeLength", np.uint8)
        g.node_seq = arr("nodeSeq", np.uint64)
        g.in_start = arr("inStart", np.uint32)
        g.in_nbr = arr("inNbr", np.uint32)
        g.out_start = arr("outStart", np.uint32)
        g.out_nbr = arr("outNbr", np.uint32)
        g.component_number = arr("componentNumber", np.uint32)
        g.linearizable = arr("linearizable", np.uint8)
        if "compMap" in index:
            g.num_components = len(index["compStart"]) - 1
            g.comp_map = arr("compMap", np.uint32)
            g.comp_idx = arr("compIdx", np.uint32)
            g.comp_start = arr("compStart", np.uint32)
            g.topo_ids = arr("topoIds", np.uint32)
            g.paths_start = arr("pathsStart", np.uint32)
            g.paths_k = arr("pathsK", np.uint32)
            g.back_start = arr("backStart", np.uint32)
            g.back_node = arr("backNode", np.uint32)
            g.back_k = arr("backK", np.uint32)
        if "origIds" in index and "nodeIDs" in index:
            g.node_ids = arr("nodeIDs", np.int32)
            g.node_offset = arr("nodeOffset", np.uint32)
            g.num_orig = len(index["origIds"])
            g.orig_ids = arr("origIds", np.int32)
            g.orig_start = arr("origStart", np.uint32)
            g.orig_nodes = arr("origNodes", np.uint32)
            g.orig_size = arr("origSize", np.uint32)
        p = ParamsStruct(bandwidth)
        h = C.c_void_p()
        rc = self.lib.gcgpu_create(device, C.byref(g), C.byref(p), C.byref(h))
        if rc != 0:
            raise GcgpuError(f"gcgpu_create failed ({rc}): {self.lib.gcgpu_last_error().decode()}")
        self.handle = h

    def close(self):
        if getattr(self, "handle", None):
            self.lib.gcgpu_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- split-node lookup on the host (AlignmentGraph::GetUnitigNode, AlignmentGraph.cpp:832-848)
    def unitig_node(self, bigraph_node: int, offset: int):
        idx = self.index
        if not hasattr(self, "_orig"):
            self._orig = {int(i): k for k, i in enumerate(idx["origIds"])}
        k = self._orig[bigraph_node]
        nodes = idx["origNodes"][idx["origStart"][k]:idx["origStart"][k + 1]]
        offs = idx["nodeOffset"][nodes]
        j = int(np.searchsorted(offs, offset, side="right")) - 1
        node = int(nodes[j])
        return node, offset - int(idx["nodeOffset"][node])

    def extend(self, seq_codes: np.ndarray, items: np.ndarray, allow_internal: bool = False):
        """K1: run a batch of extension work items; returns (results, traces)."""
        seq_codes = np.ascontiguousarray(seq_codes, dtype=np.uint8)
        items = np.ascontiguousarray(items, dtype=EXT_ITEM)
        n = len(items)
        results = np.zeros(n, dtype=EXT_RESULT)
        cap = int((2 * items["seq_len"].astype(np.int64) + 72).sum()) + 1
        traces = np.zeros(cap, dtype=np.uint64)
        used = C.c_uint64(0)
        rc = self.lib.gcgpu_extend(self.handle, _ptr(seq_codes), seq_codes.size, _ptr(items), n, _ptr(results), _ptr(traces), cap, C.byref(used))
        if rc != 0 and not (allow_internal and rc == -4):
            raise GcgpuError(f"gcgpu_extend failed ({rc}): {self.lib.gcgpu_last_error().decode()}")
        return results, traces[:used.value]

    def nw(self, seqs: bytes, items: np.ndarray):
        """K3: batch of edlib-style NW alignments over raw characters; returns (results, ops)."""
        buf = np.frombuffer(seqs, dtype=np.uint8) if isinstance(seqs, (bytes, bytearray)) else np.ascontiguousarray(seqs, dtype=np.uint8)
        items = np.ascontiguousarray(items, dtype=NW_ITEM)
        n = len(items)
        results = np.zeros(n, dtype=NW_RESULT)
        want = items["want_path"] != 0
        cap = int((items["query_len"].astype(np.int64) + items["target_len"] + 8)[want].sum()) + 1
        ops = np.zeros(cap, dtype=np.uint8)
        used = C.c_uint64(0)
        rc = self.lib.gcgpu_nw(self.handle, _ptr(buf), buf.size, _ptr(items), n, _ptr(results), _ptr(ops), cap, C.byref(used))
        if rc != 0:
            raise GcgpuError(f"gcgpu_nw failed ({rc}): {self.lib.gcgpu_last_error().decode()}")
        return results, ops[:used.value]

    def chain(self, anchors: np.ndarray, read_offsets: np.ndarray):
        """K2: co-linear chaining per read; returns (chain, chain_len, chain_score)."""
        anchors = np.ascontiguousarray(anchors, dtype=ANCHOR)
        read_offsets = np.ascontiguousarray(read_offsets, dtype=np.uint64)
        nreads = len(read_offsets) - 1
        chain = np.zeros(max(1, len(anchors)), dtype=np.uint32)
        chain_len = np.zeros(max(1, nreads), dtype=np.uint32)
        chain_score = np.zeros(max(1, nreads), dtype=np.int64)
        rc = self.lib.gcgpu_chain(self.handle, _ptr(anchors), _ptr(read_offsets), nreads, _ptr(chain), _ptr(chain_len), _ptr(chain_score))
        if rc != 0:
            raise GcgpuError(f"gcgpu_chain failed ({rc}): {self.lib.gcgpu_last_error().decode()}")
        return chain, chain_len[:nreads], chain_score[:nreads]

    @property
    def last_kernel_ms(self) -> float:
        return float(self.lib.gcgpu_last_kernel_ms(self.handle))

    @property
    def launch_count(self) -> int:
        return int(self.lib.gcgpu_launch_count(self.handle))


def unpack_trace(t: np.ndarray):
    """Packed trace entries -> (node, offset, seqPos, nodeSwitch) arrays."""
    t = t.astype(np.uint64)
    node = (t & np.uint64(0xFFFFFFFF)).astype(np.int64)
    off = ((t >> np.uint64(32)) & np.uint64(63)).astype(np.int64)
    sw = ((t >> np.uint64(38)) & np.uint64(1)).astype(np.int64)
    sp = ((t >> np.uint64(39)) & np.uint64(0x1FFFFFF)).astype(np.int64) - 1
    return node, off, sp, sw


def pack_trace(node, off, sp, sw) -> np.ndarray:
    return (np.asarray(node, dtype=np.uint64) | (np.asarray(off, dtype=np.uint64) << np.uint64(32))
            | (np.asarray(sw, dtype=np.uint64) << np.uint64(38)) | ((np.asarray(sp, dtype=np.int64) + 1).astype(np.uint64) << np.uint64(39)))
