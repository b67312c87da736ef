"""ctypes mirror of libgcalign (include/gcalign.h): the whole per-read pipeline for a batch of
reads in host memory -> GAM records + per-read summaries.  No fallback: needs libgcgpu + a GPU."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgcalign.so")


class Options(C.Structure):
    _fields_ = [("device", C.c_int32), ("host_threads", C.c_int32), ("initial_bandwidth", C.c_int32), ("streams", C.c_int32),
                ("colinear_gap", C.c_int64), ("colinear_split_len", C.c_int64), ("colinear_split_gap", C.c_int64), ("batch_bp", C.c_uint64), ("gzip_level", C.c_int32), ("threads_per_stream", C.c_int32), ("no_colinear_chaining", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("k1_ms", C.c_double), ("k2_ms", C.c_double), ("k3_ms", C.c_double),
                ("k1_items", C.c_uint64), ("k1_columns", C.c_uint64), ("k2_anchors", C.c_uint64), ("k3_items", C.c_uint64), ("k3_blocks", C.c_uint64),
                ("launches", C.c_uint64), ("s1_rounds", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("seeds_found", C.c_uint64), ("seeds_extended", C.c_uint64), ("s0_ms", C.c_double)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


SUMMARY = np.dtype([("num_alignments", "<u4"), ("used_chain", "<u4"), ("anchors", "<u4"), ("chained", "<u4"), ("path_bp", "<u8"),
                    ("clc_score", "<u8"), ("long_edit_distance", "<u8"), ("gam_offset", "<u8"), ("gam_size", "<u8")])

_lib = None


def load(path: str | None = None):
    """Load libgcalign.so (or an explicitly given build of it) and declare its entry points."""
    global _lib
    explicit = path is not None   # an explicitly given build (tests: the C-ABI test double) never becomes the default
    if _lib is None or explicit:
        path = path or LIB_PATH
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: run __graft_entry__.build() (no CPU fallback exists)")
        lib = C.CDLL(path)
        lib.gcalign_default_options.argtypes = [C.POINTER(Options)]
        lib.gcalign_last_error.restype = C.c_char_p
        lib.gcalign_open.argtypes = [C.c_char_p, C.POINTER(Options), C.POINTER(C.c_void_p)]
        lib.gcalign_close.argtypes = [C.c_void_p]
        lib.gcalign_int_peak.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        lib.gcalign_align.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_void_p, C.POINTER(Stats)]
        if explicit:
            return lib
        _lib = lib
    return _lib


class ReadBatch:
    """Reads packed into the flat host buffers gcalign_align takes."""

    def subset(self, indices):
        """The reads `indices` (a rank's shard) as a new batch."""
        seqs = [bytes(self.seq_buf[int(self.seq_off[i]):int(self.seq_off[i + 1])]).decode() for i in indices]
        names = [bytes(self.name_buf[int(self.name_off[i]):int(self.name_off[i + 1])]).decode() for i in indices]
        return ReadBatch(names, seqs)

    def lengths(self):
        return np.diff(self.seq_off.astype(np.int64))

    def __init__(self, names, seqs):
        self.n = len(seqs)
        self.seq_buf = np.frombuffer("".join(seqs).encode(), dtype=np.uint8).copy()
        self.seq_off = np.zeros(self.n + 1, dtype=np.uint64)
        np.cumsum([len(s) for s in seqs], out=self.seq_off[1:])
        self.name_buf = np.frombuffer("".join(names).encode(), dtype=np.uint8).copy()
        self.name_off = np.zeros(self.n + 1, dtype=np.uint64)
        np.cumsum([len(s) for s in names], out=self.name_off[1:])
        self.total_bp = int(self.seq_off[-1])

    @classmethod
    def from_fasta(cls, path, limit=None):
        names, seqs = [], []
        with open(path) as f:
            name, parts = None, []
            for line in f:
                if line.startswith(">"):
                    if name is not None:
                        names.append(name); seqs.append("".join(parts))
                        if limit and len(seqs) >= limit:
                            name = None
                            break
                    name, parts = line[1:].rstrip("\n"), []
                else:
                    parts.append(line.strip())
            if name is not None:
                names.append(name); seqs.append("".join(parts))
        return cls(names, seqs)


class Aligner:
    def __init__(self, graph_path: str, device: int = 0, host_threads: int = 0, split_len: int = 35, split_gap: int = 35, colinear_gap: int = 10000, batch_bp: int = 0, streams: int = 0, gzip_level: int = 0, threads_per_stream: int = 0,
                 lib_path: str | None = None, colinear_chaining: bool = True):
        self.lib = load(lib_path)
        o = Options()
        self.lib.gcalign_default_options(C.byref(o))
        o.device, o.host_threads, o.streams = device, host_threads, streams
        o.colinear_split_len, o.colinear_split_gap, o.colinear_gap, o.batch_bp = split_len, split_gap, colinear_gap, batch_bp
        o.gzip_level = gzip_level
        o.threads_per_stream = threads_per_stream
        o.no_colinear_chaining = 0 if colinear_chaining else 1
        h = C.c_void_p()
        rc = self.lib.gcalign_open(graph_path.encode(), C.byref(o), C.byref(h))
        if rc != 0:
            raise RuntimeError(f"gcalign_open failed ({rc}): {self.lib.gcalign_last_error().decode()}")
        self.handle = h

    def int_peak(self) -> float:
        """Measured int32 LOP3/IADD3 instruction rate of the device (thread-level ops/s), gcgpu_int_peak."""
        v = C.c_double(0)
        rc = self.lib.gcalign_int_peak(self.handle, C.byref(v))
        if rc != 0:
            raise RuntimeError(f"gcalign_int_peak failed ({rc}): {self.lib.gcalign_last_error().decode()}")
        return v.value

    def close(self):
        if getattr(self, "handle", None):
            self.lib.gcalign_close(self.handle)
            self.handle = None

    def align(self, batch: ReadBatch, gam: bool = True):
        """Returns (GAM bytes as a uint8 array view, summaries, stats dict).  The GAM buffer belongs to the
        aligner and is reused by the next call (copy it -- bytes(view) -- to keep it)."""
        cap = (2 * batch.total_bp + (1 << 20)) if gam else 0
        if gam and (getattr(self, "_gam_buf", None) is None or self._gam_buf.size < cap):
            self._gam_buf = np.empty(cap, dtype=np.uint8)
        out = self._gam_buf if gam else None
        cap = out.size if gam else 0
        used = C.c_uint64(0)
        summ = np.zeros(batch.n, dtype=SUMMARY)
        st = Stats()
        rc = self.lib.gcalign_align(self.handle, batch.seq_buf.ctypes.data, batch.seq_off.ctypes.data, batch.name_buf.ctypes.data, batch.name_off.ctypes.data, batch.n,
                                    out.ctypes.data if gam else None, cap, C.byref(used), summ.ctypes.data, C.byref(st))
        if rc != 0:
            raise RuntimeError(f"gcalign_align failed ({rc}): {self.lib.gcalign_last_error().decode()}")
        return (out[:used.value] if gam else np.empty(0, dtype=np.uint8)), summ, st.as_dict()
