"""Parser for the stage records written by oracle/_ref/gc_refdump (oracle/ref_dump.cpp)."""
from __future__ import annotations

import numpy as np


def _trace(tokens):
    n = int(tokens[0])
    node = np.empty(n, dtype=np.int64); off = np.empty(n, dtype=np.int64); sp = np.empty(n, dtype=np.int64); sw = np.empty(n, dtype=np.int64)
    for i, tok in enumerate(tokens[1:1 + n]):
        a, b, c, d, _ = tok.split(",", 4)
        node[i] = int(a); off[i] = int(b); sp[i] = int(c); sw[i] = int(d)
    return node, off, sp, sw


def parse(path: str):
    """Return a list of reads; each read is a dict of its stage records."""
    reads = []
    cur = None
    pending = None
    with open(path) as f:
        for line in f:
            parts = line.rstrip("\n").split(" ")
            tag = parts[0]
            if tag == "READ":
                cur = dict(name=parts[1], seq=parts[2], ext=[], seeds={}, anchors=[], chainpaths=[], ga_all=[], ga_selected=[], clc=[], final=[])
                reads.append(cur)
                seedlist = None
            elif tag in ("SEEDS_RAW", "SEEDS_ORDERED", "SEEDS_BYPOS"):
                seedlist = []
                cur["seeds"][tag] = seedlist
            elif tag == "S":
                seedlist.append(tuple(int(x) for x in parts[1:]))
            elif tag == "EXT":
                pending = dict(stage=parts[1], frag=int(parts[2]), seed=int(parts[3]), dir=parts[4], node=int(parts[5]), offset=int(parts[6]), seq="" if parts[7] == "-" else parts[7])
            elif tag == "RES":
                if parts[1] == "F":
                    pending["failed"] = True
                else:
                    pending["failed"] = False
                    pending["score"] = int(parts[1])
                    pending["trace"] = _trace(parts[2:])
                cur["ext"].append(pending)
                pending = None
            elif tag in ("GA_ALL", "GA_SELECTED", "CLC"):
                alist = cur[tag.lower()]
                alist.clear()
                curlist = alist
            elif tag == "A":
                curlist.append(dict(start=int(parts[1]), end=int(parts[2]), score=int(parts[3]), goodness=int(parts[4]), trace_score=int(parts[5]), trace=_trace(parts[6:])))
            elif tag == "GA_PATHSEQ":
                cur["ga_pathseq"] = parts[1]
                cur["long_edit_distance"] = int(parts[2])
            elif tag == "AN":
                v = [int(x) for x in parts[1:]]
                cur["anchors"].append(dict(x=v[0], y=v[1], first=(v[2], v[3]), last=(v[4], v[5]), path=v[7:7 + v[6]]))
            elif tag == "CHAIN":
                cur["chain"] = [int(x) for x in parts[2:2 + int(parts[1])]]
            elif tag == "CHAINPATH":
                n = int(parts[4])
                cur["chainpaths"].append(dict(s=int(parts[1]), t=int(parts[2]), limit=int(parts[3]), path=[int(x) for x in parts[5:5 + n]]))
            elif tag == "LONGEST":
                cur["longest"] = [tuple(int(x) for x in tok.split(",")) for tok in parts[2:]]
            elif tag == "PATHSEQ":
                cur["pathseq"] = "" if parts[1] == "-" else parts[1]
            elif tag == "EDLIB":
                if parts[1] == "ERR":
                    cur["edlib"] = None
                else:
                    cur["edlib"] = dict(distance=int(parts[1]), length=int(parts[2]), start=int(parts[3]), end=int(parts[4]), ops=parts[5] if len(parts) > 5 else "")
            elif tag == "DECISION":
                cur["decision"] = parts[1]
                cur["clc_score"] = int(parts[2])
            elif tag == "FINAL":
                cur["final_status"] = parts[1]
            elif tag == "J":
                cur["final"].append(line[2:].rstrip("\n"))
    return reads
