"""Accuracy report next to the speed report: the per-read table of the reference's evaluation script
(/root/reference scripts/summary.py:87-201) for one or two GAM files --

    name, length, {long,clcs}_pathcnt, {long,clcs}_path_bps, {long,clcs}_revcnt, long_align_rate,
    global_ed_read_long, global_ed_read_clcs

pathcnt = mappings of the read's (last) alignment, path_bps = bases of the ORIGINAL nodes on its path, revcnt = mappings on the
reverse strand, align_rate = path_bps / read length, global_ed = edit distance (global, unit costs) between the read and the
concatenated node sequences of the path (summary.py:79-90: whole nodes, reverse-complemented where the mapping is reversed).
The reference script takes the distance from the edlib Python package, which is not in this image: `edit_distance` below is the
same quantity by Myers' bit-vector recurrence on Python integers (one |read|-bit word per path base).

    python -m graphchainer_b200.summary graph.gfa reads.fa out_long.gam [out_clc.gam] > summary.csv
"""
from __future__ import annotations

import sys

from . import gam as gamlib

_RC = {"A": "T", "C": "G", "G": "C", "T": "A", "a": "t", "c": "g", "g": "c", "t": "a", "N": "N", "n": "n"}


def revc(s: str) -> str:
    return "".join(_RC[c] for c in reversed(s))


def edit_distance(a: str, b: str) -> int:
    """Global edit distance (insertions, deletions, substitutions at cost 1): Myers 1999 with `a` down the rows of one big-integer
    word; the score of the last row is carried across the columns (horizontal delta into row 0 is +1, as for edlib's NW mode)."""
    m = len(a)
    if m == 0:
        return len(b)
    mask = (1 << m) - 1
    top = 1 << (m - 1)
    peq = {}
    for i, c in enumerate(a):
        peq[c] = peq.get(c, 0) | (1 << i)
    vp, vn, score = mask, 0, m
    for c in b:
        eq = peq.get(c, 0)
        xv = eq | vn
        xh = (((eq & vp) + vp) ^ vp) | eq
        ph = vn | ~(xh | vp)
        mh = vp & xh
        if ph & top:
            score += 1
        elif mh & top:
            score -= 1
        ph = ((ph << 1) | 1) & mask
        mh = (mh << 1) & mask
        vp = (mh | ~(xv | ph)) & mask
        vn = ph & xv
    return score


def load_gfa(path: str) -> dict:
    """segment name -> sequence (summary.py:19-33 reads integer names; any name works here)"""
    seqs = {}
    with open(path) as f:
        for line in f:
            if line.startswith("S"):
                p = line.split()
                seqs[p[1]] = p[2]
    return seqs


def read_sequences(path: str) -> dict:
    """read name (first word of the header) -> (sequence, header) of a FASTA or FASTQ file, plain or gzipped"""
    import gzip
    op = gzip.open if path.endswith(".gz") else open
    out = {}
    with op(path, "rt") as f:
        lines = [l.rstrip("\n") for l in f]
    i = 0
    while i < len(lines):
        l = lines[i]
        if l.startswith(">"):
            j = i + 1
            seq = []
            while j < len(lines) and not lines[j].startswith(">"):
                seq.append(lines[j])
                j += 1
            out[l[1:].split()[0]] = ("".join(seq), l[1:])
            i = j
        elif l.startswith("@") and i + 3 < len(lines) + 1:
            out[l[1:].split()[0]] = (lines[i + 1], l[1:])
            i += 4
        else:
            i += 1
    return out


def path_of(alignment: dict, segments: dict, by_id: dict) -> dict:
    """summary.py:79-90 for one decoded vg.Alignment (gam.decode_alignment)"""
    seq = []
    rev = 0
    for m in alignment["mappings"]:
        s = segments.get(m.get("name", "")) if m.get("name") else None
        if s is None:
            s = by_id[int(m.get("node_id", 0))]
        if m.get("is_reverse"):
            rev += 1
            s = revc(s)
        seq.append(s)
    seq = "".join(seq)
    return dict(seq=seq, path_cnt=len(alignment["mappings"]), revcnt=rev, path_bps=len(seq))


def table(gfa: str, reads: str, gam_long: str, gam_clc: str | None = None, distances: bool = True):
    """rows of the reference's summary CSV (header row first)"""
    segments = load_gfa(gfa)
    by_id = {i: s for i, s in enumerate(segments.values())}   # GFA segment ids = order of first appearance (GfaGraph.cpp:164-174)
    rd = read_sequences(reads)
    files = [gamlib.read_gam(gam_long), gamlib.read_gam(gam_clc) if gam_clc else {}]
    rows = [["name", "length", "long_pathcnt", "long_path_bps", "long_revcnt", "clcs_pathcnt", "clcs_path_bps", "clcs_revcnt", "long_align_rate", "global_ed_read_long", "global_ed_read_clcs"]]
    for name, (seq, _header) in rd.items():
        row = [name, str(len(seq))] + [""] * 9
        long_bps = 0
        for k, alns in enumerate(files):
            if name not in alns or not alns[name]:
                continue
            p = path_of(alns[name][-1], segments, by_id)   # parse_gam keeps the last alignment of a name (summary.py:92-97)
            row[2 + 3 * k:5 + 3 * k] = [str(p["path_cnt"]), str(p["path_bps"]), str(p["revcnt"])]
            if distances:
                row[9 + k] = str(edit_distance(seq, p["seq"]))
            if k == 0:
                long_bps = p["path_bps"]
        row[8] = str(long_bps / len(seq)) if seq else "0"
        rows.append(row)
    return rows


def accuracy(rows) -> dict:
    """the numbers the GraphChainer paper reports per data set: reads aligned, mean align rate, mean distance per read base"""
    body = rows[1:]
    aligned = [r for r in body if r[2]]
    out = dict(reads=len(body), aligned=len(aligned))
    if aligned:
        out["mean_align_rate"] = sum(float(r[8]) for r in aligned) / len(aligned)
        with_d = [r for r in aligned if r[9]]
        if with_d:
            out["edit_distance_per_read_base"] = sum(int(r[9]) for r in with_d) / sum(int(r[1]) for r in with_d)
    return out


if __name__ == "__main__":
    if len(sys.argv) < 4:
        sys.exit(__doc__)
    rows = table(sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
    for r in rows:
        print(",".join(r))
    print(accuracy(rows), file=sys.stderr)
