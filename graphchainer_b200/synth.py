"""Synthetic acyclic variation graphs and simulated long reads (no network, fixed seeds).

Workloads follow BASELINE.json `configs` / SURVEY.md §8(d):

* graph: random ACGT backbone of length L; variant sites with exponential spacing
  (mean 50 bp, min 20 bp); 85 % SNP bubbles (two 1-bp alleles), 15 % indels (a 1-5 bp
  segment plus a bypass link).  Written as GFA 1.0 `S`/`L ... 0M` lines, forward
  links only, integer segment names (first-appearance order = id order, which is
  what the reference's GfaGraph numbering uses, GfaGraph.cpp:164-174).
* reads: a random haplotype walk through the bubbles, random start, 50 % reverse
  complemented, then the error process of the reference's read simulator
  (src/SimulateReads.cpp:13-42): per base delete w.p. d, else substitute w.p. s by a
  uniform base (may equal the original), then w.p. i/10 insert U[0,19] random bases;
  d = s = i = error/3.  A fraction of the reads carries a novel 400-bp insertion
  (the read class where the chained alignment beats the whole-read one).
"""
from __future__ import annotations

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.array([3, 2, 1, 0], dtype=np.uint8)


class SynthGraph:
    """Backbone + variant sites; `gfa()` renders it, `haplotype()` samples a walk."""

    def __init__(self, length: int, seed: int = 1, mean_spacing: float = 50.0, min_spacing: int = 20,
                 snp_frac: float = 0.85, extra_alleles: int = 0):
        rng = np.random.default_rng(seed)
        self.length = int(length)
        self.bb = rng.integers(0, 4, size=self.length, dtype=np.uint8)  # backbone incl. ref alleles / indel segments
        pos = []
        p = int(min_spacing + rng.exponential(mean_spacing))
        while p + 6 + min_spacing < self.length:
            pos.append(p)
            p += 6 + int(max(min_spacing, rng.exponential(mean_spacing)))
        self.site_pos = np.asarray(pos, dtype=np.int64)
        n = len(pos)
        self.site_is_snp = rng.random(n) < snp_frac
        self.site_len = np.where(self.site_is_snp, 1, rng.integers(1, 6, size=n)).astype(np.int64)
        # alt allele differs from the backbone base
        self.site_alt = ((self.bb[self.site_pos] + rng.integers(1, 4, size=n)) % 4).astype(np.uint8)
        # optional extra haplotype-specific alleles per SNP site (raises the MPC width, config 5)
        self.extra_alleles = int(extra_alleles)

    # ------------------------------------------------------------------ GFA
    def gfa(self) -> str:
        out = []
        links = []
        bb = _ACGT[self.bb].tobytes().decode()
        nid = 1
        prev_shared = None
        cur = 0
        n = len(self.site_pos)
        for i in range(n + 1):
            end = int(self.site_pos[i]) if i < n else self.length
            shared = nid
            nid += 1
            out.append(f"S\t{shared}\t{bb[cur:end]}")
            if prev_shared is not None:
                for a in pending_alleles:
                    links.append((a, shared))
                if pending_bypass:
                    links.append((prev_shared, shared))
            if i == n:
                break
            sp, sl = int(self.site_pos[i]), int(self.site_len[i])
            pending_alleles = []
            pending_bypass = False
            a = nid
            nid += 1
            out.append(f"S\t{a}\t{bb[sp:sp + sl]}")
            links.append((shared, a))
            pending_alleles.append(a)
            if self.site_is_snp[i]:
                b = nid
                nid += 1
                out.append(f"S\t{b}\t{'ACGT'[int(self.site_alt[i])]}")
                links.append((shared, b))
                pending_alleles.append(b)
                for k in range(self.extra_alleles):
                    c = nid
                    nid += 1
                    base = 'ACGT'[(int(self.site_alt[i]) + 1 + k) % 4]
                    out.append(f"S\t{c}\t{base}")
                    links.append((shared, c))
                    pending_alleles.append(c)
            else:
                pending_bypass = True
            prev_shared = shared
            cur = sp + sl
        out.extend(f"L\t{u}\t+\t{v}\t+\t0M" for u, v in links)
        return "\n".join(out) + "\n"

    # ------------------------------------------------------------- haplotype
    def haplotype(self, rng: np.random.Generator, start: int, length: int) -> np.ndarray:
        """A `length`-bp haplotype walk starting at backbone position `start` (codes 0..3)."""
        slack = int(length * 0.02) + 64
        end = min(self.length, start + length + slack)
        win = self.bb[start:end].copy()
        lo = np.searchsorted(self.site_pos, start, side="left")
        hi = np.searchsorted(self.site_pos, end - 6, side="left")
        if hi > lo:
            sp = self.site_pos[lo:hi] - start
            snp = self.site_is_snp[lo:hi]
            pick = rng.random(hi - lo) < 0.5
            s_idx = sp[snp & pick]
            win[s_idx] = self.site_alt[lo:hi][snp & pick]
            drop = (~snp) & pick
            if drop.any():
                keep = np.ones(len(win), dtype=bool)
                for p, l in zip(sp[drop], self.site_len[lo:hi][drop]):
                    keep[p:p + l] = False
                win = win[keep]
        return win[:length]


def introduce_errors(rng: np.random.Generator, seq: np.ndarray, error_rate: float) -> np.ndarray:
    """Vectorised restatement of introduceErrors (SimulateReads.cpp:13-42), rates = error/3 each."""
    s = i = d = error_rate / 3.0
    n = len(seq)
    deleted = rng.random(n) < d
    subst = (~deleted) & (rng.random(n) < s)
    base = seq.copy()
    base[subst] = rng.integers(0, 4, size=int(subst.sum()), dtype=np.uint8)
    ins = rng.random(n) < i / 10.0
    ins_len = np.where(ins, rng.integers(0, 20, size=n), 0)
    out_len = (~deleted).astype(np.int64) + ins_len
    total = int(out_len.sum())
    out = rng.integers(0, 4, size=total, dtype=np.uint8)  # inserted bases are uniform random
    ends = np.cumsum(out_len)
    starts = ends - out_len
    kept = ~deleted
    out[starts[kept]] = base[kept]
    return out


def simulate_reads(graph: SynthGraph, n_reads: int, read_len, error_rate: float, seed: int = 2,
                   novel_insertion_frac: float = 0.05, novel_insertion_len: int = 400):
    """Yield (name, sequence str).  `read_len` is an int or a (lo, hi) range."""
    rng = np.random.default_rng(seed)
    for r in range(n_reads):
        length = int(read_len if np.isscalar(read_len) else rng.integers(read_len[0], read_len[1] + 1))
        length = min(length, graph.length - 256)
        start = int(rng.integers(0, max(1, graph.length - length - int(length * 0.02) - 128)))
        hap = graph.haplotype(rng, start, length)
        if novel_insertion_frac > 0 and rng.random() < novel_insertion_frac and len(hap) > 3 * novel_insertion_len:
            mid = len(hap) // 2
            hap = np.concatenate([hap[:mid], rng.integers(0, 4, size=novel_insertion_len, dtype=np.uint8), hap[mid:]])
        if rng.random() < 0.5:
            hap = _COMP[hap[::-1]]
        seq = introduce_errors(rng, hap, error_rate)
        yield f"read_{r}", _ACGT[seq].tobytes().decode()


def write_fasta(path: str, reads) -> int:
    total = 0
    with open(path, "w") as f:
        for name, seq in reads:
            f.write(f">{name}\n{seq}\n")
            total += len(seq)
    return total


#: named workloads = BASELINE.json configs (C2..C5); C1 is the reference's own test/ fixture
WORKLOADS = {
    "tiny": dict(graph_len=20_000, n_reads=20, read_len=2_000, error=0.15),
    "small": dict(graph_len=200_000, n_reads=100, read_len=5_000, error=0.15),
    "c2": dict(graph_len=5_000_000, n_reads=10_000, read_len=10_000, error=0.15),
    "c3": dict(graph_len=51_000_000, n_reads=100_000, read_len=15_000, error=0.10),
    "c4": dict(graph_len=51_000_000, n_reads=2_000, read_len=(50_000, 100_000), error=0.12),
    # --sampling-step 0.5 of BASELINE config 5 = fragments every 18 bp (the reference cannot parse that flag, AlignerMain.cpp:43,233: --colinear-split-gap 18)
    "c5": dict(graph_len=5_000_000, n_reads=5_000, read_len=20_000, error=0.01, extra_alleles=2, split_gap=18),
}


def make_workload(name: str, out_prefix: str, n_reads: int | None = None, graph_seed: int = 1, read_seed: int = 2):
    """Write `<out_prefix>.gfa` and `<out_prefix>.fa`; return (gfa_path, fa_path, total_bp)."""
    cfg = dict(WORKLOADS[name])
    g = SynthGraph(cfg["graph_len"], seed=graph_seed, extra_alleles=cfg.get("extra_alleles", 0))
    gfa_path, fa_path = out_prefix + ".gfa", out_prefix + ".fa"
    with open(gfa_path, "w") as f:
        f.write(g.gfa())
    n = cfg["n_reads"] if n_reads is None else n_reads
    total = write_fasta(fa_path, simulate_reads(g, n, cfg["read_len"], cfg["error"], seed=read_seed))
    return gfa_path, fa_path, total


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("workload", choices=sorted(WORKLOADS))
    ap.add_argument("out_prefix")
    ap.add_argument("--reads", type=int, default=None)
    a = ap.parse_args()
    print(make_workload(a.workload, a.out_prefix, a.reads))
