"""bench.py -- aligned read bp/s of the GraphChainer per-read hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2] [--reads R]

One step = one pass of the whole hot path (seeding -> whole-read extension -> fragment anchoring
-> chaining -> path connection -> NW) over the workload's reads on one GPU.  N > 1 is launched by
torchrun (one rank per GPU): the graph index is replicated, every rank aligns its own equally
sized read set (weak scaling, no collective on the data path), the job value is total bp / max
time over ranks.
  value  : bp / (sum of the CUDA-event durations of all kernels of the step) -- inputs of every
           kernel are resident in HBM when its event pair starts.
  e2e    : bp / wall time of the reference-facing call gcalign_align() on HOST buffers (reads in,
           GAM records out; every host<->device copy and the host stages are inside).
--impl reference times the unmodified reference (oracle/_ref/GraphChainer_ref, all host cores)
on a bounded sample of the same reads.
"""
from __future__ import annotations

import argparse
import json
import os
import re

# the in-flight batches share the host cores: idle OpenMP teams (a batch waiting for its kernels) must sleep, not spin
os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
REFBIN = os.path.join(ROOT, "oracle", "_ref", "GraphChainer_ref")

# algorithmic bytes per work unit (DESIGN.md "Kernels"): one unit W = one 64-row Myers column step
K1_BYTES_PER_W = 2.9   # 184 B per node calculation (16 B sequence + 2 x 72 B slice items + 24 B per incoming edge) / ~64 columns
K3_BYTES_PER_W = 8.0   # one 8-byte Peq word per block step; block state is register/L1 resident by design
K1_OPS_PER_W = 44      # int32-equivalent ALU ops of getNextSlice (BVCommon.h:248-260)
K3_OPS_PER_W = 40      # calculateBlock (edlib.cpp:409-444)


def make_inputs(workload: str, n_reads: int, rank: int, tmp: str):
    from graphchainer_b200 import synth
    cfg = dict(synth.WORKLOADS[workload])
    g = synth.SynthGraph(cfg["graph_len"], seed=1, extra_alleles=cfg.get("extra_alleles", 0))
    gfa = os.path.join(tmp, f"{workload}.gfa")
    with open(gfa, "w") as f:
        f.write(g.gfa())
    reads = list(synth.simulate_reads(g, n_reads, cfg["read_len"], cfg["error"], seed=2 + rank))
    return gfa, reads


class ClockSampler(threading.Thread):
    def __init__(self, device: int):
        super().__init__(daemon=True)
        self.device = device
        self.samples = []
        self.reasons = set()
        self.stop_flag = False
        self.max_mhz = None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.device)], capture_output=True, text=True, timeout=5).stdout.strip()
                parts = [p.strip() for p in out.split(",")]
                self.samples.append(float(parts[0]))
                self.max_mhz = float(parts[1])
                for n, v in zip(names, parts[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def time_reference(gfa: str, fasta: str, threads: int, keep_gam: str | None = None, extra=()):
    """Align-phase seconds of the unmodified reference: wall clock between its "Align" and
    "Alignment finished" lines (src/Aligner.cpp:1258,1296); index build excluded.  keep_gam: where to leave its GAM output."""
    with tempfile.TemporaryDirectory() as d:
        out = keep_gam or os.path.join(d, "ref.gam")
        p = subprocess.Popen([REFBIN, "-t", str(threads), "-g", gfa, "-f", fasta, "-a", out, *extra], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        t0 = t1 = None
        for line in p.stdout:
            if line.startswith("Align") and not line.startswith("Alignment") and t0 is None:
                t0 = time.perf_counter()
            elif line.startswith("Alignment finished"):
                t1 = time.perf_counter()
        p.wait()
    if p.returncode != 0 or t0 is None or t1 is None:
        raise RuntimeError("reference run failed")
    return t1 - t0


def write_repeated_fasta(path: str, reads, reps: int) -> int:
    """The sample `reps` times in one file (names rep<r>_<name>): one reference process then aligns `reps` steps back to
    back, so its index build, thread start-up and queue sleeps (Aligner.cpp:179,214,252,504) are paid once, not per step."""
    total = 0
    with open(path, "w") as f:
        for r in range(reps):
            for name, seq in reads:
                f.write(f">rep{r}_{name}\n{seq}\n")
                total += len(seq)
    return total


def parity_on_sample(ref_gam_path: str, our_members: dict, names, gfa: str | None = None, reads_by_name: dict | None = None, extra=()):
    """Decoded-GAM equality of the reference's output and ours on the sample reads (no speed number
    counts before it).  our_members: {name: gzip member bytes of that read's record(s)}.
    The reference's multi-threaded runs are not run-to-run deterministic (c2, read_2220 / read_2329: 3 of 6 runs at -t 16 differ
    from the others and from -t 1 in `identity` and a few edits, profiles/r04h_reference_nondeterminism.txt; our records did not
    change in 8 M read-steps).  Reads that differ are therefore aligned again by the reference with ONE thread -- its canonical
    output, the one the golden fixtures hold -- and only what differs from that counts."""
    from graphchainer_b200 import gam as gamlib
    with open(ref_gam_path, "rb") as f:
        ref = gamlib.read_gam_messages(f.read())
    ours = {}
    for name, member in our_members.items():
        if len(member):
            ours.update(gamlib.read_gam_messages(member))
    n, diffs = gamlib.diff_messages(ours, ref, names=names, limit=1000)
    aligned = sum(1 for x in names if x in ref)
    out = {"reads": n, "reads_with_alignment_in_reference": aligned, "diffs": len(diffs), "first": diffs[:3]}
    if diffs and gfa and reads_by_name and len(diffs) <= 200:
        again = sorted({d.split(":")[0].split("[")[0] for d in diffs})
        with tempfile.TemporaryDirectory() as d:
            fa, g1 = os.path.join(d, "again.fa"), os.path.join(d, "again.gam")
            with open(fa, "w") as f:
                for name in again:
                    f.write(f">{name}\n{reads_by_name[name]}\n")
            time_reference(gfa, fa, 1, keep_gam=g1, extra=extra)
            with open(g1, "rb") as f:
                ref1 = gamlib.read_gam_messages(f.read())
        _, diffs1 = gamlib.diff_messages(ours, ref1, names=again, limit=1000)
        out = {"reads": n, "reads_with_alignment_in_reference": aligned, "diffs": len(diffs1), "first": diffs1[:3],
               "rechecked": {"reads": again[:20], "differed_from_the_multi_threaded_run": len(diffs), "differ_from_the_single_threaded_run": len(diffs1),
                             "note": "the reference's multi-threaded output is not run-to-run deterministic; differing reads were aligned again with -t 1"}}
    return out


def accuracy_on_sample(gfa: str, sample_reads, our_members: dict):
    """The reference's evaluation table (scripts/summary.py; graphchainer_b200/summary.py) on the first reads of the timed output:
    reads aligned, path bases per read base, edit distance between a read and the node sequences of its path per read base."""
    from graphchainer_b200 import gam as gamlib, summary
    segments = summary.load_gfa(gfa)
    by_id = {i: s for i, s in enumerate(segments.values())}
    rows = [["name", "length", "long_pathcnt", "long_path_bps", "long_revcnt", "", "", "", "long_align_rate", "global_ed_read_long", ""]]
    for name, seq in sample_reads:
        row = [name, str(len(seq))] + [""] * 9
        member = our_members.get(name, b"")
        msgs = gamlib.read_gam_messages(member).get(name, []) if len(member) else []
        if msgs:
            p = summary.path_of(gamlib.decode_alignment(msgs[-1]), segments, by_id)
            row[2:5] = [str(p["path_cnt"]), str(p["path_bps"]), str(p["revcnt"])]
            row[8] = str(p["path_bps"] / max(1, len(seq)))
            row[9] = str(summary.edit_distance(seq, p["seq"]))
        rows.append(row)
    out = summary.accuracy(rows)
    out["note"] = "first 100 reads of the timed output; edit distance = read against the whole nodes of its path (scripts/summary.py:79-90)"
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, help="c2 (default at --gpus 1: the configuration the metric is quoted on), c3 (default at --gpus > 1: BASELINE.json configs[2]), c4, c5")
    ap.add_argument("--reads", type=int, default=None, help="reads per step and GPU (default: the workload's full read set)")
    ap.add_argument("--cpu-sample", type=int, default=4000, help="reads in the CPU baseline sample (per step of the reference arm)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--streams", type=int, default=6, help="read batches in flight per GPU in the e2e measurement")
    ap.add_argument("--batch-bp", type=int, default=0, help="read bases per internal GPU batch (0 = library default)")
    ap.add_argument("--value-batch-bp", type=int, default=0, help="read bases per GPU batch in the kernel-time (value) measurement (0 = the whole read set in one batch)")
    ap.add_argument("--host-threads", type=int, default=0, help="host threads of this rank (0 = all cores / ranks); 4 at --gpus 1 reproduces one rank's share of a 32-core 8-GPU box")
    ap.add_argument("--threads-per-stream", type=int, default=0, help="host threads per in-flight batch (0 = host threads / streams)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    from graphchainer_b200 import synth
    if args.workload is None:
        args.workload = "c2" if args.gpus <= 1 else "c3"
    cfg = dict(synth.WORKLOADS[args.workload])
    # c3's full read set (100 k x 15 kb = 1.5 Gbp per GPU) is a 15-s step: the default is a stated tenth of it on the full 51 Mbp graph
    n_reads = args.reads or (cfg["n_reads"] if args.workload != "c3" else cfg["n_reads"] // 10)
    host_cores = os.cpu_count() or 1
    split_gap = int(cfg.get("split_gap", 35))
    ref_extra = ("--colinear-split-gap", str(split_gap)) if split_gap != 35 else ()
    read_len = cfg["read_len"] if isinstance(cfg["read_len"], int) else (cfg["read_len"][0] + cfg["read_len"][1]) // 2
    args.cpu_sample = max(1, min(args.cpu_sample, n_reads, int(args.cpu_sample * 10_000 / read_len)))  # ~40 Mbp of reads per reference step
    config = {"workload": f"{args.workload}: synthetic {cfg['graph_len'] / 1e6:g} Mbp acyclic SNP/indel graph + {n_reads} simulated reads/GPU, length {cfg['read_len']}, {cfg['error'] * 100:g}% error (5% with a novel 400-bp insertion)" + ("" if n_reads == cfg["n_reads"] else f" [{n_reads} of the configuration's {cfg['n_reads']} reads]"),
              "reads_per_gpu": n_reads, "graph_bp": cfg["graph_len"], "colinear_split_gap": split_gap, "l2": "inputs larger than L2: slice/trace workspaces of a step exceed 126 MB",
              "value_timing": "sum of CUDA-event durations of the step's kernels, the read set as one GPU batch, one batch in flight", "e2e_timing": "wall time of gcalign_align() on host buffers incl. all copies and host stages"}
    tmp = tempfile.mkdtemp(prefix="gcbench_")

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        if not os.path.exists(REFBIN):
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/GraphChainer_ref not built"}))
            return
        sample = args.cpu_sample
        gfa, reads = make_inputs(args.workload, sample, 0, tmp)
        # one process aligns all timed steps back to back (and one before it the warm-up steps): the reference rebuilds its
        # index in every process (86 s for c3 at -t 8), and a 0.5-s align phase per process would mostly measure its start-up
        if args.warmup > 0:
            warm = os.path.join(tmp, "warm.fa")
            write_repeated_fasta(warm, reads, args.warmup)
            time_reference(gfa, warm, host_cores, extra=ref_extra)
        fa = os.path.join(tmp, "sample.fa")
        bp_total = write_repeated_fasta(fa, reads, args.steps)
        t_all = time_reference(gfa, fa, host_cores, extra=ref_extra)
        bp = bp_total // args.steps
        t = t_all / args.steps
        v = bp / t
        line = {"metric": "aligned read bp/sec", "value": v, "unit": "bp/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic", "impl": "reference", "config": config,
                "cpu_baseline": {"value": v, "unit": "bp/s", "cores": host_cores, "kind": "reference",
                                 "sample": f"first {sample} reads of the workload ({bp} bp) per step, {args.steps} steps back to back in one process, unmodified reference sources built with shim headers (oracle/Makefile), -t {host_cores}, align phase only"},
                "e2e": {"value": v, "unit": "bp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ our arm
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (libgcgpu has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl")
    from graphchainer_b200 import align
    gfa, reads = make_inputs(args.workload, n_reads, rank, tmp)
    batch = align.ReadBatch([r[0] for r in reads], [r[1] for r in reads])
    threads = args.host_threads or max(1, host_cores // world)
    t_index = time.perf_counter()
    aligner = align.Aligner(gfa, device=local_rank, host_threads=threads, split_len=35, split_gap=split_gap, streams=args.streams, batch_bp=args.batch_bp, threads_per_stream=args.threads_per_stream)
    index_s = time.perf_counter() - t_index

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- e2e: the reference-facing call on host buffers, `streams` read batches in flight
    for _ in range(args.warmup):
        aligner.align(batch, gam=True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    gam_bytes = 0
    launches = 0
    pcie_h2d = pcie_d2h = 0
    for _ in range(args.steps):
        gam, summ, st = aligner.align(batch, gam=True)
        gam_bytes = len(gam)
        launches += st["launches"]
        pcie_h2d, pcie_d2h = st["h2d_bytes"], st["d2h_bytes"]
    barrier()
    wall = time.perf_counter() - t0
    # the sample reads' GAM records of the last timed step (the buffer is reused by the next call): parity gate below
    sample_n = args.cpu_sample if (rank == 0 and not args.no_cpu_baseline) else 0
    our_members = {reads[i][0]: bytes(gam[int(summ["gam_offset"][i]):int(summ["gam_offset"][i]) + int(summ["gam_size"][i])]) for i in range(sample_n)}
    try:
        int_peak = aligner.int_peak()
    except Exception:
        int_peak = None
    aligner.close()
    # ---- kernel time: the same steps with ONE batch in flight, so that every CUDA-event pair brackets a
    # kernel that has the GPU to itself (with several streams the event durations of overlapping kernels add up)
    # ... and the whole read set as ONE batch: the kernels then see the launch sizes they are built for (a 16 Mbp batch gives the
    # lane-per-item K1 kernels a few hundred warps; its kernel-time sum measures launch latency, not throughput)
    value_batch_bp = args.value_batch_bp or (batch.total_bp + 1)
    aligner1 = align.Aligner(gfa, device=local_rank, host_threads=threads, split_len=35, split_gap=split_gap, streams=1, batch_bp=value_batch_bp)
    for _ in range(args.warmup):
        aligner1.align(batch, gam=False)
    barrier()
    steps = []
    for _ in range(args.steps):
        _, summ1, st = aligner1.align(batch, gam=False)
        steps.append(st)
    # the kernel-time configuration (one big batch: other launch sizes, lanes that work through several items) must give the same
    # alignments as the end-to-end one whose records went through the parity gate: every summary field of every read
    fields = ("num_alignments", "used_chain", "anchors", "chained", "path_bp", "clc_score", "long_edit_distance")
    value_mismatch = int(sum(int((summ1[f] != summ[f]).sum()) for f in fields))
    barrier()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    aligner1.close()
    kernel_ms = sum(s["s0_ms"] + s["k1_ms"] + s["k2_ms"] + s["k3_ms"] for s in steps)
    agg = torch.tensor([wall, kernel_ms / 1e3, float(batch.total_bp * args.steps)], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = agg.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = agg.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        wall_max, kern_max, bp_total = float(mx[0]), float(mx[1]), float(sm[2])
    else:
        wall_max, kern_max, bp_total = float(agg[0]), float(agg[1]), float(agg[2])
    if rank != 0:
        return
    k = args.steps
    k1_ms = sum(s["k1_ms"] for s in steps) / k
    k3_ms = sum(s["k3_ms"] for s in steps) / k
    k2_ms = sum(s["k2_ms"] for s in steps) / k
    s0_ms = sum(s["s0_ms"] for s in steps) / k
    k1_cols = sum(s["k1_columns"] for s in steps) / k
    k3_blocks = sum(s["k3_blocks"] for s in steps) / k
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    if k3_ms >= k1_ms:
        dom, dom_ms, dom_bytes, dom_units, dom_ops = "K3: gc_k3w_distance_kernel + gc_k3l_level_kernel (NW distances and edit paths), all launches of a step", k3_ms, k3_blocks * K3_BYTES_PER_W, k3_blocks, K3_OPS_PER_W
    else:
        dom, dom_ms, dom_bytes, dom_units, dom_ops = "K1: gc_k1s_forward_kernel + gc_k1s_backtrace_kernel (whole-read extensions) + gc_k1_kernel + gc_k1_bt_kernel (fragments), all launches of a step", k1_ms, k1_cols * K1_BYTES_PER_W, k1_cols, K1_OPS_PER_W
    achieved = dom_bytes / (dom_ms / 1e3) / 1e9 if dom_ms > 0 else 0.0
    int_ops = dom_units * dom_ops / (dom_ms / 1e3) if dom_ms > 0 else 0.0
    # DRAM traffic and pipe utilisation of the dominant kernel: from the committed ncu --set full capture (profiles/ncu_dominant_launch.json,
    # written by profiles/ncu_summary.py from the .ncu-rep of this code state), never typed in
    ncu_dom = None
    try:
        ncu_dom = json.load(open(os.path.join(ROOT, "profiles", "ncu_dominant_launch.json")))
    except Exception:
        pass
    value = bp_total / kern_max
    e2e_value = bp_total / wall_max
    line = {"metric": "aligned read bp/sec", "value": value, "unit": "bp/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": kern_max * 1e3 / k, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": config,
            "e2e": {"value": e2e_value, "unit": "bp/s", "ms_per_step": wall_max * 1e3 / k, "h2d_bytes_per_step": int(pcie_h2d), "d2h_bytes_per_step": int(pcie_d2h),
                    "pcie_bytes_per_bp": {"h2d": pcie_h2d / batch.total_bp, "d2h": pcie_d2h / batch.total_bp}, "gam_bytes_per_step": int(gam_bytes),
                    "e2e_over_value": e2e_value / value if value > 0 else None,
                    "note": "h2d/d2h = every byte libgcgpu copied across PCIe during one step on rank 0 (gcgpu_transfer_bytes); the caller's buffers (reads in, GAM out) are host memory; "
                            "e2e_over_value = end-to-end throughput / kernel-only throughput (1.0 = nothing but kernels on the critical path; > 1 = kernels of different batches overlap)"},
            "gpu_launches": int(launches),
            "value_config_summary_mismatches": value_mismatch,
            "roofline": {"bound": "int", "kernel": dom, "achieved": int_ops / 1e12, "peak": (int_peak or 0.0) / 1e12, "unit": "T int32-op/s", "frac": (int_ops / int_peak) if int_peak else None,
                         "traffic": (ncu_dom or {}).get("dram_bytes_per_launch"),
                         "peak_source": "measured live: gcgpu_int_peak (independent LOP3+IADD3 chains, best of 4; ncu of that kernel: profiles/)",
                         "note": "integer-pipe bound bit-parallel kernels (SURVEY 8d): achieved = work units x int32-equivalent ops per unit (44 for K1 = getNextSlice BVCommon.h:248-260, 40 for K3 = calculateBlock edlib.cpp:409-444) / kernel time; traffic = dram bytes of the dominant launch from the committed ncu capture",
                         "units_per_step": dom_units,
                         "hbm": {"achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "peak_source": peak_src, "note": "algorithmic bytes = work units x bytes/unit (DESIGN.md) / kernel time"},
                         "ncu_dominant_launch": ncu_dom},
            "kernels_ms_per_step": {"s0_seed": s0_ms, "k1_extend": k1_ms, "k2_chain": k2_ms, "k3_nw": k3_ms},
            "gcups": {"k1": 64 * k1_cols / (k1_ms / 1e3) / 1e9 if k1_ms > 0 else None, "k3": 64 * k3_blocks / (k3_ms / 1e3) / 1e9 if k3_ms > 0 else None,
                      "note": "64 DP cells per work unit W (one Myers column step on a 64-row word), SURVEY 8d"},
            "work_per_step": {"k1_column_steps": k1_cols, "k3_block_steps": k3_blocks, "k1_items": steps[0]["k1_items"], "k3_items": steps[0]["k3_items"], "s1_rounds": steps[0]["s1_rounds"]},
            "clocks": sampler.summary(), "index_build_s": index_s, "host_threads_per_rank": threads, "streams": args.streams, "batch_bp": args.batch_bp, "threads_per_stream": args.threads_per_stream}
    if not args.no_cpu_baseline and os.path.exists(REFBIN):
        sample = args.cpu_sample
        fa = os.path.join(tmp, "sample.fa")
        bp = synth.write_fasta(fa, reads[:sample])
        ref_gam = os.path.join(tmp, "ref_sample.gam")
        secs = time_reference(gfa, fa, host_cores, keep_gam=ref_gam, extra=ref_extra)
        line["cpu_baseline"] = {"value": bp / secs, "unit": "bp/s", "cores": host_cores, "kind": "reference",
                                "sample": f"first {sample} reads of rank 0's set ({bp} bp), unmodified reference sources built with shim headers (oracle/Makefile), -t {host_cores}, align phase only, {secs:.2f} s"}
        # parity gate: the timed workload's own reads, reference output vs the GAM records the timed e2e step produced
        par = parity_on_sample(ref_gam, our_members, [r[0] for r in reads[:sample]], gfa=gfa, reads_by_name={r[0]: r[1] for r in reads[:sample]}, extra=ref_extra)
        line["parity_on_sample"] = par
        if par["diffs"] != 0:
            line["value"] = None
            line["e2e"]["value"] = None
            line["rejected"] = "decoded GAM records of the sample differ from the reference's: no speed number is reported"
        line["accuracy_on_sample"] = accuracy_on_sample(gfa, reads[:100], our_members)
    if value_mismatch:
        line["value"] = None
        line["rejected"] = (line.get("rejected", "") + " the one-batch configuration of the kernel-time measurement gave different alignment summaries than the end-to-end one").strip()
    print(json.dumps(line))


def _shutdown():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    try:
        main()
    finally:
        _shutdown()
