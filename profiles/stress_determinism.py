"""Self-consistency of the GAM output across repeated steps of the same reads: every step must give the same records (the parity
gate of bench.py compares one step with the reference; this looks for run-to-run differences).
    python profiles/stress_determinism.py <steps> [workload]
r04h (a build with a streaming submit / wait API, not kept): ~900 steps x 10000 reads of c2, calls one at a time and streamed, no
record ever changed -- the one parity difference of r04g was the reference's own multi-threaded nondeterminism
(profiles/r04h_reference_nondeterminism.txt)."""
import sys, os, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from graphchainer_b200 import align, gam as gamlib

mode, steps, workload = "sync", int(sys.argv[1]), (sys.argv[2] if len(sys.argv) > 2 else "c2")
from graphchainer_b200 import synth
n_reads = synth.WORKLOADS[workload]["n_reads"]
os.makedirs("/tmp/stress", exist_ok=True)
gfa, reads = bench.make_inputs(workload, n_reads, 0, "/tmp/stress")
batch = align.ReadBatch([r[0] for r in reads], [r[1] for r in reads])
a = align.Aligner(gfa, device=0, host_threads=16, split_len=35, split_gap=35, streams=6)
first_records = {}
step_no = [0]
def members(g, summ):
    """per-read digests of a step; a read whose record differs from step 0 is decoded and written out with both versions"""
    out = []
    for i in range(batch.n):
        rec = bytes(g[int(summ["gam_offset"][i]):int(summ["gam_offset"][i]) + int(summ["gam_size"][i])])
        d = hashlib.md5(rec).hexdigest()
        out.append(d)
        if step_no[0] == 0:
            first_records[i] = (d, rec, int(summ["used_chain"][i]), int(summ["num_alignments"][i]))
        elif d != first_records[i][0]:
            a0 = gamlib.read_gam_messages(first_records[i][1]); a1 = gamlib.read_gam_messages(rec)
            n, diffs = gamlib.diff_messages(a1, a0, names=[reads[i][0]], limit=3)
            print("DIFF step", step_no[0], "read", i, reads[i][0], "len", len(reads[i][1]), "used_chain", int(summ["used_chain"][i]), "was", first_records[i][2], "alignments", int(summ["num_alignments"][i]), "was", first_records[i][3],
                  "record bytes", len(rec), "was", len(first_records[i][1]), "|", str(diffs)[:1500], flush=True)
    step_no[0] += 1
    return out
outs = []
for _ in range(steps):
    g, summ, st = a.align(batch, gam=True)
    outs.append(members(g, summ))
a.close()
bad = {}
for s in range(1, len(outs)):
    for i in range(batch.n):
        if outs[s][i] != outs[0][i]:
            bad.setdefault(i, []).append(s)
print(mode, workload, "steps", steps, "reads", batch.n, "reads whose record differs from step 0 in some step:", len(bad), dict(list(bad.items())[:10]))
