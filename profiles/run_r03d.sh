#!/bin/bash
# r03d: lane-per-item K1 (gc_k1s) vs the lock-step kernels: parity tests, per-launch trace of one 839-read batch in both modes, bench in both modes
cd "$GRAFT_REPO_ROOT"
O=gpurun_out; mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/r03d_tests.log 2>&1
tail -3 $O/r03d_tests.log
python -c "
from graphchainer_b200 import synth
print(synth.make_workload('c2', '/tmp/c2s', n_reads=839))
"
D=graphchainer_b200/GraphChainerB200
$D -g /tmp/c2s.gfa --gc-save-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o.gam -t 16 --gc-streams 1 > /dev/null 2>&1
GCGPU_TRACE=1 GC_TRACE=1 $D --gc-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o.gam -t 16 --gc-streams 1 > $O/r03d_trace_simt.txt 2>&1
GCGPU_K1_LOCKSTEP=1 GCGPU_TRACE=1 GC_TRACE=1 $D --gc-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o2.gam -t 16 --gc-streams 1 > $O/r03d_trace_lockstep.txt 2>&1
cmp /tmp/o.gam /tmp/o2.gam && echo "GAM identical between modes"
grep "k1 (long" $O/r03d_trace_simt.txt $O/r03d_trace_lockstep.txt
timeout 900 python bench.py --no-cpu-baseline > $O/r03d_bench_simt.json 2> $O/r03d_bench_simt.err
GCGPU_K1_LOCKSTEP=1 timeout 900 python bench.py --no-cpu-baseline > $O/r03d_bench_lockstep.json 2> $O/r03d_bench_lockstep.err
python - <<'PY'
import json
for m in ("simt","lockstep"):
    try:
        d=json.load(open(f"gpurun_out/r03d_bench_{m}.json"))
        print(m, "value %.1f Mbp/s e2e %.1f Mbp/s" % (d["value"]/1e6, d["e2e"]["value"]/1e6), d["kernels_ms_per_step"])
    except Exception as e: print(m, "failed", e)
PY
