#!/bin/bash
# r03e: hybrid K1 (lane-per-item for launches >= 4096 long items, lock-step below), block-parallel piece / token kernels; new bench (parity gate)
cd "$GRAFT_REPO_ROOT"
O=gpurun_out; mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/r03e_tests.log 2>&1
tail -3 $O/r03e_tests.log
python -c "
from graphchainer_b200 import synth
print(synth.make_workload('c2', '/tmp/c2s', n_reads=839))
"
D=graphchainer_b200/GraphChainerB200
$D -g /tmp/c2s.gfa --gc-save-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o.gam -t 16 --gc-streams 1 > /dev/null 2>&1
GCGPU_TRACE=1 GC_TRACE=1 $D --gc-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o.gam -t 16 --gc-streams 1 > $O/r03e_trace.txt 2>&1
grep "gcgpu\]\|phase" $O/r03e_trace.txt | grep -v hint
timeout 1200 python bench.py > $O/r03e_bench.json 2> $O/r03e_bench.err
tail -3 $O/r03e_bench.err
for S in 8 12; do
timeout 900 python bench.py --no-cpu-baseline --streams $S > $O/r03e_bench_s$S.json 2> $O/r03e_bench_s$S.err
done
GCGPU_K1_SIMT_MIN=1000000000 timeout 900 python bench.py --no-cpu-baseline > $O/r03e_bench_lockstep.json 2> $O/r03e_bench_lockstep.err
python - <<'PY'
import json
for m in ("","_s8","_s12","_lockstep"):
    try:
        d=json.load(open(f"gpurun_out/r03e_bench{m}.json"))
        print(m or "default", "value %.1f Mbp/s e2e %.1f Mbp/s" % ((d["value"] or 0)/1e6, (d["e2e"]["value"] or 0)/1e6), d["kernels_ms_per_step"], d.get("parity_on_sample"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(m, "failed", e)
PY
