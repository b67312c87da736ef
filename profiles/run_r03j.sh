#!/bin/bash
# r03j: GAM kernel one working lane per warp, K2 form per read; batch/stream sweep at 16 and 4 host threads
cd "$GRAFT_REPO_ROOT"
O=gpurun_out; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/r03j_tests.log 2>&1
tail -3 $O/r03j_tests.log
python -c "
from graphchainer_b200 import synth
print(synth.make_workload('c2', '/tmp/c2s', n_reads=1678))
"
D=graphchainer_b200/GraphChainerB200
$D -g /tmp/c2s.gfa --gc-save-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o.gam -t 16 --gc-streams 1 > /dev/null 2>&1
GCGPU_TRACE=1 GC_TRACE=1 $D --gc-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o.gam -t 16 --gc-streams 1 > $O/r03j_trace.txt 2>&1
grep "gcgpu\] s7\|gcgpu\] k2\|phase" $O/r03j_trace.txt
timeout 1200 python bench.py > $O/r03j_bench.json 2> $O/r03j_bench.err
tail -3 $O/r03j_bench.err
for cfg in "4 25000000 16" "3 34000000 16" "2 50000000 16" "8 12600000 16" "6 0 4" "4 25000000 4" "3 34000000 4"; do
set -- $cfg
timeout 900 python bench.py --no-cpu-baseline --streams $1 --batch-bp $2 --host-threads $3 > $O/r03j_bench_s$1_b$2_t$3.json 2> $O/r03j_bench_s$1_b$2_t$3.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r03j_bench*.json")):
    try:
        d=json.load(open(f))
        print(f.split("/")[-1], "value %.1f Mbp/s e2e %.1f Mbp/s" % ((d["value"] or 0)/1e6, (d["e2e"]["value"] or 0)/1e6), {k: round(v,1) for k,v in d["kernels_ms_per_step"].items()}, d.get("parity_on_sample"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(f, "failed", e)
PY
