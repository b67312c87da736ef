#!/bin/bash
# r03s: GAM records' gzip members made by the 32 lanes of a warp (chunk-parallel); K1 forms per launch again
cd "$GRAFT_REPO_ROOT"
O=gpurun_out; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/r03s_tests.log 2>&1
tail -3 $O/r03s_tests.log
python -c "
from graphchainer_b200 import synth
print(synth.make_workload('c2', '/tmp/c2s', n_reads=1678))
"
D=graphchainer_b200/GraphChainerB200
$D -g /tmp/c2s.gfa --gc-save-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o.gam -t 16 --gc-streams 1 > /dev/null 2>&1
GCGPU_TRACE=1 GC_TRACE=1 $D --gc-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o.gam -t 16 --gc-streams 1 > $O/r03s_trace.txt 2>&1
grep "k1 (long\|s7 gam\|phase final\|phase gam" $O/r03s_trace.txt | tail -12
ls -la /tmp/o.gam
timeout 1200 python bench.py > $O/r03s_bench_c2.json 2> $O/r03s_bench_c2.err
tail -3 $O/r03s_bench_c2.err
timeout 1200 python bench.py --no-cpu-baseline --host-threads 4 > $O/r03s_bench_c2_t4.json 2> $O/r03s_bench_c2_t4.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r03s_bench*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "value %.1f Mbp/s e2e %.1f Mbp/s" % ((d["value"] or 0)/1e6, (d["e2e"]["value"] or 0)/1e6), {k: round(v,1) for k,v in d["kernels_ms_per_step"].items()}, d.get("parity_on_sample"), d["roofline"]["frac"], "gam bytes/step", d["e2e"].get("gam_bytes_per_step"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(f, "failed", e)
PY
