#!/bin/bash
# r03h: K3 block form (wide bands), own path cover in the index builder: parity tests, c4 + c2 bench, trace of a c4 batch
cd "$GRAFT_REPO_ROOT"
O=gpurun_out; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/r03h_tests.log 2>&1
tail -3 $O/r03h_tests.log
timeout 1800 python bench.py --workload c4 --reads 600 --steps 2 > $O/r03h_bench_c4.json 2> $O/r03h_bench_c4.err
tail -2 $O/r03h_bench_c4.err
timeout 1200 python bench.py > $O/r03h_bench_c2.json 2> $O/r03h_bench_c2.err
tail -2 $O/r03h_bench_c2.err
python -c "
from graphchainer_b200 import synth
print(synth.make_workload('c4', '/tmp/c4s', n_reads=100))
"
D=graphchainer_b200/GraphChainerB200
$D -g /tmp/c4s.gfa --gc-save-index /tmp/c4s.gcidx -f /tmp/c4s.fa -a /tmp/o.gam -t 16 --gc-streams 1 > /dev/null 2>&1
GCGPU_TRACE=1 GC_TRACE=1 $D --gc-index /tmp/c4s.gcidx -f /tmp/c4s.fa -a /tmp/o.gam -t 16 --gc-streams 1 > $O/r03h_trace_c4_100reads.txt 2>&1
grep "gcgpu\]\|phase" $O/r03h_trace_c4_100reads.txt | grep -v hint | tail -40
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r03h_bench*.json")):
    try:
        d=json.load(open(f))
        print(f.split("/")[-1], "value %.1f Mbp/s e2e %.1f Mbp/s" % ((d["value"] or 0)/1e6, (d["e2e"]["value"] or 0)/1e6), {k: round(v,1) for k,v in d["kernels_ms_per_step"].items()}, d.get("parity_on_sample"), (d.get("cpu_baseline") or {}).get("value"), "index_s", round(d["index_build_s"],1))
    except Exception as e: print(f, "failed", e)
PY
