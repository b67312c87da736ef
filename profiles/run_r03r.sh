#!/bin/bash
# r03r: K1 launches split between the two forms (longest items lock-step on the side stream, the rest lane-per-item); parallel GAM gather
cd "$GRAFT_REPO_ROOT"
O=gpurun_out; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/r03r_tests.log 2>&1
tail -3 $O/r03r_tests.log
python -c "
from graphchainer_b200 import synth
print(synth.make_workload('c2', '/tmp/c2s', n_reads=1678))
"
D=graphchainer_b200/GraphChainerB200
$D -g /tmp/c2s.gfa --gc-save-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o.gam -t 16 --gc-streams 1 > /dev/null 2>&1
for W in 0 0.5 1 2; do
GCGPU_K1_SPLIT=$W GCGPU_TRACE=1 GC_TRACE=1 $D --gc-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o_$W.gam -t 16 --gc-streams 1 > $O/r03r_trace_split$W.txt 2>&1
echo "== split weight $W"; grep "k1 (long\|k1 forms" $O/r03r_trace_split$W.txt | head -10
done
cmp /tmp/o_0.gam /tmp/o_1.gam && cmp /tmp/o_0.5.gam /tmp/o_1.gam && cmp /tmp/o_2.gam /tmp/o_1.gam && echo "GAM identical across splits"
for W in 1 0 0.5 2; do
GCGPU_K1_SPLIT=$W GC_TRACE_CALL=1 timeout 900 python bench.py --no-cpu-baseline > $O/r03r_bench_c2_split$W.json 2> $O/r03r_bench_c2_split$W.err
grep "gather" $O/r03r_bench_c2_split$W.err | tail -2
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r03r_bench*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "value %.1f Mbp/s e2e %.1f Mbp/s" % ((d["value"] or 0)/1e6, (d["e2e"]["value"] or 0)/1e6), {k: round(v,1) for k,v in d["kernels_ms_per_step"].items()}, d.get("parity_on_sample"), d["roofline"]["frac"])
    except Exception as e: print(f, "failed", e)
PY
