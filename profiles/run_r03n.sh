#!/bin/bash
# r03n: K1 lane form with 16-column windows, duplicate pops skipped, walk in 16-cell rounds; launch bounds (64,6)/(64,8)
cd "$GRAFT_REPO_ROOT"
O=gpurun_out; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q -k "k1 or pipeline or align" ) > $O/r03n_tests.log 2>&1
tail -3 $O/r03n_tests.log
python -c "
from graphchainer_b200 import synth
print(synth.make_workload('c2', '/tmp/c2s', n_reads=1678))
"
D=graphchainer_b200/GraphChainerB200
$D -g /tmp/c2s.gfa --gc-save-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o.gam -t 16 --gc-streams 1 > /dev/null 2>&1
for L in 8 32 4; do
GCGPU_K1_LANES=$L GCGPU_K1_FORM=lane GCGPU_TRACE=1 GC_TRACE=1 $D --gc-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o_$L.gam -t 16 --gc-streams 1 > $O/r03n_trace_lanes$L.txt 2>&1
echo "== lanes $L (all whole-read launches in lane form)"; grep "k1 (long" $O/r03n_trace_lanes$L.txt | tail -6
done
GCGPU_K1_FORM=lockstep GCGPU_TRACE=1 GC_TRACE=1 $D --gc-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o_ls.gam -t 16 --gc-streams 1 > $O/r03n_trace_lockstep.txt 2>&1
cmp /tmp/o_8.gam /tmp/o_ls.gam && cmp /tmp/o_32.gam /tmp/o_ls.gam && cmp /tmp/o_4.gam /tmp/o_ls.gam && echo "GAM identical across forms"
timeout 1200 python bench.py --no-cpu-baseline > $O/r03n_bench_c2.json 2> $O/r03n_bench_c2.err
tail -3 $O/r03n_bench_c2.err
GCGPU_K1_FORM=lane timeout 900 python bench.py --no-cpu-baseline > $O/r03n_bench_c2_lane.json 2> $O/r03n_bench_c2_lane.err
GCGPU_K1_FORM=lockstep timeout 900 python bench.py --no-cpu-baseline > $O/r03n_bench_c2_lockstep.json 2> $O/r03n_bench_c2_lockstep.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r03n_bench*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "value %.1f Mbp/s e2e %.1f Mbp/s" % ((d["value"] or 0)/1e6, (d["e2e"]["value"] or 0)/1e6), {k: round(v,1) for k,v in d["kernels_ms_per_step"].items()}, d.get("parity_on_sample"), d["roofline"]["frac"])
    except Exception as e: print(f, "failed", e)
PY
