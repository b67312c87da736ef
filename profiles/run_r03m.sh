#!/bin/bash
# r03m: K1 lane-per-item kernels as persistent warps (lanes take item after item), unrolled column loop, ordered slice lookups
cd "$GRAFT_REPO_ROOT"
O=gpurun_out; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/r03m_tests.log 2>&1
tail -3 $O/r03m_tests.log
python -c "
from graphchainer_b200 import synth
print(synth.make_workload('c2', '/tmp/c2s', n_reads=1678))
"
D=graphchainer_b200/GraphChainerB200
$D -g /tmp/c2s.gfa --gc-save-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o.gam -t 16 --gc-streams 1 > /dev/null 2>&1
for L in 8 32 4 16; do
GCGPU_K1_LANES=$L GCGPU_K1_FORM=lane GCGPU_TRACE=1 GC_TRACE=1 $D --gc-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o_$L.gam -t 16 --gc-streams 1 > $O/r03m_trace_lanes$L.txt 2>&1
echo "== lanes $L (all whole-read launches in lane form)"; grep "gcgpu\] k1 (long" $O/r03m_trace_lanes$L.txt | tail -6
done
GCGPU_K1_FORM=lockstep GCGPU_TRACE=1 GC_TRACE=1 $D --gc-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o_ls.gam -t 16 --gc-streams 1 > $O/r03m_trace_lockstep.txt 2>&1
echo "== lockstep"; grep "gcgpu\] k1 (long" $O/r03m_trace_lockstep.txt | tail -6
GCGPU_TRACE=1 GC_TRACE=1 $D --gc-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o_def.gam -t 16 --gc-streams 1 > $O/r03m_trace_default.txt 2>&1
echo "== default"; grep "gcgpu\] k1 (long\|B200:" $O/r03m_trace_default.txt | tail -7
cmp /tmp/o_8.gam /tmp/o_ls.gam && cmp /tmp/o_32.gam /tmp/o_ls.gam && cmp /tmp/o_4.gam /tmp/o_ls.gam && echo "GAM identical across forms"
timeout 1200 python bench.py > $O/r03m_bench_c2.json 2> $O/r03m_bench_c2.err
tail -3 $O/r03m_bench_c2.err
for L in 32 16; do
GCGPU_K1_LANES=$L timeout 900 python bench.py --no-cpu-baseline > $O/r03m_bench_c2_lanes$L.json 2> $O/r03m_bench_c2_lanes$L.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r03m_bench*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "value %.1f Mbp/s e2e %.1f Mbp/s" % ((d["value"] or 0)/1e6, (d["e2e"]["value"] or 0)/1e6), {k: round(v,1) for k,v in d["kernels_ms_per_step"].items()}, d.get("parity_on_sample"), d["roofline"]["frac"])
    except Exception as e: print(f, "failed", e)
PY
