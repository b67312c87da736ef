#!/bin/bash
# r03o: K1 lane form: persistent lanes (32 per warp), duplicate pops skipped, full column loop; register budget A/B (GCGPU_K1_OCC)
cd "$GRAFT_REPO_ROOT"
O=gpurun_out; mkdir -p $O
python -c "
from graphchainer_b200 import synth
print(synth.make_workload('c2', '/tmp/c2s', n_reads=1678))
"
D=graphchainer_b200/GraphChainerB200
$D -g /tmp/c2s.gfa --gc-save-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o.gam -t 16 --gc-streams 1 > /dev/null 2>&1
for OCC in 0 1; do
GCGPU_K1_OCC=$OCC GCGPU_K1_FORM=lane GCGPU_TRACE=1 GC_TRACE=1 $D --gc-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o_$OCC.gam -t 16 --gc-streams 1 > $O/r03o_trace_occ$OCC.txt 2>&1
echo "== occ $OCC (all whole-read launches in lane form)"; grep "k1 (long" $O/r03o_trace_occ$OCC.txt | tail -6
done
GCGPU_K1_FORM=lockstep $D --gc-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o_ls.gam -t 16 --gc-streams 1 > /dev/null 2>&1
cmp /tmp/o_0.gam /tmp/o_ls.gam && cmp /tmp/o_1.gam /tmp/o_ls.gam && echo "GAM identical across forms"
timeout 1200 python bench.py --no-cpu-baseline > $O/r03o_bench_c2.json 2> $O/r03o_bench_c2.err
tail -3 $O/r03o_bench_c2.err
GCGPU_K1_OCC=1 timeout 900 python bench.py --no-cpu-baseline > $O/r03o_bench_c2_occ1.json 2> $O/r03o_bench_c2_occ1.err
timeout 1200 python bench.py --no-cpu-baseline > $O/r03o_bench_c2_again.json 2> $O/r03o_bench_c2_again.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r03o_bench*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "value %.1f Mbp/s e2e %.1f Mbp/s" % ((d["value"] or 0)/1e6, (d["e2e"]["value"] or 0)/1e6), {k: round(v,1) for k,v in d["kernels_ms_per_step"].items()}, d.get("parity_on_sample"), d["roofline"]["frac"])
    except Exception as e: print(f, "failed", e)
PY
