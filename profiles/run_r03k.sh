#!/bin/bash
# r03k: c3 (BASELINE configs[2]) on one GPU with 16 and 4 host threads; warm per-phase trace of a second batch
cd "$GRAFT_REPO_ROOT"
O=gpurun_out; mkdir -p $O
timeout 1800 python bench.py --workload c3 --reads 10000 > $O/r03k_bench_c3.json 2> $O/r03k_bench_c3.err
tail -2 $O/r03k_bench_c3.err
timeout 1800 python bench.py --workload c3 --reads 10000 --host-threads 4 --no-cpu-baseline > $O/r03k_bench_c3_t4.json 2> $O/r03k_bench_c3_t4.err
python -c "
from graphchainer_b200 import synth
print(synth.make_workload('c2', '/tmp/c2s', n_reads=3356))
"
D=graphchainer_b200/GraphChainerB200
$D -g /tmp/c2s.gfa --gc-save-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o.gam -t 16 --gc-streams 1 > /dev/null 2>&1
GCGPU_TRACE=1 GC_TRACE=1 $D --gc-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o.gam -t 16 --gc-streams 1 > $O/r03k_trace_two_batches.txt 2>&1
grep "gcgpu\]\|phase" $O/r03k_trace_two_batches.txt | grep -v hint | tail -34
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r03k_bench*.json")):
    try:
        d=json.load(open(f))
        print(f.split("/")[-1], "value %.1f Mbp/s e2e %.1f Mbp/s" % ((d["value"] or 0)/1e6, (d["e2e"]["value"] or 0)/1e6), {k: round(v,1) for k,v in d["kernels_ms_per_step"].items()}, d.get("parity_on_sample"), (d.get("cpu_baseline") or {}).get("value"), "index_s", round(d["index_build_s"],1))
    except Exception as e: print(f, "failed", e)
PY
