#!/bin/bash
# r03x: c4 (ultra-long reads) after r03w: slabs released before the dense trace buffer grows; device buffer growth traced
cd "$GRAFT_REPO_ROOT"
O=gpurun_out; mkdir -p $O
for W in c4; do
( time GCGPU_TRACE_MEM=1 timeout 1500 python bench.py --workload $W ) > $O/r03x_bench_$W.json 2> $O/r03x_bench_$W.err
grep "device buffer grows" $O/r03x_bench_$W.err | awk '$6 > 20' | head -40
tail -3 $O/r03x_bench_$W.err | cut -c1-400
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r03x_bench*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "value %.1f Mbp/s e2e %.1f Mbp/s" % ((d["value"] or 0)/1e6, (d["e2e"]["value"] or 0)/1e6), {k: round(v,1) for k,v in d["kernels_ms_per_step"].items()}, d.get("parity_on_sample"), d["roofline"]["frac"], (d.get("cpu_baseline") or {}).get("value"), "index_s %.1f" % d.get("index_build_s", 0))
    except Exception as e: print(f, "failed", e)
PY
