#!/bin/bash
# r04b: c4 with the default batches (three in flight for ultra-long reads), reference sample and parity gate
cd "$GRAFT_REPO_ROOT"
O=gpurun_out; mkdir -p $O
( time timeout 1500 python bench.py --workload c4 ) > $O/r04b_bench_c4.json 2> $O/r04b_bench_c4.err
tail -3 $O/r04b_bench_c4.err | cut -c1-400
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r04b_bench*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "value %.1f Mbp/s e2e %.1f Mbp/s" % ((d["value"] or 0)/1e6, (d["e2e"]["value"] or 0)/1e6), {k: round(v,1) for k,v in d["kernels_ms_per_step"].items()}, d.get("parity_on_sample"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(f, "failed", e)
PY
