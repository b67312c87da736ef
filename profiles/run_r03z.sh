#!/bin/bash
# r03z: c3 (150 Mbp per step) with 6, 9 (default 16.8 Mbp batches) and 12 batches on the 6 streams
cd "$GRAFT_REPO_ROOT"
O=gpurun_out; mkdir -p $O
for B in 25200000 12600000 0; do
timeout 1200 python bench.py --workload c3 --no-cpu-baseline --batch-bp $B > $O/r03z_bench_c3_b$B.json 2> $O/r03z_bench_c3_b$B.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r03z_bench*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "value %.1f Mbp/s e2e %.1f Mbp/s" % ((d["value"] or 0)/1e6, (d["e2e"]["value"] or 0)/1e6), {k: round(v,1) for k,v in d["kernels_ms_per_step"].items()})
    except Exception as e: print(f, "failed", e)
PY
