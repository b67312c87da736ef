#!/bin/bash
# r04k: batches assigned to the workers round-robin (worker w: batches w, w + W, ...): c5 twice, c2, c3
cd "$GRAFT_REPO_ROOT"
O=gpurun_out; mkdir -p $O
for i in 1 2; do
GCGPU_TRACE_MEM=1 timeout 1500 python bench.py --workload c5 --no-cpu-baseline > $O/r04k_bench_c5_$i.json 2> $O/r04k_bench_c5_$i.err
grep -c "device buffer grows" $O/r04k_bench_c5_$i.err
done
timeout 1500 python bench.py --no-cpu-baseline > $O/r04k_bench_c2.json 2> $O/r04k_bench_c2.err
timeout 1500 python bench.py --workload c3 --no-cpu-baseline > $O/r04k_bench_c3.json 2> $O/r04k_bench_c3.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r04k_bench*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        e=d["e2e"]
        print(f.split("/")[-1], "value %.1f Mbp/s e2e %.1f Mbp/s (%.0f ms/step)" % ((d["value"] or 0)/1e6, (e["value"] or 0)/1e6, e["ms_per_step"]))
    except Exception as e: print(f, "failed", e)
PY
