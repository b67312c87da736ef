#!/bin/bash
# r04i: bench c2 (reference sample, parity gate with the single-threaded re-check) and c5, calls streamed and one at a time
cd "$GRAFT_REPO_ROOT"
O=gpurun_out; mkdir -p $O
timeout 1500 python bench.py > $O/r04i_bench_c2.json 2> $O/r04i_bench_c2.err
tail -2 $O/r04i_bench_c2.err | cut -c1-300
GCGPU_TRACE_MEM=1 timeout 1500 python bench.py --workload c5 --no-cpu-baseline > $O/r04i_bench_c5.json 2> $O/r04i_bench_c5.err
grep -c "device buffer grows" $O/r04i_bench_c5.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r04i_bench*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        e=d["e2e"]
        print(f.split("/")[-1], "value %.1f Mbp/s e2e %.1f Mbp/s (%.0f ms/step; one call at a time %.1f, %.0f ms/step)" % ((d["value"] or 0)/1e6, (e["value"] or 0)/1e6, e["ms_per_step"], e["one_call_at_a_time"]["value"]/1e6, e["one_call_at_a_time"]["ms_per_step"]), d.get("parity_on_sample"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(f, "failed", e)
PY
