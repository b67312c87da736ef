#!/bin/bash
# r04c: c3 and c5 with the default batches, reference sample and parity gate; c3 with 4 host threads
cd "$GRAFT_REPO_ROOT"
O=gpurun_out; mkdir -p $O
for W in c3 c5; do
( time timeout 1500 python bench.py --workload $W ) > $O/r04c_bench_$W.json 2> $O/r04c_bench_$W.err
tail -1 $O/r04c_bench_$W.err | cut -c1-300
done
timeout 1500 python bench.py --workload c3 --no-cpu-baseline --host-threads 4 > $O/r04c_bench_c3_t4.json 2> $O/r04c_bench_c3_t4.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r04c_bench*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "value %.1f Mbp/s e2e %.1f Mbp/s" % ((d["value"] or 0)/1e6, (d["e2e"]["value"] or 0)/1e6), {k: round(v,1) for k,v in d["kernels_ms_per_step"].items()}, d.get("parity_on_sample"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(f, "failed", e)
PY
