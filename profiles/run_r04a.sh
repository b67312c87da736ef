#!/bin/bash
# r04a: default batches = one round of equal batches over the workers (26 Mbp at most, 8 Mbp at least): all four configurations
cd "$GRAFT_REPO_ROOT"
O=gpurun_out; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/r04a_tests.log 2>&1
tail -3 $O/r04a_tests.log
timeout 1500 python bench.py > $O/r04a_bench_c2.json 2> $O/r04a_bench_c2.err
for W in c3 c4 c5; do
timeout 1500 python bench.py --workload $W --no-cpu-baseline > $O/r04a_bench_$W.json 2> $O/r04a_bench_$W.err
tail -1 $O/r04a_bench_$W.err | cut -c1-300
done
timeout 1500 python bench.py --no-cpu-baseline --host-threads 4 > $O/r04a_bench_c2_t4.json 2> $O/r04a_bench_c2_t4.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r04a_bench*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "value %.1f Mbp/s e2e %.1f Mbp/s" % ((d["value"] or 0)/1e6, (d["e2e"]["value"] or 0)/1e6), {k: round(v,1) for k,v in d["kernels_ms_per_step"].items()}, d.get("parity_on_sample"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(f, "failed", e)
PY
