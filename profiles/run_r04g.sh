#!/bin/bash
# r04g: gcalign_submit / gcalign_wait (worker pool, calls as a stream): GPU tests, bench c2 / c3 with the reference sample, c4 / c5
cd "$GRAFT_REPO_ROOT"
O=gpurun_out; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/r04g_tests.log 2>&1
tail -3 $O/r04g_tests.log
timeout 1500 python bench.py > $O/r04g_bench_c2.json 2> $O/r04g_bench_c2.err
tail -2 $O/r04g_bench_c2.err | cut -c1-300
timeout 1500 python bench.py --workload c3 > $O/r04g_bench_c3.json 2> $O/r04g_bench_c3.err
for W in c5 c4; do
timeout 1500 python bench.py --workload $W --no-cpu-baseline > $O/r04g_bench_$W.json 2> $O/r04g_bench_$W.err
tail -1 $O/r04g_bench_$W.err | cut -c1-300
done
timeout 1500 python bench.py --no-cpu-baseline --host-threads 4 > $O/r04g_bench_c2_t4.json 2> $O/r04g_bench_c2_t4.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r04g_bench*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "value %.1f Mbp/s e2e %.1f Mbp/s (one call at a time %.1f)" % ((d["value"] or 0)/1e6, (d["e2e"]["value"] or 0)/1e6, d["e2e"]["one_call_at_a_time"]["value"]/1e6), d.get("parity_on_sample"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(f, "failed", e)
PY
