#!/bin/bash
# r04j: final state: GPU tests, smoke(), bench c2 (reference sample, parity gate) and c5
cd "$GRAFT_REPO_ROOT"
O=gpurun_out; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/r04j_tests.log 2>&1
tail -3 $O/r04j_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r04j_smoke.log 2>&1; tail -1 $O/r04j_smoke.log
timeout 1500 python bench.py > $O/r04j_bench_c2.json 2> $O/r04j_bench_c2.err
tail -2 $O/r04j_bench_c2.err | cut -c1-300
timeout 1500 python bench.py --workload c5 --no-cpu-baseline > $O/r04j_bench_c5.json 2> $O/r04j_bench_c5.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r04j_bench*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        e=d["e2e"]
        print(f.split("/")[-1], "value %.1f Mbp/s e2e %.1f Mbp/s (%.0f ms/step)" % ((d["value"] or 0)/1e6, (e["value"] or 0)/1e6, e["ms_per_step"]), d.get("parity_on_sample"), d.get("accuracy_on_sample"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(f, "failed", e)
PY
