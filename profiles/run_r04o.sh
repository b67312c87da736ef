#!/bin/bash
# r04o: the round's final state: all GPU tests, smoke(), bench c2 with the reference sample
cd "$GRAFT_REPO_ROOT"
O=gpurun_out; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/r04o_tests.log 2>&1
tail -3 $O/r04o_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r04o_smoke.log 2>&1; tail -1 $O/r04o_smoke.log
timeout 1500 python bench.py > $O/r04o_bench_c2.json 2> $O/r04o_bench_c2.err
tail -2 $O/r04o_bench_c2.err | cut -c1-300
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r04o_bench_c2.json") if l.startswith("{")][-1])
e=d["e2e"]
print("value %.1f Mbp/s e2e %.1f Mbp/s (%.0f ms/step)" % ((d["value"] or 0)/1e6, (e["value"] or 0)/1e6, e["ms_per_step"]), d.get("parity_on_sample"), "value-config mismatches", d.get("value_config_summary_mismatches"), d.get("accuracy_on_sample"), (d.get("cpu_baseline") or {}).get("value"))
PY
