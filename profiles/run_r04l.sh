#!/bin/bash
# r04l: GPU tests incl. the lane-per-item kernels on 2-3 blocks (every lane works through many items); bench c2 with the summaries of
# the one-batch (kernel-time) configuration checked against the end-to-end ones
cd "$GRAFT_REPO_ROOT"
O=gpurun_out; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/r04l_tests.log 2>&1
tail -3 $O/r04l_tests.log
timeout 1500 python bench.py > $O/r04l_bench_c2.json 2> $O/r04l_bench_c2.err
tail -2 $O/r04l_bench_c2.err | cut -c1-300
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r04l_bench*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        e=d["e2e"]
        print(f.split("/")[-1], "value %.1f Mbp/s e2e %.1f Mbp/s (%.0f ms/step)" % ((d["value"] or 0)/1e6, (e["value"] or 0)/1e6, e["ms_per_step"]), d.get("parity_on_sample"), "value-config mismatches", d.get("value_config_summary_mismatches"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(f, "failed", e)
PY
