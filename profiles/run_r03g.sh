#!/bin/bash
# r03g: sweep-form K2 (parity tests incl. MPC width 16 and ~2 k anchors per read), bench lines for c2 / c3 / c4 / c5
cd "$GRAFT_REPO_ROOT"
O=gpurun_out; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/r03g_tests.log 2>&1
tail -3 $O/r03g_tests.log
timeout 1200 python bench.py > $O/r03g_bench_c2.json 2> $O/r03g_bench_c2.err
tail -2 $O/r03g_bench_c2.err
timeout 1500 python bench.py --workload c5 --reads 2500 > $O/r03g_bench_c5.json 2> $O/r03g_bench_c5.err
tail -2 $O/r03g_bench_c5.err
timeout 1800 python bench.py --workload c3 --reads 10000 > $O/r03g_bench_c3.json 2> $O/r03g_bench_c3.err
tail -2 $O/r03g_bench_c3.err
timeout 1800 python bench.py --workload c4 --reads 600 --steps 2 > $O/r03g_bench_c4.json 2> $O/r03g_bench_c4.err
tail -2 $O/r03g_bench_c4.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r03g_bench*.json")):
    try:
        d=json.load(open(f))
        print(f.split("/")[-1], "value %.1f Mbp/s e2e %.1f Mbp/s" % ((d["value"] or 0)/1e6, (d["e2e"]["value"] or 0)/1e6), {k: round(v,1) for k,v in d["kernels_ms_per_step"].items()}, d.get("parity_on_sample"), (d.get("cpu_baseline") or {}).get("value"), "index_s", round(d["index_build_s"],1))
    except Exception as e: print(f, "failed", e)
PY
