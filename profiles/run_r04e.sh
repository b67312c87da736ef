#!/bin/bash
# r04e: final state of the round: GPU tests, smoke(), bench c2 with the reference sample (parity gate, accuracy_on_sample)
cd "$GRAFT_REPO_ROOT"
O=gpurun_out; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/r04e_tests.log 2>&1
tail -3 $O/r04e_tests.log
( time python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > $O/r04e_smoke.log 2>&1
tail -4 $O/r04e_smoke.log
timeout 1500 python bench.py > $O/r04e_bench_c2.json 2> $O/r04e_bench_c2.err
tail -2 $O/r04e_bench_c2.err | cut -c1-300
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 1 ) > $O/r04e_bench_c2_reference.json 2> $O/r04e_bench_c2_reference.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r04e_bench*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "value %.1f Mbp/s e2e %.1f Mbp/s" % ((d["value"] or 0)/1e6, (d["e2e"]["value"] or 0)/1e6), d.get("parity_on_sample"), d.get("accuracy_on_sample"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(f, "failed", e)
PY
