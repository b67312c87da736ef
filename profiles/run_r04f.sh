#!/bin/bash
# r04f: whole-read K1 launches on a low-priority stream (the batch's other kernels on a high-priority one): A/B on c2 and c3
cd "$GRAFT_REPO_ROOT"
O=gpurun_out; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q -k "k1 or pipeline or align" ) > $O/r04f_tests.log 2>&1
tail -3 $O/r04f_tests.log
for P in 1 0 1 0; do
GCGPU_K1_PRIORITY=$P timeout 900 python bench.py --no-cpu-baseline > $O/r04f_bench_c2_prio${P}_$RANDOM.json 2> $O/r04f_err.txt
done
for P in 1 0; do
GCGPU_K1_PRIORITY=$P timeout 900 python bench.py --workload c3 --no-cpu-baseline > $O/r04f_bench_c3_prio$P.json 2> $O/r04f_err.txt
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r04f_bench*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "value %.1f Mbp/s e2e %.1f Mbp/s" % ((d["value"] or 0)/1e6, (d["e2e"]["value"] or 0)/1e6), {k: round(v,1) for k,v in d["kernels_ms_per_step"].items()})
    except Exception as e: print(f, "failed", e)
PY
