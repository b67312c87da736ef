#!/bin/bash
# r04p: gc_k3l_level_kernel at 128 registers (launch bounds 128 x 4; 255 before): K2/K3 tests, K3 kernel time of a c2 step
cd "$GRAFT_REPO_ROOT"
O=gpurun_out; mkdir -p $O
( timeout 600 python -m pytest tests/test_k2_k3.py -m gpu -x -q ) > $O/r04p_tests.log 2>&1
tail -1 $O/r04p_tests.log
timeout 600 python bench.py --no-cpu-baseline > $O/r04p_bench_c2.json 2> $O/r04p_bench_c2.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r04p_bench_c2.json") if l.startswith("{")][-1])
print("value %.1f e2e %.1f" % (d["value"]/1e6, d["e2e"]["value"]/1e6), d["kernels_ms_per_step"], d.get("value_config_summary_mismatches"))
PY
