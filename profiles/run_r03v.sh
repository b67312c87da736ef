#!/bin/bash
# r03v: the other configurations with this round's final code (c3 = default at N > 1, c4 ultra-long, c5 HiFi / wide path cover), each with
# the reference sample and the parity gate; ncu --set full of the S7 kernels
cd "$GRAFT_REPO_ROOT"
O=gpurun_out; mkdir -p $O
for W in c3 c4 c5; do
( time timeout 1500 python bench.py --workload $W ) > $O/r03v_bench_$W.json 2> $O/r03v_bench_$W.err
tail -2 $O/r03v_bench_$W.err
done
timeout 1200 python bench.py --workload c3 --no-cpu-baseline --host-threads 4 > $O/r03v_bench_c3_t4.json 2> $O/r03v_bench_c3_t4.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r03v_bench*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "value %.1f Mbp/s e2e %.1f Mbp/s" % ((d["value"] or 0)/1e6, (d["e2e"]["value"] or 0)/1e6), {k: round(v,1) for k,v in d["kernels_ms_per_step"].items()}, d.get("parity_on_sample"), d["roofline"]["frac"], (d.get("cpu_baseline") or {}).get("value"), "index_s %.1f" % d.get("index_build_s", 0))
    except Exception as e: print(f, "failed", e)
PY
python -c "
from graphchainer_b200 import synth
print(synth.make_workload('c2', '/tmp/c2s', n_reads=1678))
"
D=graphchainer_b200/GraphChainerB200
$D -g /tmp/c2s.gfa --gc-save-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o.gam -t 16 --gc-streams 1 > /dev/null 2>&1
for spec in "gc_gam_kernel:0" "gc_gam_size_kernel:0"; do
	K=${spec%%:*}; S=${spec##*:}; F=${K//[<>]/_}
	timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^$K" -s $S -c 1 -o $O/r03v_$F -f $D --gc-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o.gam -t 16 --gc-streams 1 > $O/r03v_$F.log 2>&1
	tail -1 $O/r03v_$F.log
	python profiles/ncu_summary.py kernel $O/r03v_$F.ncu-rep >> $O/r03v_ncu_full_summary_s7.txt 2>/dev/null
done
rm -f $O/r03v_gc_gam_size_kernel.ncu-rep
cat $O/r03v_ncu_full_summary_s7.txt | head -60
