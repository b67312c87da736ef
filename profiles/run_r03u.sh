#!/bin/bash
# r03u: record-size kernel one record per warp; full GPU tests; bench c2 (16 and 4 host threads); launch list; ncu --set full of the
# kernels changed this round (K1 lane-per-item forward / backtrace = the dominant launch, GAM records, K2 sweep, K3 block form)
cd "$GRAFT_REPO_ROOT"
O=gpurun_out; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/r03u_tests.log 2>&1
tail -3 $O/r03u_tests.log
python -c "
from graphchainer_b200 import synth
print(synth.make_workload('c2', '/tmp/c2s', n_reads=1678))
"
D=graphchainer_b200/GraphChainerB200
$D -g /tmp/c2s.gfa --gc-save-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o.gam -t 16 --gc-streams 1 > /dev/null 2>&1
GCGPU_TRACE=1 GC_TRACE=1 $D --gc-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o.gam -t 16 --gc-streams 1 > /dev/null 2>&1
GCGPU_TRACE=1 GC_TRACE=1 $D --gc-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o.gam -t 16 --gc-streams 1 > $O/r03u_trace.txt 2>&1
grep "gcgpu\]\|phase" $O/r03u_trace.txt | grep -v hint | tail -40
timeout 1200 python bench.py > $O/r03u_bench_c2.json 2> $O/r03u_bench_c2.err
tail -3 $O/r03u_bench_c2.err
timeout 1200 python bench.py --no-cpu-baseline --host-threads 4 > $O/r03u_bench_c2_t4.json 2> $O/r03u_bench_c2_t4.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r03u_bench*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "value %.1f Mbp/s e2e %.1f Mbp/s" % ((d["value"] or 0)/1e6, (d["e2e"]["value"] or 0)/1e6), {k: round(v,1) for k,v in d["kernels_ms_per_step"].items()}, d.get("parity_on_sample"), d["roofline"]["frac"], (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(f, "failed", e)
PY
[ "$1" = "skip-ncu" ] && exit 0
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/r03u_launches.csv python bench.py --reads 1678 --steps 1 --warmup 3 --no-cpu-baseline > $O/r03u_launches_bench.log 2>&1
python profiles/ncu_summary.py launches $O/r03u_launches.csv > $O/r03u_launches_c2_1678reads.txt 2>&1
head -30 $O/r03u_launches_c2_1678reads.txt
for spec in "gc_k1s_forward_kernel:0" "gc_k1s_backtrace_kernel:0" "gc_gam_kernel:1" "gc_gam_size_kernel:1"; do
	K=${spec%%:*}; S=${spec##*:}; F=${K//[<>]/_}
	timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$K" -s $S -c 1 -o $O/r03u_$F -f $D --gc-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o.gam -t 16 --gc-streams 1 > $O/r03u_$F.log 2>&1
	tail -1 $O/r03u_$F.log
	python profiles/ncu_summary.py kernel $O/r03u_$F.ncu-rep >> $O/r03u_ncu_full_summary.txt 2>/dev/null
done
python profiles/ncu_summary.py dominant $O/r03u_ncu_dominant_launch.json "S1 round 2 of a 1678-read c2 batch: 18116 whole-read extensions (lane-per-item kernels, forward + backtrace)" $O/r03u_gc_k1s_forward_kernel.ncu-rep $O/r03u_gc_k1s_backtrace_kernel.ncu-rep > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gc_k2_chain_kernel" -s 0 -c 1 -o $O/r03u_gc_k2_chain_kernel_sweep -f python -m pytest tests/test_k2_k3.py -m gpu -q -k "ultralong_anchor_sets" > $O/r03u_gc_k2_chain_kernel_sweep.log 2>&1
tail -1 $O/r03u_gc_k2_chain_kernel_sweep.log
python profiles/ncu_summary.py kernel $O/r03u_gc_k2_chain_kernel_sweep.ncu-rep >> $O/r03u_ncu_full_summary.txt 2>/dev/null
python -c "
from graphchainer_b200 import synth
print(synth.make_workload('c4', '/tmp/c4s', n_reads=40))
"
$D -g /tmp/c4s.gfa --gc-save-index /tmp/c4s.gcidx -f /tmp/c4s.fa -a /tmp/o4.gam -t 16 --gc-streams 1 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gc_k3b_distance_kernel" -s 0 -c 1 -o $O/r03u_gc_k3b_distance_kernel -f $D --gc-index /tmp/c4s.gcidx -f /tmp/c4s.fa -a /tmp/o4.gam -t 16 --gc-streams 1 > $O/r03u_gc_k3b_distance_kernel.log 2>&1
tail -1 $O/r03u_gc_k3b_distance_kernel.log
python profiles/ncu_summary.py kernel $O/r03u_gc_k3b_distance_kernel.ncu-rep >> $O/r03u_ncu_full_summary.txt 2>/dev/null
for F in gc_gam_size_kernel gc_k2_chain_kernel_sweep gc_k3b_distance_kernel; do rm -f $O/r03u_$F.ncu-rep; done
cat $O/r03u_ncu_full_summary.txt | head -150
