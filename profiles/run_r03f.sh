#!/bin/bash
# r03f: parity tests, bench, batch-size sweep, launch list and ncu --set full captures of the new kernels
cd "$GRAFT_REPO_ROOT"
O=gpurun_out; mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/r03f_tests.log 2>&1
tail -3 $O/r03f_tests.log
python -c "
from graphchainer_b200 import synth
print(synth.make_workload('c2', '/tmp/c2s', n_reads=839))
"
D=graphchainer_b200/GraphChainerB200
$D -g /tmp/c2s.gfa --gc-save-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o.gam -t 16 --gc-streams 1 > /dev/null 2>&1
GCGPU_TRACE=1 GC_TRACE=1 $D --gc-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o.gam -t 16 --gc-streams 1 > /dev/null 2>&1
GCGPU_TRACE=1 GC_TRACE=1 $D --gc-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o.gam -t 16 --gc-streams 1 > $O/r03f_trace.txt 2>&1
grep "gcgpu\]\|phase" $O/r03f_trace.txt | grep -v hint
timeout 1200 python bench.py > $O/r03f_bench.json 2> $O/r03f_bench.err
tail -3 $O/r03f_bench.err
for cfg in "25000000 4" "17000000 6" "34000000 3" "100000000 1" "50000000 2"; do
set -- $cfg
timeout 900 python bench.py --no-cpu-baseline --batch-bp $1 --streams $2 > $O/r03f_bench_b$1_s$2.json 2> $O/r03f_bench_b$1_s$2.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r03f_bench*.json")):
    try:
        d=json.load(open(f))
        print(f.split("/")[-1], "value %.1f Mbp/s e2e %.1f Mbp/s" % ((d["value"] or 0)/1e6, (d["e2e"]["value"] or 0)/1e6), {k: round(v,1) for k,v in d["kernels_ms_per_step"].items()}, d.get("parity_on_sample"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(f, "failed", e)
PY
[ "$1" = "skip-ncu" ] && exit 0
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r03f_launches.csv python bench.py --reads 839 --steps 1 --warmup 3 --no-cpu-baseline > $O/r03f_launches_bench.log 2>&1
for spec in "gc_k1s_forward_kernel:0" "gc_k1s_backtrace_kernel:0" "gc_k1_kernel:0" "gc_k1_bt_kernel:0" "gc_k1_long_kernel:0" "gc_k3w_distance_kernel:1" "gc_k3l_level_kernel:0" "gc_piece_kernel:1" "gc_tokens_kernel:1"; do
	K=${spec%%:*}; S=${spec##*:}; F=${K//[<>]/_}
	timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$K" -s $S -c 1 -o $O/r03f_$F -f $D --gc-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o.gam -t 16 --gc-streams 1 > $O/r03f_$F.log 2>&1
	tail -1 $O/r03f_$F.log
	python profiles/ncu_summary.py kernel $O/r03f_$F.ncu-rep >> $O/r03f_ncu_full_summary.txt 2>/dev/null
done
python profiles/ncu_summary.py dominant $O/r03f_ncu_dominant_launch.json "S1 round 2 of an 839-read c2 batch: 9094 whole-read extensions (lane-per-item kernels, forward + backtrace)" $O/r03f_gc_k1s_forward_kernel.ncu-rep $O/r03f_gc_k1s_backtrace_kernel.ncu-rep > /dev/null 2>&1
# integer peak kernel under ncu (pipe utilisation of the peak measurement itself)
timeout 600 ncu --set full --clock-control none -k regex:gc_int_peak_kernel -s 2 -c 1 -o $O/r03f_gc_int_peak_kernel -f python -c "
import sys; sys.path.insert(0,'.')
from graphchainer_b200 import align
a = align.Aligner('/tmp/c2s.gfa', device=0, host_threads=8, streams=1)
print(a.int_peak())
" > $O/r03f_gc_int_peak_kernel.log 2>&1
python profiles/ncu_summary.py kernel $O/r03f_gc_int_peak_kernel.ncu-rep >> $O/r03f_ncu_full_summary.txt 2>/dev/null
for F in gc_k1_kernel gc_k1_bt_kernel gc_k1_long_kernel gc_k3w_distance_kernel gc_k3l_level_kernel gc_piece_kernel gc_tokens_kernel gc_int_peak_kernel; do rm -f $O/r03f_$F.ncu-rep; done
ls -la $O | tail -30
