#!/bin/bash
# One GPU-box visit: parity tests, bench line, per-kernel trace, ncu launch list, ncu --set full of the hot kernels.
# usage (via gpurun): bash profiles/gpu_round.sh <tag> [skip-ncu]
cd "$GRAFT_REPO_ROOT"
TAG=${1:-r01}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
nproc >> $O/${TAG}_smi.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/${TAG}_tests.log 2>&1
tail -3 $O/${TAG}_tests.log
timeout 900 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
cat $O/${TAG}_bench.json
python -c "
from graphchainer_b200 import synth
print(synth.make_workload('c2', '/tmp/c2s', n_reads=839))
"
D=graphchainer_b200/GraphChainerB200
$D -g /tmp/c2s.gfa --gc-save-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o.gam -t 16 --gc-streams 1 > /dev/null 2>&1
GCGPU_TRACE=1 GC_TRACE=1 $D --gc-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o.gam -t 16 --gc-streams 1 > $O/${TAG}_trace.txt 2>&1
[ "$2" = "skip-ncu" ] && exit 0
# launch list of the bench command (small read set so that the serialised replay stays short)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/${TAG}_launches.csv python bench.py --reads 839 --steps 1 --warmup 3 --no-cpu-baseline > $O/${TAG}_launches_bench.log 2>&1
for spec in "gc_k1_long_kernel:1" "gc_k1_long_bt_kernel:1" "gc_k1_kernel:0" "gc_k1_bt_kernel:0" "gc_k3w_distance_kernel:1" "gc_k3l_level_kernel:0" "gc_seed_kernel:0"; do
	K=${spec%%:*}; S=${spec##*:}; F=${K//[<>]/_}
	timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$K" -s $S -c 1 -o $O/${TAG}_$F -f $D --gc-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o.gam -t 16 --gc-streams 1 > $O/${TAG}_$F.log 2>&1
	tail -2 $O/${TAG}_$F.log
	# the box may return at most 64 MiB: summarise here, keep only the reports of the dominant kernel pair
	python profiles/ncu_summary.py kernel $O/${TAG}_$F.ncu-rep >> $O/${TAG}_ncu_full_summary.txt 2>/dev/null
	case "$K" in gc_k1_long_kernel|gc_k1_long_bt_kernel) ;; *) rm -f $O/${TAG}_$F.ncu-rep ;; esac
done
ls -la $O
