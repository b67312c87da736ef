cd "$GRAFT_REPO_ROOT"; O=gpurun_out; mkdir -p $O
python -c "
from graphchainer_b200 import synth
print(synth.make_workload('c2', '/tmp/c2s', n_reads=839))
"
D=graphchainer_b200/GraphChainerB200
$D -g /tmp/c2s.gfa --gc-save-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o.gam -t 16 --gc-streams 1 > /dev/null 2>&1
for spec in "gc_k1_kernel:4" "gc_k1_bt_kernel:4"; do K=${spec%%:*}; S=${spec##*:}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$K" -s $S -c 1 -o $O/r02f_${K}_s2 -f $D --gc-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o.gam -t 16 --gc-streams 1 > $O/r02f_$K.log 2>&1
tail -1 $O/r02f_$K.log
python profiles/ncu_summary.py kernel $O/r02f_${K}_s2.ncu-rep >> $O/r02f_ncu_s2_summary.txt
done
