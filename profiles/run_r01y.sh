cd "$GRAFT_REPO_ROOT"; O=gpurun_out; mkdir -p $O
python -c "
from graphchainer_b200 import synth
print(synth.make_workload('c2', '/tmp/c2s', n_reads=839))
"
D=graphchainer_b200/GraphChainerB200
$D -g /tmp/c2s.gfa --gc-save-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o.gam -t 16 --gc-streams 1 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gc_k1_long_kernel" -s 3 -c 1 -o $O/r01y_fwd_round4 -f $D --gc-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o.gam -t 16 --gc-streams 1 > $O/r01y_ncu.log 2>&1
tail -2 $O/r01y_ncu.log
