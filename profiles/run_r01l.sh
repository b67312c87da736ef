cd "$GRAFT_REPO_ROOT"; O=gpurun_out; mkdir -p $O
python -c "
from graphchainer_b200 import synth
print(synth.make_workload('c2', '/tmp/c2s', n_reads=839))
"
D=graphchainer_b200/GraphChainerB200
$D -g /tmp/c2s.gfa --gc-save-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o.gam -t 16 --gc-streams 1 > /dev/null 2>&1
for i in 1 2; do GCGPU_TRACE=1 GC_TRACE=1 $D --gc-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o.gam -t 16 --gc-streams 1 > $O/r01l_trace.txt 2>&1; done
grep -E "k1 \(long" $O/r01l_trace.txt | tr '\n' ' '; echo
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gc_k1_long_kernel" -s 1 -c 1 -o $O/r01l_gc_k1_long_kernel -f $D --gc-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o.gam -t 16 --gc-streams 1 > $O/r01l_ncu.log 2>&1
tail -2 $O/r01l_ncu.log
timeout 300 python -m pytest tests -m gpu -x -q -k "k1 or pipeline" 2>&1 | tail -2
