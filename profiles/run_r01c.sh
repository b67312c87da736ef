cd "$GRAFT_REPO_ROOT"; O=gpurun_out; mkdir -p $O
python -c "
from graphchainer_b200 import synth
print(synth.make_workload('c2', '/tmp/c2m', n_reads=3400))
"
D=graphchainer_b200/GraphChainerB200
$D -g /tmp/c2m.gfa --gc-save-index /tmp/c2m.gcidx -f /tmp/c2m.fa -a /tmp/o.gam -t 16 --gc-streams 1 --gc-max-reads 100 > /dev/null 2>&1
for i in 1 2; do
GCGPU_TRACE=1 GC_TRACE=1 $D --gc-index /tmp/c2m.gcidx -f /tmp/c2m.fa -a /tmp/o.gam -t 16 --gc-streams 1 --gc-batch-bp 17000000 > $O/r01c_trace_simt.txt 2>&1
done
GCGPU_K1_SIMT_MIN=100000000 GCGPU_TRACE=1 GC_TRACE=1 $D --gc-index /tmp/c2m.gcidx -f /tmp/c2m.fa -a /tmp/o.gam -t 16 --gc-streams 1 --gc-batch-bp 17000000 > $O/r01c_trace_lockstep.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^gc_k1_long_simt_kernel" -s 1 -c 1 -o $O/r01c_k1_simt -f $D --gc-index /tmp/c2m.gcidx -f /tmp/c2m.fa -a /tmp/o.gam -t 16 --gc-streams 1 --gc-batch-bp 17000000 --gc-max-reads 1700 > $O/r01c_k1_simt.log 2>&1
tail -2 $O/r01c_k1_simt.log
