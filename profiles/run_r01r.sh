cd "$GRAFT_REPO_ROOT"; O=gpurun_out; mkdir -p $O
python -c "
from graphchainer_b200 import synth
print(synth.make_workload('c2', '/tmp/c2m', n_reads=3400))
"
D=graphchainer_b200/GraphChainerB200
$D -g /tmp/c2m.gfa --gc-save-index /tmp/c2m.gcidx -f /tmp/c2m.fa -a /tmp/o.gam -t 16 --gc-streams 1 --gc-max-reads 100 > /dev/null 2>&1
for i in 1 2; do GCGPU_TRACE=1 GC_TRACE=1 $D --gc-index /tmp/c2m.gcidx -f /tmp/c2m.fa -a /tmp/o.gam -t 16 --gc-streams 1 --gc-batch-bp 40000000 > $O/r01r_trace_lockstep_3400.txt 2>&1; done
echo lockstep; grep -E "k1 \(long" $O/r01r_trace_lockstep_3400.txt | awk '{print $4, $5, $6}' | tr '\n' ' '; echo
for i in 1 2; do GCGPU_K1_SIMT_MIN=3000 GCGPU_TRACE=1 GC_TRACE=1 $D --gc-index /tmp/c2m.gcidx -f /tmp/c2m.fa -a /tmp/o1.gam -t 16 --gc-streams 1 --gc-batch-bp 40000000 > $O/r01r_trace_simt_3400.txt 2>&1; done
echo simt; grep -E "k1 \(long" $O/r01r_trace_simt_3400.txt | awk '{print $4, $5, $6}' | tr '\n' ' '; echo
cmp /tmp/o.gam /tmp/o1.gam && echo same-gam
