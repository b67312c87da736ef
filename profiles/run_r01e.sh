cd "$GRAFT_REPO_ROOT"; O=gpurun_out; mkdir -p $O
python -c "
from graphchainer_b200 import synth
print(synth.make_workload('c2', '/tmp/c2m', n_reads=3400))
print(synth.make_workload('c2', '/tmp/c2s', n_reads=839))
"
D=graphchainer_b200/GraphChainerB200
$D -g /tmp/c2m.gfa --gc-save-index /tmp/c2m.gcidx -f /tmp/c2m.fa -a /tmp/o.gam -t 16 --gc-streams 1 --gc-max-reads 100 > /dev/null 2>&1
# one batch of 3400 reads (36k long items in round 2): lock-step vs SIMT form, second run of each is the warm one
for i in 1 2; do
GCGPU_TRACE=1 GC_TRACE=1 $D --gc-index /tmp/c2m.gcidx -f /tmp/c2m.fa -a /tmp/o.gam -t 16 --gc-streams 1 --gc-batch-bp 40000000 > $O/r01e_trace_lockstep_3400.txt 2>&1
done
for i in 1 2; do
GCGPU_K1_SIMT_MIN=3000 GCGPU_TRACE=1 GC_TRACE=1 $D --gc-index /tmp/c2m.gcidx -f /tmp/c2m.fa -a /tmp/o.gam -t 16 --gc-streams 1 --gc-batch-bp 40000000 > $O/r01e_trace_simt_3400.txt 2>&1
done
GCGPU_K1_SIMT_MIN=3000 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^gc_k1_long_simt_kernel" -s 0 -c 1 -o $O/r01e_k1_simt -f $D --gc-index /tmp/c2m.gcidx -f /tmp/c2m.fa -a /tmp/o.gam -t 16 --gc-streams 1 --gc-max-reads 839 > $O/r01e_k1_simt.log 2>&1
tail -2 $O/r01e_k1_simt.log
bash profiles/sweep.sh r01e "8388608 33554432 134217728" "4 2"
