cd "$GRAFT_REPO_ROOT"; O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "
from graphchainer_b200 import synth
print(synth.make_workload('c2', '/tmp/c2s', n_reads=839))
"
D=graphchainer_b200/GraphChainerB200
$D -g /tmp/c2s.gfa --gc-save-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o.gam -t 16 --gc-streams 1 > /dev/null 2>&1
for i in 1 2; do GCGPU_TRACE=1 GC_TRACE=1 $D --gc-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o.gam -t 16 --gc-streams 1 > $O/r01s_trace.txt 2>&1; done
grep -E "k3" $O/r01s_trace.txt | tr '\n' ' '; echo
for i in 1 2; do GCGPU_K3_PATH_LEVELS=0 GCGPU_TRACE=1 GC_TRACE=1 $D --gc-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o0.gam -t 16 --gc-streams 1 > $O/r01s_trace_dfs.txt 2>&1; done
grep -E "k3" $O/r01s_trace_dfs.txt | tr '\n' ' '; echo
cmp /tmp/o.gam /tmp/o0.gam && echo same-gam
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r01s_a.json 2> $O/r01s_a.err
python -c "
import sys, json
l = json.loads(open('$O/r01s_a.json').read().strip().splitlines()[-1])
print('a', json.dumps({k: l[k] for k in ('value', 'ms_per_step', 'e2e', 'kernels_ms_per_step', 'gpu_launches')}))"
