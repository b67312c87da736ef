cd "$GRAFT_REPO_ROOT"; O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "
from graphchainer_b200 import synth
print(synth.make_workload('c2', '/tmp/c2s', n_reads=839))
"
D=graphchainer_b200/GraphChainerB200
$D -g /tmp/c2s.gfa --gc-save-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o.gam -t 16 --gc-streams 1 > /dev/null 2>&1
for i in 1 2 3; do GCGPU_TRACE=1 GC_TRACE=1 $D --gc-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o.gam -t 16 --gc-streams 1 > $O/r01t_trace.txt 2>&1; done
grep -E "k1 \(long" $O/r01t_trace.txt | awk '{print $5, $6}' | tr '\n' ' '; echo
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r01t_a.json 2> $O/r01t_a.err
python -c "
import sys, json
l = json.loads(open('$O/r01t_a.json').read().strip().splitlines()[-1])
print('a', json.dumps({k: l[k] for k in ('value', 'ms_per_step', 'e2e', 'kernels_ms_per_step', 'gpu_launches')}))"
