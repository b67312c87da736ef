cd "$GRAFT_REPO_ROOT"; O=gpurun_out; mkdir -p $O
for W in 16 8; do echo "== parity with $W lanes per long item"; GCGPU_K1_LONG_WIDTH=$W timeout 600 python -m pytest tests -m gpu -x -q -k "k1 or pipeline" 2>&1 | tail -2; done
python -c "
from graphchainer_b200 import synth
print(synth.make_workload('c2', '/tmp/c2s', n_reads=839))
"
D=graphchainer_b200/GraphChainerB200
$D -g /tmp/c2s.gfa --gc-save-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o.gam -t 16 --gc-streams 1 > /dev/null 2>&1
for W in 32 16 8; do
for i in 1 2; do GCGPU_K1_LONG_WIDTH=$W GCGPU_TRACE=1 GC_TRACE=1 $D --gc-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o.gam -t 16 --gc-streams 1 > $O/r01p_trace_w$W.txt 2>&1; done
echo "W=$W"; grep -E "k1 \(long" $O/r01p_trace_w$W.txt | awk '{print $5, $6}' | tr '\n' ' '; echo
done
for cfg in 32:5 16:5 8:5 8:8; do W=${cfg%%:*}; MB=${cfg##*:}
GCGPU_K1_LONG_WIDTH=$W GCGPU_K1_LONG_BLOCKS=$MB timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r01p_w${W}_mb$MB.json 2> $O/r01p_w${W}_mb$MB.err
python -c "
import sys, json
l = json.loads(open('$O/r01p_w${W}_mb$MB.json').read().strip().splitlines()[-1])
print('W=$W MB=$MB', json.dumps({k: l[k] for k in ('value', 'ms_per_step', 'e2e', 'kernels_ms_per_step')}))"
done
