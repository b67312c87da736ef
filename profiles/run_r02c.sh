cd "$GRAFT_REPO_ROOT"; O=gpurun_out; mkdir -p $O
nproc; uptime
for i in 1 2; do
GC_TRACE_CALL=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r02c_$i.json 2> $O/r02c_$i.err
python -c "
import sys, json
l = json.loads(open('$O/r02c_$i.json').read().strip().splitlines()[-1])
print('run $i', json.dumps({k: l[k] for k in ('value', 'e2e', 'kernels_ms_per_step')}))"
grep "gcalign\] batches" $O/r02c_$i.err | head -6
done
grep "worker" $O/r02c_2.err | sed -n 40,52p
