cd "$GRAFT_REPO_ROOT"; O=gpurun_out; mkdir -p $O
bash profiles/gpu_round.sh r01o
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r01o_bench_reference.json 2> $O/r01o_bench_reference.err
tail -1 $O/r01o_bench_reference.json | cut -c1-400
GCGPU_K1_LONG_BLOCKS=8 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r01o_mb8.json 2> $O/r01o_mb8.err
python -c "
import sys, json
l = json.loads(open('$O/r01o_mb8.json').read().strip().splitlines()[-1])
print('mb8', json.dumps({k: l[k] for k in ('value', 'ms_per_step', 'e2e', 'kernels_ms_per_step', 'gpu_launches')}))"
