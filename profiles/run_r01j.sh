cd "$GRAFT_REPO_ROOT"; O=gpurun_out; mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/r01j_tests.log 2>&1
tail -3 $O/r01j_tests.log
GC_TRACE_CALL=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --streams 6 --threads-per-stream 6 > $O/r01j_a.json 2> $O/r01j_a.err
MALLOC_MMAP_THRESHOLD_=1073741824 MALLOC_TRIM_THRESHOLD_=8589934592 MALLOC_TOP_PAD_=268435456 GC_TRACE_CALL=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --streams 6 --threads-per-stream 6 > $O/r01j_b.json 2> $O/r01j_b.err
MALLOC_MMAP_THRESHOLD_=1073741824 MALLOC_TRIM_THRESHOLD_=8589934592 MALLOC_TOP_PAD_=268435456 GC_TRACE_CALL=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --streams 4 --threads-per-stream 8 --batch-bp 16777216 > $O/r01j_c.json 2> $O/r01j_c.err
for f in a b c; do python -c "
import sys, json
l = json.loads(open('$O/r01j_$f.json').read().strip().splitlines()[-1])
print('$f', json.dumps({k: l[k] for k in ('value', 'ms_per_step', 'e2e', 'kernels_ms_per_step', 'gpu_launches')}))"; grep gcalign $O/r01j_$f.err | tail -16; done
