#!/bin/bash
# streams x host-threads-per-stream x batch sweep of bench.py on the GPU box: bash profiles/sweep2.sh <tag> "<streams:tps:batch_bp> ..."
cd "$GRAFT_REPO_ROOT"
TAG=$1; O=gpurun_out; mkdir -p $O
for cfg in $2; do
	IFS=: read st tps bb <<< "$cfg"
	echo "== streams=$st threads_per_stream=$tps batch_bp=$bb" >> $O/${TAG}_sweep.txt
	timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --batch-bp $bb --streams $st --threads-per-stream $tps $3 2>&1 | tail -1 | python -c "
import sys, json
l = json.loads(sys.stdin.readline())
print(json.dumps({k: l[k] for k in ('value', 'ms_per_step', 'e2e', 'kernels_ms_per_step', 'gpu_launches')}))" >> $O/${TAG}_sweep.txt 2>&1
done
cat $O/${TAG}_sweep.txt
