#!/bin/bash
# batch size / streams sweep of bench.py on the GPU box: bash profiles/sweep.sh <tag> "<batch_bp list>" "<streams list>" [extra bench args]
cd "$GRAFT_REPO_ROOT"
TAG=$1; O=gpurun_out; mkdir -p $O
for bb in $2; do for st in $3; do
	echo "== batch_bp=$bb streams=$st" >> $O/${TAG}_sweep.txt
	timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --batch-bp $bb --streams $st $4 2>&1 | tail -1 | python -c "
import sys, json
l = json.loads(sys.stdin.readline())
print(json.dumps({k: l[k] for k in ('value', 'ms_per_step', 'e2e', 'kernels_ms_per_step', 'gpu_launches')}))" >> $O/${TAG}_sweep.txt 2>&1
done; done
cat $O/${TAG}_sweep.txt
