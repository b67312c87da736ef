#!/bin/bash
# ncu capture of the K1 / K3 kernels on a c2 sample (839 reads): run on the GPU box via gpurun
set -e
cd "$GRAFT_REPO_ROOT"
python -c "
from graphchainer_b200 import synth
print(synth.make_workload('c2', '/tmp/c2s', n_reads=839))
"
D=graphchainer_b200/GraphChainerB200
$D -g /tmp/c2s.gfa --gc-save-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o.gam -t 16 --gc-streams 1 | tail -2
ncu --set full --clock-control none --import-source on -k regex:"$1" -s ${2:-0} -c ${3:-1} -o gpurun_out/$4 -f $D --gc-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o.gam -t 16 --gc-streams 1 > gpurun_out/$4.log 2>&1
tail -3 gpurun_out/$4.log
