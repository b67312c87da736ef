#!/bin/bash
# r03t: durations of the S7 kernels (token streams, record sizes, GAM records, gather) of one 1678-read c2 batch
cd "$GRAFT_REPO_ROOT"
O=gpurun_out; mkdir -p $O
python -c "
from graphchainer_b200 import synth
print(synth.make_workload('c2', '/tmp/c2s', n_reads=1678))
"
D=graphchainer_b200/GraphChainerB200
$D -g /tmp/c2s.gfa --gc-save-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o.gam -t 16 --gc-streams 1 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gc_gam|gc_tokens|gc_piece|DeviceScan" --csv --log-file $O/r03t_s7_launches.csv $D --gc-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o.gam -t 16 --gc-streams 1 > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(l for l in open("gpurun_out/r03t_s7_launches.csv") if l.startswith('"'))]
h=rows[0]; k=h.index("Kernel Name"); v=h.index("Metric Value"); u=h.index("Metric Unit")
c=collections.defaultdict(lambda:[0,0.0])
for r in rows[1:]:
    t=float(r[v].replace(",","")); 
    if r[u]=="ns": t/=1e6
    elif r[u]=="us": t/=1e3
    name=r[k].split("(")[0]
    c[name][0]+=1; c[name][1]+=t
for name,(n,t) in sorted(c.items(), key=lambda x:-x[1][1]): print(f"{name:60s} n={n:4d} {t:9.3f} ms")
PY
