#!/bin/bash
# r03p: where a batch's wall time goes when 6 batches are in flight (c2, 10000 reads, 16 host threads): phase times per batch, 1 stream vs 6 streams
cd "$GRAFT_REPO_ROOT"
O=gpurun_out; mkdir -p $O
python -c "
from graphchainer_b200 import synth
print(synth.make_workload('c2', '/tmp/c2f', n_reads=10000))
"
D=graphchainer_b200/GraphChainerB200
$D -g /tmp/c2f.gfa --gc-save-index /tmp/c2f.gcidx -f /tmp/c2f.fa -a /tmp/o.gam -t 16 > /dev/null 2>&1
for S in 1 6; do
GCGPU_TRACE=1 GC_TRACE=1 GC_TRACE_CALL=1 $D --gc-index /tmp/c2f.gcidx -f /tmp/c2f.fa -a /tmp/o_$S.gam -t 16 --gc-streams $S > $O/r03p_trace_streams$S.txt 2>&1
grep "B200:" $O/r03p_trace_streams$S.txt
done
python - <<'PY'
import re,collections
for S in (1,6):
    ph=collections.defaultdict(lambda:[0,0,0,0]); k=collections.defaultdict(lambda:[0,0])
    for l in open(f"gpurun_out/r03p_trace_streams{S}.txt", errors="replace"):
        m=re.match(r"\[gc\] phase (\S+)\s+([\d.]+) ms(?: \(libgcgpu calls ([\d.]+) ms, host ([\d.]+) ms\))?", l)
        if m:
            p=ph[m.group(1)]; p[0]+=1; p[1]+=float(m.group(2)); p[2]+=float(m.group(3) or 0); p[3]+=float(m.group(4) or 0)
        m=re.match(r"\[gcgpu\] (.+?)\s+n=(\d+)\s+([\d.]+) ms", l)
        if m:
            kk=k[m.group(1).strip()]; kk[0]+=1; kk[1]+=float(m.group(3))
    print(f"== streams {S}: phases (count, total ms, in libgcgpu, host)")
    for name,v in ph.items(): print(f"  {name:8s} n={v[0]:3d} total {v[1]:8.1f}  gpu-calls {v[2]:8.1f}  host {v[3]:8.1f}")
    print("  sum total %.1f gpu-calls %.1f host %.1f" % (sum(v[1] for v in ph.values()), sum(v[2] for v in ph.values()), sum(v[3] for v in ph.values())))
    print(f"== streams {S}: kernel groups (event time incl. waiting for SMs)")
    for name,v in sorted(k.items(), key=lambda x:-x[1][1])[:14]: print(f"  {name:40s} n={v[0]:4d} {v[1]:8.1f} ms")
PY
