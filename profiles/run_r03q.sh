#!/bin/bash
# r03q: where a batch's wall time goes in the bench's warm e2e steps (6 batches in flight) vs one batch in flight
cd "$GRAFT_REPO_ROOT"
O=gpurun_out; mkdir -p $O
GCGPU_TRACE=1 GC_TRACE=1 timeout 900 python bench.py --no-cpu-baseline > $O/r03q_bench_traced.json 2> $O/r03q_trace_s6.txt
GCGPU_TRACE=1 GC_TRACE=1 timeout 900 python bench.py --no-cpu-baseline --streams 1 > $O/r03q_bench_traced_s1.json 2> $O/r03q_trace_s1.txt
python - <<'PY'
import re,collections,json
for S in ("s6","s1"):
    lines=open(f"gpurun_out/r03q_trace_{S}.txt", errors="replace").read().split("\n")
    gam=[i for i,l in enumerate(lines) if l.startswith("[gc] phase gam")]
    # 36 batches of the e2e part (warm-up 18 + timed 18), then 6 whole-set batches
    lo,hi=gam[17]+1,gam[35]+1
    ph=collections.defaultdict(lambda:[0,0,0,0]); k=collections.defaultdict(lambda:[0,0])
    for l in lines[lo:hi]:
        m=re.match(r"\[gc\] phase (\S+)\s+([\d.]+) ms(?: \(libgcgpu calls ([\d.]+) ms, host ([\d.]+) ms\))?", l)
        if m:
            p=ph[m.group(1)]; p[0]+=1; p[1]+=float(m.group(2)); p[2]+=float(m.group(3) or 0); p[3]+=float(m.group(4) or 0)
        m=re.match(r"\[gcgpu\] (.+?)\s+n=(\d+)\s+([\d.]+) ms", l)
        if m:
            kk=k[m.group(1).strip()]; kk[0]+=1; kk[1]+=float(m.group(3))
    d=json.loads([l for l in open(f"gpurun_out/r03q_bench_traced{'' if S=='s6' else '_s1'}.json") if l.startswith("{")][-1])
    print(f"== {S}: e2e {d['e2e']['value']/1e6:.1f} Mbp/s ({d['e2e']['ms_per_step']:.0f} ms/step, tracing on); 18 timed batches: phase (count, wall ms per batch, in libgcgpu, host)")
    for name,v in ph.items(): print(f"  {name:8s} n={v[0]:3d} wall {v[1]/18:8.1f}  gpu-calls {v[2]/18:8.1f}  host {v[3]/18:8.1f}")
    print("  per batch: wall %.1f gpu-calls %.1f host %.1f" % (sum(v[1] for v in ph.values())/18, sum(v[2] for v in ph.values())/18, sum(v[3] for v in ph.values())/18))
    print(f"  kernel groups (CUDA-event time per batch, incl. waiting for SMs)")
    for name,v in sorted(k.items(), key=lambda x:-x[1][1])[:12]: print(f"    {name:40s} n={v[0]:4d} {v[1]/18:8.1f} ms")
PY
