#!/bin/bash
# r04d: 2 GPUs, the driver's launch line (torchrun, one rank per GPU): our arm and the reference arm; default workload at N > 1 = c3
cd "$GRAFT_REPO_ROOT"
O=gpurun_out; mkdir -p $O
nvidia-smi -L > $O/r04d_smi.txt; nproc >> $O/r04d_smi.txt
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 ) > $O/r04d_bench_n2.json 2> $O/r04d_bench_n2.err
tail -2 $O/r04d_bench_n2.err
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 ) > $O/r04d_bench_n2_reference.json 2> $O/r04d_bench_n2_reference.err
tail -4 $O/r04d_bench_n2_reference.err
python - <<'PY'
import json
for f in ("gpurun_out/r04d_bench_n2.json","gpurun_out/r04d_bench_n2_reference.json"):
    try:
        line=[l for l in open(f) if l.startswith("{")][-1]
        d=json.loads(line)
        print(f, "value %.1f Mbp/s e2e %.1f Mbp/s" % ((d["value"] or 0)/1e6, (d["e2e"]["value"] or 0)/1e6), d["config"]["workload"], d.get("parity_on_sample"), (d.get("cpu_baseline") or {}).get("cores"))
    except Exception as e: print(f, "failed", e)
PY
