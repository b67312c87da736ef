cd "$GRAFT_REPO_ROOT"; O=gpurun_out; mkdir -p $O
nvidia-smi -L; nproc
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 3 > $O/r01_n2.json 2> $O/r01_n2.err
tail -1 $O/r01_n2.json | cut -c1-1500; tail -3 $O/r01_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > $O/r01_n2_ref.json 2> $O/r01_n2_ref.err
tail -1 $O/r01_n2_ref.json | cut -c1-600
D=graphchainer_b200/GraphChainerB200
python -c "
from graphchainer_b200 import synth
print(synth.make_workload('c2', '/tmp/c2m', n_reads=3000))
"
$D -g /tmp/c2m.gfa --gc-save-index /tmp/c2m.gcidx -f /tmp/c2m.fa -a /tmp/o1.gam -t 16 --gc-gpus 1 2>&1 | tail -1
$D --gc-index /tmp/c2m.gcidx -f /tmp/c2m.fa -a /tmp/o2.gam -t 16 --gc-gpus 2 2>&1 | tail -1
python -c "
from graphchainer_b200 import gam
a=gam.read_gam('/tmp/o1.gam'); b=gam.read_gam('/tmp/o2.gam'); print('driver 1 GPU vs 2 GPUs:', len(a), len(b), gam.diff_gam(a,b)[:2])"
