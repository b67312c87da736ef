cd "$GRAFT_REPO_ROOT"; O=gpurun_out; mkdir -p $O
bash profiles/gpu_round.sh r01u
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r01u_bench_reference.json 2> $O/r01u_bench_reference.err
tail -1 $O/r01u_bench_reference.json | cut -c1-300
