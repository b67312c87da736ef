cd "$GRAFT_REPO_ROOT"; O=gpurun_out; mkdir -p $O
bash profiles/gpu_round.sh r02d
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r02d_bench_reference.json 2> $O/r02d_bench_reference.err
du -sh $O
