cd "$GRAFT_REPO_ROOT"; O=gpurun_out; mkdir -p $O
bash profiles/gpu_round.sh r01v
du -sh $O
