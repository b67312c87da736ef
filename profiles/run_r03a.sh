cd "$GRAFT_REPO_ROOT"; bash profiles/gpu_round.sh r03a skip-ncu
