cd "$GRAFT_REPO_ROOT"; O=gpurun_out; mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -x -q ) > $O/r01b_tests.log 2>&1; tail -3 $O/r01b_tests.log
( GCGPU_K1_SIMT_MIN=1 timeout 900 python -m pytest tests -m gpu -x -q ) > $O/r01b_tests_simt.log 2>&1; tail -3 $O/r01b_tests_simt.log
bash profiles/sweep.sh r01b "8388608 33554432 134217728" "4 2"
