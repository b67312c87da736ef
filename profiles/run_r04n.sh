#!/bin/bash
# r04n: the ragged-read fixtures on the GPU
cd "$GRAFT_REPO_ROOT"
O=gpurun_out; mkdir -p $O
( time timeout 600 python -m pytest tests/test_pipeline.py -m gpu -x -q -k "ragged" ) > $O/r04n_tests.log 2>&1
tail -4 $O/r04n_tests.log
