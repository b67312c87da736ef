#!/bin/bash
# r04m: GPU tests after --no-colinear-chaining and the single-threaded live reference runs
cd "$GRAFT_REPO_ROOT"
O=gpurun_out; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/r04m_tests.log 2>&1
tail -4 $O/r04m_tests.log
