#!/bin/bash
# r04h: run-to-run determinism of the GAM records over many steps (after r04g's single differing read)
cd "$GRAFT_REPO_ROOT"
O=gpurun_out; mkdir -p $O
timeout 420 python profiles/stress_determinism.py mixed 200 c2 > $O/r04h_mixed_long.txt 2> $O/r04h_mixed_long.err
grep -c . $O/r04h_mixed_long.txt; grep DIFF $O/r04h_mixed_long.txt | head -5 | cut -c1-1200; tail -1 $O/r04h_mixed_long.txt | cut -c1-300; tail -2 $O/r04h_mixed_long.err | cut -c1-300
