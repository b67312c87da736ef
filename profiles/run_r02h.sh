cd "$GRAFT_REPO_ROOT"; O=gpurun_out; mkdir -p $O
bash profiles/gpu_round.sh r02h
D=graphchainer_b200/GraphChainerB200
for spec in "gc_k1_kernel:4" "gc_k1_bt_kernel:4"; do K=${spec%%:*}; S=${spec##*:}
timeout 600 ncu --set full --clock-control none -k regex:"$K" -s $S -c 1 -o $O/r02h_${K}_s2 -f $D --gc-index /tmp/c2s.gcidx -f /tmp/c2s.fa -a /tmp/o.gam -t 16 --gc-streams 1 > $O/r02h_${K}_s2.log 2>&1
echo "## S2 launch (512 k fragment items)" >> $O/r02h_ncu_full_summary.txt
python profiles/ncu_summary.py kernel $O/r02h_${K}_s2.ncu-rep >> $O/r02h_ncu_full_summary.txt
rm -f $O/r02h_${K}_s2.ncu-rep
done
du -sh $O
