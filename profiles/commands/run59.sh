O=gpurun_out/r01p5; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"select_kernel|expand_kernel|encode_kernel|apply_kernel" -s 150 -c 8 -o $O/prof_tree_final python bench.py --steps 1 --warmup 3 --no-cpu-baseline --midgame-ply 0 > $O/ncu_tree.log 2>&1; echo "ncu tree rc=$?" >> $O/summary.txt
