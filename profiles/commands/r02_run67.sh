O=gpurun_out/r02bo; mkdir -p $O
timeout 300 ncu --set full --clock-control none --import-source on -k regex:heads_kernel -s 30 -c 2 -o $O/prof_heads_narrow_128 python tools/single_game_latency.py --searches 2 > $O/ncu_heads_128.log 2>&1; echo "ncu heads128 rc=$?" >> $O/summary.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:heads_kernel -s 30 -c 2 -o $O/prof_heads_narrow_256 python bench.py --channels 256 --blocks 19 --hidden 256 --sims 1600 --games 8 --steps 1 --warmup 3 --no-extra-configs --no-cpu-baseline --midgame-ply 0 > $O/ncu_heads_256.log 2>&1; echo "ncu heads256 rc=$?" >> $O/summary.txt
cat $O/summary.txt
