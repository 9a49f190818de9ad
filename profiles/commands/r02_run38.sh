O=gpurun_out/r02al; mkdir -p $O
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_1gpu.json 2> $O/bench_1gpu.err; echo "bench rc=$?" >> $O/summary.txt; tail -c 6000 $O/bench_1gpu.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_pair -s 24 -c 2 -o $O/prof_conv256 python tools/nn_forward.py --boards 16384 --channels 256 --blocks 19 --hidden 256 --iters 3 --warmup 1 > $O/ncu_conv256.log 2>&1; echo "ncu conv256 rc=$?" >> $O/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tower_cluster -s 5 -c 1 -o $O/prof_tower python tools/single_game_latency.py --searches 1 > $O/ncu_tower.log 2>&1; echo "ncu tower rc=$?" >> $O/summary.txt
cat $O/summary.txt
