set -x
O=gpurun_out/r01i; mkdir -p $O
timeout 600 python bench.py --steps 4 --warmup 3 > $O/bench.log 2> $O/bench.err; echo "bench rc=$?" >> $O/summary.txt
timeout 600 python tools/perft_bench.py --depth 13 --cpu > $O/perft.log 2>&1; echo "perft rc=$?" >> $O/summary.txt
timeout 600 python tools/perft_bench.py --depth 12 --mode 1 > $O/perft_mode1.log 2>&1
timeout 300 python tools/perft_bench.py --depth 11 > $O/perft_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dfs_kernel|bfs_" -o $O/prof_perft python tools/perft_bench.py --depth 11 > $O/ncu_perft.log 2>&1; echo "ncu perft rc=$?" >> $O/summary.txt
