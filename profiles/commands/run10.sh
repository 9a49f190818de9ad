set -x
O=gpurun_out/r01j; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/summary.txt
timeout 600 python bench.py --steps 4 --warmup 3 > $O/bench.log 2> $O/bench.err; echo "bench rc=$?" >> $O/summary.txt
timeout 600 python tools/perft_bench.py --depth 12 --cpu > $O/perft.log 2>&1; echo "perft rc=$?" >> $O/summary.txt
timeout 600 python tools/perft_bench.py --depth 12 --mode 1 > $O/perft_mode1.log 2>&1
timeout 300 python tools/perft_bench.py --depth 11 --only > $O/perft_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dfs_kernel -c 1 -o $O/prof_perft python tools/perft_bench.py --depth 11 --only > $O/ncu_perft.log 2>&1; echo "ncu perft rc=$?" >> $O/summary.txt
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"select_kernel|expand_kernel|encode_kernel|apply_kernel" -s 180 -c 7 -o $O/prof_tree python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_tree.log 2>&1; echo "ncu tree rc=$?" >> $O/summary.txt
ls -la $O
