O=gpurun_out/r02a; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/summary.txt
timeout 300 python tools/resnet_error.py > $O/resnet_error.json 2> $O/resnet_error.err; echo "resnet_error rc=$?" >> $O/summary.txt
timeout 900 python bench.py > $O/bench.log 2> $O/bench.err; echo "bench rc=$?" >> $O/summary.txt
timeout 300 ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,gpu__time_duration.sum --clock-control none -k regex:dfs_kernel --csv --log-file $O/perft_ncu.csv python tools/perft_bench.py --depth 11 --only > $O/perft_ncu.log 2>&1; echo "perft ncu rc=$?" >> $O/summary.txt
