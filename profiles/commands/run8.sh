set -x
O=gpurun_out/r01h; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/summary.txt
timeout 600 python bench.py --steps 4 --warmup 3 > $O/bench.log 2> $O/bench.err; echo "bench rc=$?" >> $O/summary.txt
timeout 300 python bench.py --impl reference --steps 4 --warmup 1 > $O/bench_ref.log 2> $O/bench_ref.err; echo "bench ref rc=$?" >> $O/summary.txt
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 700 --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_launch.log 2>&1; echo "ncu launches rc=$?" >> $O/summary.txt
timeout 120 python tools/nn_forward.py --boards 32768 --iters 3 --warmup 1 > $O/nn_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv3x3|heads" -s 22 -c 4 -o $O/prof_pair python tools/nn_forward.py --boards 32768 --iters 3 --warmup 1 > $O/ncu_pair.log 2>&1; echo "ncu pair rc=$?" >> $O/summary.txt
