set -x
O=gpurun_out/r01e; mkdir -p $O
timeout 600 python bench.py --steps 3 --warmup 3 > $O/bench.log 2> $O/bench.err; echo "bench rc=$?" >> $O/summary.txt
timeout 120 python tools/nn_forward.py --boards 32768 --iters 3 --warmup 1 > $O/nn_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3 -s 24 -c 2 -o $O/prof_pair python tools/nn_forward.py --boards 32768 --iters 3 --warmup 1 > $O/ncu_pair.log 2>&1; echo "ncu pair rc=$?" >> $O/summary.txt
