set -x
O=gpurun_out/r01f; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/summary.txt
timeout 120 python tools/nn_forward.py --boards 32768 --iters 10 > $O/nn_pair.log 2>&1; echo "nn pair rc=$?" >> $O/summary.txt
timeout 120 python tools/nn_forward.py --boards 4096 --iters 20 > $O/nn_pair_4096.log 2>&1
timeout 120 python tools/nn_forward.py --boards 32000 --iters 10 > $O/nn_pair_32000.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 21 -c 42 --csv --log-file $O/launches_nn.csv python tools/nn_forward.py --boards 32768 --iters 1 --warmup 1 > $O/ncu_nn.log 2>&1
