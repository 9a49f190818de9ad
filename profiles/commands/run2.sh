set -x
O=gpurun_out/r01c; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_resnet.py -x -q > $O/pytest_resnet.log 2>&1; echo "pytest resnet rc=$?" >> $O/summary.txt
timeout 120 python tools/nn_forward.py --boards 32768 --iters 10 > $O/nn_pair.log 2>&1; echo "nn pair rc=$?" >> $O/summary.txt
OZ_CONV_VARIANT=1 timeout 120 python tools/nn_forward.py --boards 32768 --iters 10 > $O/nn_v1.log 2>&1; echo "nn v1 rc=$?" >> $O/summary.txt
timeout 120 python tools/nn_forward.py --boards 4096 --iters 20 > $O/nn_pair_4096.log 2>&1
