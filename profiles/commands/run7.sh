set -x
O=gpurun_out/r01g; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_resnet.py -x -q > $O/pytest_resnet.log 2>&1; echo "pytest rc=$?" >> $O/summary.txt
for f in 7 0 1 2 4 6 3 5; do
  OZ_CONV_FLAGS=$f timeout 120 python tools/nn_forward.py --boards 32768 --iters 10 > $O/nn_flags$f.log 2>&1
done
