O=gpurun_out/r01c1; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_resnet.py -x -q > $O/pytest_resnet.log 2>&1; echo "resnet rc=$?" >> $O/summary.txt
for c in 0 1; do
  OZ_CONV_CHAIN=$c timeout 300 python tools/nn_forward.py --boards 32000 --iters 20 > $O/nn_chain$c.log 2>&1
  OZ_CONV_CHAIN=$c timeout 300 python tools/nn_forward.py --boards 32768 --iters 20 > $O/nn32768_chain$c.log 2>&1
  OZ_CONV_CHAIN=$c timeout 300 python bench.py --steps 4 --no-cpu-baseline --midgame-ply 0 > $O/bench_chain$c.log 2> $O/bench_chain$c.err
done
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "all rc=$?" >> $O/summary.txt
