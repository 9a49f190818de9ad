set -x
O=gpurun_out/r01l; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/summary.txt
for f in 14 6; do OZ_CONV_FLAGS=$f timeout 120 python tools/nn_forward.py --boards 32000 --iters 10 > $O/nn_flags$f.log 2>&1; done
timeout 300 python tools/nn_forward.py --boards 8192 --iters 5 --channels 256 --blocks 19 --hidden 256 > $O/nn_256x20b.log 2>&1
OZ_CONV_VARIANT=1 timeout 300 python tools/nn_forward.py --boards 8192 --iters 5 --channels 256 --blocks 19 --hidden 256 > $O/nn_256x20b_v1.log 2>&1
timeout 600 python bench.py --steps 4 --warmup 3 > $O/bench.log 2> $O/bench.err; echo "bench rc=$?" >> $O/summary.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 21 -c 26 --csv --log-file $O/launches_nn.csv python tools/nn_forward.py --boards 32000 --iters 1 --warmup 1 > $O/ncu_nn.log 2>&1
