set -x
mkdir -p gpurun_out/r01b
O=gpurun_out/r01b
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/gpu.txt
nproc >> $O/gpu.txt
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/summary.txt
python bench.py --steps 3 --warmup 3 > $O/bench.log 2> $O/bench.err; echo "bench rc=$?" >> $O/summary.txt
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 700 --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_launch.log 2>&1; echo "ncu launches rc=$?" >> $O/summary.txt
python tools/nn_forward.py --boards 32768 --iters 3 --warmup 1 > $O/nn_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv3x3 -s 24 -c 2 -o $O/prof_conv_full python tools/nn_forward.py --boards 32768 --iters 3 --warmup 1 > $O/ncu_conv.log 2>&1; echo "ncu conv rc=$?" >> $O/summary.txt
