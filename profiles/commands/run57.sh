O=gpurun_out/r01r1; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/summary.txt
timeout 600 python bench.py --steps 6 --no-cpu-baseline > $O/bench.log 2> $O/bench.err; echo "bench rc=$?" >> $O/summary.txt
timeout 600 python tools/selfplay_run.py --games 1000 --slots 1000 > $O/selfplay.log 2> $O/selfplay.err
