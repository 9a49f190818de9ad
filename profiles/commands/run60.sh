O=gpurun_out/r01e1; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/summary.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/summary.txt
timeout 600 python bench.py --steps 6 --no-cpu-baseline > $O/bench.log 2> $O/bench.err; echo "bench rc=$?" >> $O/summary.txt
timeout 300 python tools/single_game_latency.py --graphs 1 > $O/single_g1.log 2> $O/single_g1.err
