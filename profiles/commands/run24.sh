set -x
O=gpurun_out/r01x; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/summary.txt
timeout 600 python tools/match_run.py --pairs 500 --sims 800 > $O/match_500.log 2> $O/match_500.err; echo "match rc=$?" >> $O/summary.txt
timeout 600 python bench.py --games 4096 --steps 2 --no-cpu-baseline --midgame-ply 0 > $O/bench_g4096.log 2> $O/bench_g4096.err; echo "bench4096 rc=$?" >> $O/summary.txt
timeout 600 python bench.py --games 2048 --steps 2 --no-cpu-baseline --midgame-ply 0 > $O/bench_g2048.log 2> $O/bench_g2048.err; echo "bench2048 rc=$?" >> $O/summary.txt
