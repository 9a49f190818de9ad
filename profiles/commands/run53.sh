O=gpurun_out/r01f2; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/summary.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/summary.txt
timeout 600 python bench.py > $O/bench.log 2> $O/bench.err; echo "bench rc=$?" >> $O/summary.txt
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.log 2> $O/bench_ref.err; echo "bench ref rc=$?" >> $O/summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --midgame-ply 0 > $O/ncu.log 2>&1; echo "ncu rc=$?" >> $O/summary.txt
timeout 300 python tools/single_game_latency.py --graphs 1 > $O/single_g1.log 2> $O/single_g1.err
timeout 600 python tools/match_run.py --pairs 500 > $O/match.log 2> $O/match.err
