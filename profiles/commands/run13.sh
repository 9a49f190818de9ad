set -x
O=gpurun_out/r01m; mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/summary.txt
timeout 600 python bench.py > $O/bench.log 2> $O/bench.err; echo "bench rc=$?" >> $O/summary.txt
timeout 600 python bench.py --games 4096 --steps 2 --no-cpu-baseline --midgame-ply 0 > $O/bench_4096.log 2> $O/bench_4096.err; echo "bench4096 rc=$?" >> $O/summary.txt
timeout 600 python bench.py --games 256 --steps 4 --no-cpu-baseline --midgame-ply 0 > $O/bench_256.log 2> $O/bench_256.err; echo "bench256 rc=$?" >> $O/summary.txt
