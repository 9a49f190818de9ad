O=gpurun_out/r02m; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/summary.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/summary.txt
for b in 256 1024 4096; do timeout 600 python tools/train_bench.py --batch $b --steps 30 >> $O/train_bench.jsonl 2>> $O/train_bench.err; echo "train_bench $b rc=$?" >> $O/summary.txt; done
timeout 900 python bench.py --steps 10 --warmup 5 > $O/bench.log 2> $O/bench.err; echo "bench rc=$?" >> $O/summary.txt
timeout 300 python bench.py --impl reference --steps 10 --warmup 5 > $O/bench_ref.log 2> $O/bench_ref.err; echo "bench ref rc=$?" >> $O/summary.txt
timeout 300 python tools/perft_bench.py --depth 12 > $O/perft.json 2> $O/perft.err; echo "perft rc=$?" >> $O/summary.txt
timeout 300 python tools/single_game_latency.py --graphs 1 > $O/single.log 2> $O/single.err; echo "single rc=$?" >> $O/summary.txt
