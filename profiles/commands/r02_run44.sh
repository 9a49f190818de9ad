O=gpurun_out/r02ar; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/summary.txt; tail -1 $O/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_1.log 2>&1; echo "pytest rc=$? $(tail -1 $O/pytest_1.log)" >> $O/summary.txt
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?" >> $O/summary.txt
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_1gpu.json 2> $O/bench_1gpu.err; echo "bench rc=$?" >> $O/summary.txt
for b in 256 1024 4096; do timeout 300 python tools/train_bench.py --batch $b --steps 30 >> $O/train_step.jsonl 2>> $O/train.err; done
timeout 600 python tools/reference_on_gpu.py > $O/reference_on_b200.json 2> $O/reference_on_b200.err; echo "refgpu rc=$?" >> $O/summary.txt
timeout 300 python tools/single_game_latency.py --searches 40 > $O/single.json 2>&1
timeout 300 python tools/single_game_latency.py --searches 20 --channels 256 --blocks 19 --hidden 256 --sims 1600 >> $O/single.json 2>&1
cat $O/summary.txt; tail -2 $O/single.json; tail -3 $O/train_step.jsonl | cut -c1-260; tail -2 $O/reference_on_b200.json | cut -c1-600
