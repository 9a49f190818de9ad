O=gpurun_out/r02p; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/summary.txt
timeout 300 python tools/single_game_latency.py --graphs 1 > $O/single.log 2>> $O/err.log; echo "single rc=$?" >> $O/summary.txt
timeout 900 python bench.py --steps 10 --warmup 5 > $O/bench.log 2> $O/bench.err; echo "bench rc=$?" >> $O/summary.txt
timeout 300 python tools/perft_bench.py --depth 12 > $O/perft.json 2> $O/perft.err; echo "perft rc=$?" >> $O/summary.txt
python - $O/bench.log >> $O/table.txt <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(round(d['value']), round(d['e2e']['value']), d['roofline']['frac'], d['perft']['nodes_per_s'], d['config4']['value'], [(p['leaf_batch_per_gpu'], round(p['value'])) for p in d['config5']['points']])
PY
