O=gpurun_out/r02bi; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 $O/pytest_gpu.log)" >> $O/summary.txt
timeout 300 python tools/single_game_latency.py --searches 40 > $O/single_128.json 2>&1
timeout 300 python tools/single_game_latency.py --searches 20 --channels 256 --blocks 19 --hidden 256 --sims 1600 > $O/single_256.json 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config5-leaf 64 256 1024 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/summary.txt
cat $O/summary.txt; tail -5 $O/pytest_gpu.log
for f in $O/single_*.json; do grep '^{' $f | cut -c1-260; done
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02bi/bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['phases_ms_per_step'], d['midgame'])
print([ (p['leaf_batch_per_gpu'], round(p['value'])) for p in d['config5']['points']], d['config3_leaf_batch_256']['value'], d['config4']['value'])
P
