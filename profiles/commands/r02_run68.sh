O=gpurun_out/r02bp; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 $O/pytest_gpu.log)" >> $O/summary.txt
timeout 300 python tools/single_game_latency.py --searches 30 2>&1 | grep '^{' | cut -c1-260 > $O/single.json
timeout 300 python tools/single_game_latency.py --searches 15 --channels 256 --blocks 19 --hidden 256 --sims 1600 2>&1 | grep '^{' | cut -c1-260 >> $O/single.json
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --midgame-ply 0 --config5-leaf 64 256 1024 --config5-steps 8 --config4-steps 2 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/summary.txt
cat $O/summary.txt; cat $O/single.json
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02bp/bench.json').read().strip().splitlines()[-1]); ph=d['phases_ms_per_step']
print(round(d['value']), round(d['ms_per_step'],2), 'heads', round(ph['network']-ph['of_which_conv'],3), ph, d['clocks']['sm_mhz'])
print([(p['leaf_batch_per_gpu'], round(p['value'])) for p in d['config5']['points']], d['config3_leaf_batch_256']['value'])
P
