O=gpurun_out/r02bq; mkdir -p $O
timeout 400 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 $O/pytest_gpu.log)" >> $O/summary.txt
timeout 120 python tools/single_game_latency.py --searches 30 2>&1 | grep '^{' | cut -c1-260 > $O/single.json
timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra-configs --midgame-ply 24 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/summary.txt
cat $O/summary.txt; cat $O/single.json
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02bq/bench.json').read().strip().splitlines()[-1]); ph=d['phases_ms_per_step']
print(round(d['value']), round(d['ms_per_step'],2), ph, d['midgame'], d['clocks']['sm_mhz'])
P
