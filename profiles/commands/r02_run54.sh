O=gpurun_out/r02bb; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_1.log 2>&1; echo "pytest rc=$? $(tail -1 $O/pytest_1.log)" >> $O/summary.txt; cat $O/summary.txt
timeout 300 python tools/single_game_latency.py --searches 40 > $O/single.json 2>&1; tail -1 $O/single.json
timeout 600 python bench.py --steps 10 --warmup 3 --no-extra-configs --no-cpu-baseline > $O/bench.json 2> $O/bench.err; python - <<PY
import json
d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['phases_ms_per_step'], d['clocks']['sm_mhz'], d['midgame'])
PY
