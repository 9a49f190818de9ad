O=gpurun_out/r02ay; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_1.log 2>&1; echo "pytest rc=$? $(tail -1 $O/pytest_1.log)" >> $O/summary.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/summary.txt
cat $O/summary.txt
python - <<PY
import json
d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1])
print(d['value'], d['phases_ms_per_step'], d['config3_leaf_batch_256']['value'], [round(p['value']) for p in d['config5']['points']])
PY
