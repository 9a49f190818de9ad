O=gpurun_out/r02o; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_search.py -m gpu -x -q -k "temperature or rate_table" > $O/pytest_new.log 2>&1; echo "pytest new rc=$?" >> $O/summary.txt
for mode in 1 0 1 0; do
  OZ_CONV_STATIC=$mode timeout 300 python tools/single_game_latency.py --graphs 1 > $O/single_static$mode.log 2>> $O/err.log
  echo "static=$mode $(tail -1 $O/single_static$mode.log | cut -c1-200)" >> $O/table.txt
done
for mode in 1 0; do
  OZ_CONV_STATIC=$mode timeout 600 python bench.py --no-cpu-baseline --midgame-ply 0 --steps 3 --config4-steps 1 --config5-leaf 64 256 1024 > $O/bench_static$mode.log 2>> $O/err.log
  python - $O/bench_static$mode.log $mode >> $O/table.txt <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('static='+sys.argv[2], round(d['value']), [(p['leaf_batch_per_gpu'], round(p['value'])) for p in d['config5']['points']])
PY
done
