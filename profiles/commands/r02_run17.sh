O=gpurun_out/r02q; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_resnet.py tests/test_net_golden.py tests/test_cpp_binding.py -m gpu -x -q > $O/pytest_resnet.log 2>&1; echo "pytest resnet rc=$?" >> $O/summary.txt
timeout 900 python -m pytest tests/test_gpu_search.py tests/test_gpu_statistical.py tests/test_gpu_reference_callers.py -m gpu -x -q > $O/pytest_search.log 2>&1; echo "pytest search rc=$?" >> $O/summary.txt
for mode in 0 1 0 1; do
  OZ_TOWER=$mode timeout 300 python tools/single_game_latency.py --graphs 1 > $O/single_tower$mode.log 2>> $O/err.log
  echo "tower=$mode $(tail -1 $O/single_tower$mode.log | cut -c1-200)" >> $O/table.txt
done
for mode in 0 1; do
  OZ_TOWER=$mode timeout 600 python bench.py --no-cpu-baseline --midgame-ply 0 --steps 3 --config4-steps 1 --config5-leaf 64 128 256 512 > $O/bench_tower$mode.log 2>> $O/err.log
  python - $O/bench_tower$mode.log $mode >> $O/table.txt <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('tower='+sys.argv[2], round(d['value']), [(p['leaf_batch_per_gpu'], round(p['value'])) for p in d['config5']['points']])
PY
done
for mode in 0 1; do OZ_TOWER=$mode timeout 300 python tools/nn_forward.py --boards 32 64 128 256 >> $O/nn_forward_tower$mode.jsonl 2>> $O/err.log; done
