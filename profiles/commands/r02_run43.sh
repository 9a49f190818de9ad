O=gpurun_out/r02aq; mkdir -p $O
timeout 300 python tools/single_game_latency.py --searches 40 > $O/single.json 2>&1; tail -1 $O/single.json
timeout 1500 python -m pytest tests/test_gpu_search.py tests/test_gpu_statistical.py tests/test_cpp_binding.py tests/test_gpu_reference_callers.py -m gpu -x -q > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-extra-configs --no-cpu-baseline > $O/bench.json 2> $O/bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02aq/bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['phases_ms_per_step'], d['clocks']['sm_mhz'])
PY
