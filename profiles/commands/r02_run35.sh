O=gpurun_out/r02ai; mkdir -p $O
OZ_TOWER=1 timeout 300 python tools/single_game_latency.py --searches 30 > $O/single_tower1.json 2>&1; tail -1 $O/single_tower1.json
timeout 900 python -m pytest tests/test_gpu_search.py tests/test_cpp_binding.py tests/test_gpu_reference_callers.py -m gpu -x -q > $O/pytest.log 2>&1; tail -3 $O/pytest.log
