O=gpurun_out/r02ap; mkdir -p $O
timeout 300 python tools/single_game_latency.py --searches 40 > $O/single.json 2>&1; tail -1 $O/single.json
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -3 $O/pytest.log
