O=gpurun_out/r02af; mkdir -p $O
OZ_TOWER_DBG=0 timeout 120 python tools/tower_timeline.py > $O/timeline_0.log 2>&1
head -8 $O/timeline_0.log | cut -d'|' -f3-
for t in 0 1; do OZ_TOWER=$t timeout 300 python tools/single_game_latency.py --searches 30 > $O/single_tower$t.json 2>&1; tail -1 $O/single_tower$t.json; done
OZ_TOWER=1 timeout 300 python tools/single_game_latency.py --searches 30 --graphs 0 > $O/single_tower1_nograph.json 2>&1; tail -1 $O/single_tower1_nograph.json
timeout 900 python -m pytest tests/test_gpu_resnet.py tests/test_net_golden.py tests/test_gpu_search.py tests/test_cpp_binding.py -m gpu -x -q > $O/pytest.log 2>&1; tail -3 $O/pytest.log
