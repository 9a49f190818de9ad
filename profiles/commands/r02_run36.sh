O=gpurun_out/r02aj; mkdir -p $O
timeout 300 python tools/tower_check.py --iters 200 > $O/tower_check.log 2>&1; echo "tower_check rc=$?" >> $O/summary.txt; tail -1 $O/tower_check.log
grep '"boards": 32' $O/tower_check.log | cut -c1-120
for t in 0 1; do OZ_TOWER=$t timeout 300 python tools/single_game_latency.py --searches 30 > $O/single_tower$t.json 2>&1; tail -1 $O/single_tower$t.json; done
timeout 300 python tools/single_game_latency.py --searches 20 --channels 256 --blocks 19 --hidden 256 --sims 1600 > $O/single_256x20b.json 2>&1; tail -1 $O/single_256x20b.json
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -3 $O/pytest.log
