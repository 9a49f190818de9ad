O=gpurun_out/r02ag; mkdir -p $O
timeout 300 python tools/tower_check.py --iters 200 > $O/tower_check.log 2>&1; echo "tower_check rc=$?" >> $O/summary.txt; tail -1 $O/tower_check.log
grep '"boards": 32' $O/tower_check.log | cut -c1-120
OZ_TOWER_DBG=0 timeout 120 python tools/tower_timeline.py > $O/timeline_0.log 2>&1
head -8 $O/timeline_0.log | cut -d'|' -f3-
OZ_TOWER=1 timeout 300 python tools/single_game_latency.py --searches 30 > $O/single_tower1.json 2>&1; tail -1 $O/single_tower1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/single_launches.csv python tools/single_game_latency.py --searches 2 > $O/ncu.log 2>&1; echo "ncu rc=$?" >> $O/summary.txt
