O=gpurun_out/r02ah; mkdir -p $O
OZ_TOWER=1 timeout 300 python tools/single_game_latency.py --searches 30 > $O/single_tower1.json 2>&1; tail -1 $O/single_tower1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 1500 --csv --log-file $O/single_launches_warm.csv python tools/single_game_latency.py --searches 2 > $O/ncu.log 2>&1; echo "ncu rc=$?" >> $O/summary.txt
