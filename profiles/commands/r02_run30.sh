O=gpurun_out/r02ad; mkdir -p $O
timeout 300 python tools/tower_check.py --iters 200 > $O/tower_check.log 2>&1; echo "tower_check rc=$?" >> $O/summary.txt
OZ_TOWER_DBG=0 timeout 120 python tools/tower_timeline.py > $O/timeline_0.log 2>&1
grep -v '"boards": 1,\|"boards": 33\|"boards": 16' $O/tower_check.log | cut -c1-200; head -14 $O/timeline_0.log
