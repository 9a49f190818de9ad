O=gpurun_out/r02bf; mkdir -p $O
OZ_TOWER_DBG=0 timeout 120 python tools/tower_timeline.py > $O/timeline.log 2>&1; head -22 $O/timeline.log | cut -c1-330
