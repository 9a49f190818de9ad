O=gpurun_out/r02ae; mkdir -p $O
OZ_TOWER_DBG=0 timeout 120 python tools/tower_timeline.py > $O/timeline_0.log 2>&1
head -12 $O/timeline_0.log | cut -d'|' -f3-
