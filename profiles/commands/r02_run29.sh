O=gpurun_out/r02ac; mkdir -p $O
for f in 0 1 2 3; do OZ_TOWER_DBG=$f timeout 120 python tools/tower_timeline.py > $O/timeline_$f.log 2>&1; done
head -24 $O/timeline_0.log; grep "us per forward" $O/timeline_*.log
