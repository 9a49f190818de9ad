O=gpurun_out/r02ab; mkdir -p $O
timeout 300 python tools/tower_check.py --iters 200 > $O/tower_check.log 2>&1; echo "tower_check rc=$?" >> $O/summary.txt
tail -25 $O/tower_check.log
