set -x
O=gpurun_out/r01n1; mkdir -p $O
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --midgame-ply 0 > $O/ncu.log 2>&1; echo "ncu rc=$?" >> $O/summary.txt
