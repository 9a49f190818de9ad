O=gpurun_out/r01i2; mkdir -p $O
timeout 1200 python tools/iteration_run.py --iterations 3 > $O/iteration.log 2> $O/iteration.err; echo "iteration rc=$?" >> $O/summary.txt
timeout 600 python tools/selfplay_run.py --games 2000 --slots 500 --samples > $O/rolling.log 2> $O/rolling.err; echo "rolling rc=$?" >> $O/summary.txt
