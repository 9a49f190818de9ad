O=gpurun_out/r01i1; mkdir -p $O
timeout 1200 python tools/iteration_run.py --iterations 2 > $O/iteration.log 2> $O/iteration.err; echo "iteration rc=$?" >> $O/summary.txt
