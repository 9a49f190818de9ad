O=gpurun_out/r02at; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_1.log 2>&1; echo "pytest rc=$? $(tail -1 $O/pytest_1.log)" >> $O/summary.txt
timeout 900 python tools/selfplay_run.py > $O/selfplay_1000.json 2> $O/selfplay.err; echo "selfplay rc=$?" >> $O/summary.txt
cat $O/summary.txt; tail -1 $O/selfplay_1000.json | cut -c1-700
