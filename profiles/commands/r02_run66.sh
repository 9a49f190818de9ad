O=gpurun_out/r02bn; mkdir -p $O
for i in 1 2; do timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$i.log 2>&1; echo "pytest $i rc=$? $(tail -1 $O/pytest_gpu_$i.log)" >> $O/summary.txt; done
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/summary.txt
timeout 600 python tools/selfplay_run.py > $O/selfplay_1000.json 2> $O/selfplay.err; echo "selfplay rc=$?" >> $O/summary.txt
timeout 600 python tools/match_run.py --pairs 250 --sims 3200 > $O/match_3200.json 2> $O/match.err; echo "match rc=$?" >> $O/summary.txt
timeout 600 python tools/iteration_run.py --iterations 2 > $O/iteration.jsonl 2> $O/iteration.err; echo "iteration rc=$?" >> $O/summary.txt
for b in 256 4096; do timeout 300 python tools/train_bench.py --batch $b --steps 30 --skip-pytorch >> $O/train_step.jsonl 2>> $O/train.err; done
cat $O/summary.txt; tail -1 $O/selfplay_1000.json | cut -c1-400; tail -1 $O/match_3200.json | cut -c1-400; tail -2 $O/iteration.jsonl | cut -c1-500; cut -c1-300 $O/train_step.jsonl
